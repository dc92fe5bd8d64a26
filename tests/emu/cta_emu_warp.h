// tests/emu/cta_emu_warp.h -- the one warp shuffle the product's kernels use, for builds on the CTA emulator (include
// after cta_emu.h).  __shfl_down_sync(full mask, v, off) appears only in the kernels' final counter reductions, which
// every thread of the block executes the same number of times: the values are exchanged through a block-wide buffer
// between two block barriers.  Test infrastructure.
#pragma once
static unsigned long long emu_shfl_buf[EMU_MAX_CLUSTER][2048];
static inline unsigned long long __shfl_down_sync(unsigned, unsigned long long v, int off) {
    unsigned long long* b = emu_shfl_buf[emu_block_slot];
    const unsigned t = threadIdx.x, lane = t & 31u;
    b[t] = v;
    __syncthreads();
    const unsigned long long r = (lane + (unsigned)off < 32u && t + (unsigned)off < blockDim.x) ? b[t + off] : v;
    __syncthreads();
    return r;
}
