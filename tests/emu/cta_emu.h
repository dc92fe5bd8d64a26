// tests/emu/cta_emu.h -- runs CUDA thread blocks on the host: one pthread per CUDA thread, a pthread barrier per block
// for __syncthreads(), one buffer per block for dynamic shared memory, function-local statics for `__shared__`
// variables.  Three launch forms: one block; a grid whose blocks do not communicate (run one after another); one
// thread-block CLUSTER of up to 8 blocks running concurrently, with a cluster-wide barrier and distributed-shared-memory
// address mapping (ms_cluster_sync / ms_cluster_rank / ms_cluster_map -- the kernels use the same three names on the
// device, where they wrap cooperative_groups::this_cluster()).
// Test infrastructure: lets kernels written with nothing but threadIdx / blockIdx / blockDim / __syncthreads / shared
// memory (no warp intrinsics, no atomics) be executed and checked on machines without a GPU.  Not a performance model
// and not a race detector.  Kernels launched as a cluster must not declare static `__shared__` variables (the
// emulation would share them between the blocks).
#pragma once
#include <cmath>
#include <cstring>
#include <pthread.h>
#include <vector>

#define MS_CTA_EMU 1
#define __global__
#define __device__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __cluster_dims__(...)

constexpr unsigned EMU_MAX_CLUSTER = 8;
constexpr size_t EMU_SMEM_BYTES = 232448;                     // 227 KB, the per-block limit of sm_100
alignas(16) static unsigned char emu_dyn_smem[EMU_MAX_CLUSTER][EMU_SMEM_BYTES];
#define MS_DYN_SMEM(name) unsigned char* name = emu_dyn_smem[emu_block_slot]

struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 threadIdx, blockIdx;
static thread_local unsigned emu_block_slot = 0;              // which shared-memory buffer / block barrier this thread uses
static emu_dim3 blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
static pthread_barrier_t emu_block_barrier[EMU_MAX_CLUSTER], emu_cluster_barrier;
static unsigned emu_cluster_size = 1;

static inline void __syncthreads() { pthread_barrier_wait(&emu_block_barrier[emu_block_slot]); }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }

// cluster primitives (device versions: ms_sd_train_cluster.cuh)
static inline unsigned ms_cluster_rank() { return emu_block_slot; }
static inline void ms_cluster_sync() { pthread_barrier_wait(&emu_cluster_barrier); }
template <class T>
static inline T* ms_cluster_map(T* p, unsigned rank) {
    size_t off = (size_t)((unsigned char*)p - emu_dyn_smem[emu_block_slot]);
    return (T*)(emu_dyn_smem[rank] + off);
}

template <class Kernel, class Args>
struct emu_launch_ctx { Kernel k; const Args* a; unsigned tid, block, block_y, slot; };

template <class Kernel, class Args>
static void* emu_thread_main(void* p) {
    auto* c = static_cast<emu_launch_ctx<Kernel, Args>*>(p);
    threadIdx = {c->tid, 0, 0};
    blockIdx = {c->block, c->block_y, 0};
    emu_block_slot = c->slot;
    c->k(*c->a);
    return nullptr;
}

// `blocks` concurrent blocks (slots 0 .. blocks-1) of `threads` threads each; block b reports blockIdx.x = first_block + b
template <class Kernel, class Args>
static int emu_run_blocks(Kernel k, const Args& args, unsigned first_block, unsigned blocks, unsigned threads,
                          unsigned block_y = 0) {
    if (blocks < 1 || blocks > EMU_MAX_CLUSTER) return -3;
    blockDim = {threads, 1, 1};
    emu_cluster_size = blocks;
    for (unsigned b = 0; b < blocks; ++b)
        if (pthread_barrier_init(&emu_block_barrier[b], nullptr, threads)) return -1;
    if (pthread_barrier_init(&emu_cluster_barrier, nullptr, blocks * threads)) return -1;
    const unsigned total = blocks * threads;
    std::vector<pthread_t> th(total);
    std::vector<emu_launch_ctx<Kernel, Args>> ctx(total);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (unsigned i = 0; i < total; ++i) {
        ctx[i] = {k, &args, i % threads, first_block + i / threads, block_y, i / threads};
        if (pthread_create(&th[i], &attr, emu_thread_main<Kernel, Args>, &ctx[i]))
            return -2;   // a partial start would dead-lock on the barriers; the caller treats it as fatal
    }
    for (unsigned i = 0; i < total; ++i) pthread_join(th[i], nullptr);
    pthread_attr_destroy(&attr);
    for (unsigned b = 0; b < blocks; ++b) pthread_barrier_destroy(&emu_block_barrier[b]);
    pthread_barrier_destroy(&emu_cluster_barrier);
    return 0;
}

// kernel<<<1, threads>>>(args)
template <class Kernel, class Args>
static int emu_launch_cta(Kernel k, const Args& args, unsigned threads) {
    gridDim = {1, 1, 1};
    return emu_run_blocks(k, args, 0, 1, threads);
}

// kernel<<<dim3(blocks, blocks_y), threads>>>(args) for kernels whose blocks do not communicate: one after another
template <class Kernel, class Args>
static int emu_launch_grid(Kernel k, const Args& args, unsigned blocks, unsigned threads, unsigned blocks_y = 1) {
    gridDim = {blocks, blocks_y, 1};
    for (unsigned y = 0; y < blocks_y; ++y)
        for (unsigned b = 0; b < blocks; ++b) {
            int rc = emu_run_blocks(k, args, b, 1, threads, y);
            if (rc) return rc;
        }
    gridDim = {1, 1, 1};
    return 0;
}

// kernel<<<cluster_size, threads>>>(args) launched as ONE cluster: the blocks run concurrently
template <class Kernel, class Args>
static int emu_launch_cluster(Kernel k, const Args& args, unsigned cluster_size, unsigned threads) {
    gridDim = {cluster_size, 1, 1};
    int rc = emu_run_blocks(k, args, 0, cluster_size, threads);
    gridDim = {1, 1, 1};
    return rc;
}
