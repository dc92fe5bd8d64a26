// scopa_b200/csrc/ms_sdcfr.cu -- SDCFR: batched advantage-net inference and the level-batched
// external-sampling traversal.  sm_100a only.
//
// Replaces (paths relative to /root/reference/):
//   DeepCFR._state_to_features / _get_legal_actions_mask   src/algorithms/deep_cfr/deep_cfr.py:213-282
//   AdvantageNetwork.get_advantages                         deep_cfr.py:54-68   (batch-1, one device sync per node)
//   positive_regret_policy                                  src/algorithms/deep_cfr/nets.py:93-101
//   DeepCFR._external_sampling_cfr                          deep_cfr.py:284-365
//   FlexibleNet (mlp mode 34 -> 128 -> 64 -> 16, ReLU)      nets.py:151-235, 296-331
//
// The reference does one batch-1 MLP call (and a host<->device hop) per tree node.  Here B traversals
// advance together, one tree level at a time: the frontier of a level (B x 1..24 nodes) is one batched
// inference (sd_level_mlp_kernel: state -> features -> MLP -> 16 raw advantages per node), followed by
// sd_expand_kernel (policy, then expansion of a traverser node / sampling at an opponent node, with the env step).
// The two are separate launches since round 2: the tensor-core kernel holds its TMEM columns and operand buffers
// only for the MLP, and the rule evaluation runs at full occupancy instead of at the 16 warps per SM the TMEM
// budget allows (profiles/README.md R2.5).  The recursion shape is data independent (a traverser node with h cards has exactly h
// children, an opponent node 1), so level arrays are dense: child index = parent index * fan-out + i,
// no compaction and no atomics; Philox call indices (= the reference's depth-first invocation order)
// follow from the same shape.
//
// Two inference paths, selected by `precision`:
//   0  fp32 on CUDA cores, separate multiply/add in the oracle's summation order (bit-comparable with
//      oracle/ms_oracle.c; this is the parity path);
//   1  bf16 operands / fp32 accumulation on the 5th-generation tensor cores (tcgen05.mma, accumulators
//      in TMEM): the throughput path.
#include <cuda_bf16.h>

#include "ms_common.cuh"
#include "ms_state.cuh"
#include "ms_div.cuh"

#ifndef MS_DYN_SMEM   // the host emulation (tests/emu) supplies its own: one buffer per emulated block
#define MS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace ms {

constexpr int SD_IN = 34, SD_H1 = 128, SD_H2 = 64, SD_OUT = 16;
constexpr int SD_W1 = 0, SD_B1 = SD_W1 + SD_H1 * SD_IN, SD_W2 = SD_B1 + SD_H1, SD_B2 = SD_W2 + SD_H2 * SD_H1,
              SD_W3 = SD_B2 + SD_H2, SD_B3 = SD_W3 + SD_OUT * SD_H2, SD_NW = SD_B3 + SD_OUT;   // 13776 floats
constexpr int SD_TILE = 128;   // nodes per CTA tile = threads per CTA (thread t owns node/row t)

struct SdShape {
    int n[9];            // nodes per traversal at each level
    int f[8];            // fan-out at each level (hand size at traverser levels, 1 at opponent levels)
    uint32_t size[9];    // _external_sampling_cfr invocations in the subtree of a level-d node
    int sample_off[8];   // first sample slot of a traverser level inside a traversal's block of samples
    int samples;         // samples per traversal (41 for a 4+4 deal)
    int player;
};

struct SdLevelPtrs { uint4* state; uint32_t* call; float* value; float4* pol; };

struct SdArgs {
    SdShape sh;
    SdLevelPtrs lvl[9];
    const float* net[2];          // fp32 parameter blobs (SD_NW floats each), per player
    const unsigned char* img[2];  // bf16 shared-memory images of the nets (tensor-core path)
    uint32_t hand_order;
    uint2 pkey;
    unsigned long long first_trav;
    long long n_trav;
    float* out_feat; float* out_target; float* out_mask; float* out_value;
    float* raw;                   // [n_trav * max nodes of a level][16] raw advantages of the level in flight
};

// features of the CURRENT player's view (deep_cfr.py:304): hand one-hot[16] by action id, table
// one-hot[16] (order dropped), [1.0 (player == current_player), 0.0]
__device__ __forceinline__ void sd_features(const MsState& s, int cp, float* x) {
    const uint32_t hand = st_hand(s, cp);
    const uint32_t tset = table_set(s.y, st_table_len(s));
#pragma unroll
    for (int i = 0; i < 16; i++) { x[i] = (float)((hand >> i) & 1u); x[16 + i] = (float)((tset >> i) & 1u); }
    x[32] = 1.f; x[33] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// fp32 CUDA-core MLP: weights in shared memory (every lane reads the same weight -> broadcast), the
// thread's hidden activations in shared memory as [k][thread] (conflict free).  acc = b; acc += w*x in
// index order with separate multiply and add (library is compiled with --fmad=false).
struct SdSmemFp32 { float* w; float* h1; float* h2; };

__device__ __forceinline__ void mlp_fp32(const SdSmemFp32& sm, const float* x, float* out) {
    const int tid = threadIdx.x;
    for (int o = 0; o < SD_H1; o++) {
        float acc = sm.w[SD_B1 + o];
        const float* wr = sm.w + SD_W1 + o * SD_IN;
#pragma unroll
        for (int i = 0; i < SD_IN; i++) acc += wr[i] * x[i];
        sm.h1[o * SD_TILE + tid] = acc > 0.f ? acc : 0.f;
    }
    for (int o = 0; o < SD_H2; o += 4) {
        float a0 = sm.w[SD_B2 + o], a1 = sm.w[SD_B2 + o + 1], a2 = sm.w[SD_B2 + o + 2], a3 = sm.w[SD_B2 + o + 3];
        const float* w0 = sm.w + SD_W2 + o * SD_H1;
        for (int k = 0; k < SD_H1; k++) {
            const float h = sm.h1[k * SD_TILE + tid];
            a0 += w0[k] * h; a1 += w0[SD_H1 + k] * h; a2 += w0[2 * SD_H1 + k] * h; a3 += w0[3 * SD_H1 + k] * h;
        }
        sm.h2[(o + 0) * SD_TILE + tid] = a0 > 0.f ? a0 : 0.f;
        sm.h2[(o + 1) * SD_TILE + tid] = a1 > 0.f ? a1 : 0.f;
        sm.h2[(o + 2) * SD_TILE + tid] = a2 > 0.f ? a2 : 0.f;
        sm.h2[(o + 3) * SD_TILE + tid] = a3 > 0.f ? a3 : 0.f;
    }
#pragma unroll
    for (int o = 0; o < SD_OUT; o++) {
        float acc = sm.w[SD_B3 + o];
        const float* wr = sm.w + SD_W3 + o * SD_H2;
        for (int k = 0; k < SD_H2; k++) acc += wr[k] * sm.h2[k * SD_TILE + tid];
        out[o] = acc;
    }
}

constexpr size_t SD_SMEM_FP32 = sizeof(float) * (SD_NW + 16 + SD_H1 * SD_TILE + SD_H2 * SD_TILE);

__device__ __forceinline__ SdSmemFp32 sd_carve_fp32(unsigned char* raw, const float* net) {
    SdSmemFp32 sm;
    sm.w = (float*)raw; sm.h1 = sm.w + SD_NW + 16; sm.h2 = sm.h1 + SD_H1 * SD_TILE;
    for (int i = threadIdx.x; i < SD_NW; i += blockDim.x) sm.w[i] = net[i];
    __syncthreads();
    return sm;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 path.  One CTA = 128 threads = one 128-row tile; thread t builds row t of the A operand
// (features, then the ReLU'd hidden activations) in shared memory in the UMMA K-major no-swizzle
// ("interleave") canonical layout -- 8x16-byte core matrices: element (m, k) of a K=16 slice lives at
//   (m / 8) * SBO + (k / 8) * LBO + (m % 8) * 16 + (k % 8) * 2  bytes --
// the weights W[out][in] (= nn.Linear's layout = K-major B operand) are staged once per CTA in the same
// layout as bf16, one elected thread issues tcgen05.mma (M=128, N=128/64/16, K=16 per instruction)
// into TMEM, completion is signalled through tcgen05.commit -> mbarrier, and thread t reads ITS
// accumulator row back with tcgen05.ld 32x32b for the bias + ReLU epilogue.
constexpr int SD_K1 = 48;                        // 34 padded to a multiple of 16
// TMEM: 128 columns per CTA.  The three accumulators reuse the same columns: an accumulator is fully
// read back (and a CTA barrier passed) before the next layer's MMA is issued.
constexpr uint32_t SD_TM_COLS = 128, SD_TM_D1 = 0, SD_TM_D2 = 0, SD_TM_D3 = 0;
// A CTA of the tensor-core path is TWO warpgroups, each working on its own 128-row tile with its own A buffer,
// mbarrier and 128 TMEM columns, sharing one copy of the weight image: 96 KB of shared memory per CTA, two CTAs
// per SM = four tiles in flight per SM (one tile per CTA allowed three: the kernel is bound by the per-tile
// dependency chain, so tiles in flight are what count).  The warpgroups never meet inside the tile loop: they
// synchronise on named barriers (id 1 + warpgroup, 128 threads).
constexpr int SD_TC_WG = 2, SD_TC_THREADS = SD_TC_WG * SD_TILE, SD_TC_CTAS_PER_SM = 2;
constexpr int SD_CTAS_PER_SM = 3;   // fp32 path launch bound (its shared memory allows one)
// bf16 operand image of one net, as it sits in shared memory (built once per call by sd_prep_kernel):
// w1 [128 x 48] | w2 [64 x 128] | w3 [16 x 64] in the UMMA canonical layout, then the BIAS OPERANDS: the biases are
// added by the tensor cores, not by the epilogues -- each layer's accumulator is started by one extra K = 16 MMA of a
// constant A tile `ones` [128 x 16] (columns 0 and 1 = 1.0, the rest 0) with a B tile [N x 16] whose columns 0 and 1
// hold the bias split into two bf16 terms (hi + lo: 16 mantissa bits), so D = b + A W^T leaves the epilogue of a hidden
// layer with a single packed convert (cvt.rn.relu.bf16x2.f32) per two activations.  Layer 1 needs no extra MMA: its
// K = 34 operand is padded to 48, and columns 34 / 35 of W1 hold the bias terms against 1.0 in the A rows.
constexpr int SD_IMG_W1 = 0, SD_IMG_W2 = SD_IMG_W1 + 2 * 128 * 48, SD_IMG_W3 = SD_IMG_W2 + 2 * 64 * 128,
              SD_IMG_ONES = SD_IMG_W3 + 2 * 16 * 64, SD_IMG_BT2 = SD_IMG_ONES + 2 * 128 * 16,
              SD_IMG_BT3 = SD_IMG_BT2 + 2 * 64 * 16, SD_IMG_BYTES = SD_IMG_BT3 + 2 * 16 * 16;

// switches of the tensor-core level kernel, measured on whole traversals (profiles/README.md R2.5, prof_r02q): bit 0 = request
// the next tile's state before the layers of the current one (no gain: four tiles in flight already hide the load), bit 1 =
// two x32 TMEM loads per wait in the hidden epilogues (+3.5 %).  Default: 2.
#ifndef SD_VAR_DEFAULT
#define SD_VAR_DEFAULT 2
#endif
struct SdSmemTc {
    __nv_bfloat16* a;      // A operand tile, 128 rows x up to 128 k   (32 KB)
    __nv_bfloat16* w1;     // 128 x 48
    __nv_bfloat16* w2;     // 64 x 128
    __nv_bfloat16* w3;     // 16 x 64
    __nv_bfloat16* ones;   // 128 x 16 constant A tile of the bias MMAs
    __nv_bfloat16* bt2;    // bias operands of layers 2 and 3: 64 x 16, 16 x 16
    __nv_bfloat16* bt3;
    unsigned long long* bar;
    uint32_t* tmem_base;
    uint32_t tm_off;       // this warpgroup's first TMEM column
    int wg;                // warpgroup index inside the CTA
};
constexpr size_t SD_SMEM_TC = SD_TC_WG * 2 * 128 * 128 + SD_IMG_BYTES + 64 + 1024;

__device__ __forceinline__ void sd_wg_sync(int wg) {
    asm volatile("bar.sync %0, 128;\n" :: "r"(1 + wg) : "memory");
}

// canonical K-major no-swizzle tile with `rows` rows and K elements: core matrix (8 rows x 8 k) = 128 B;
// K-adjacent core matrices are contiguous (LBO = 128 B), 8-row groups are K/8 core matrices apart
__device__ __forceinline__ uint32_t sd_tile_off(int row, int k, int K) {
    return (uint32_t)((row >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

__device__ __forceinline__ uint64_t sd_smem_desc(const void* p, int K) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((K >> 3) * 128) >> 4;
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46);   // version 1, no swizzle
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t sd_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void sd_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

__device__ __forceinline__ void sd_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void sd_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" :: "r"(a), "r"(parity) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns of the accumulator: thread = its lane (row)
__device__ __forceinline__ void sd_tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive columns in one instruction: half the load -> wait round trips of the hidden-layer epilogues
__device__ __forceinline__ void sd_tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// 64 consecutive columns: two x32 loads in flight, one wait
__device__ __forceinline__ void sd_tmem_ld64(uint32_t taddr, float* v) {
    uint32_t r[64];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint32_t* q = r + 32 * h;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
              "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]),
              "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]),
              "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
            : "r"(taddr + 32u * (uint32_t)h));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; i++) v[i] = __uint_as_float(r[i]);
}

// builds the shared-memory image of a net (bf16, canonical layout) in global memory, once per call
__global__ void __launch_bounds__(256) sd_prep_kernel(const float* __restrict__ net, unsigned char* __restrict__ img) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
    for (int i = tid; i < 128 * SD_K1; i += T) {
        const int o = i / SD_K1, k = i % SD_K1;
        // columns 34 and 35 of the padded K = 48 operand carry the layer's bias (hi + lo bf16 terms); the A rows hold 1.0 there
        float w = 0.f;
        if (k < SD_IN) w = net[SD_W1 + o * SD_IN + k];
        else if (k == SD_IN || k == SD_IN + 1) {
            const float b = net[SD_B1 + o];
            const float hi = __bfloat162float(__float2bfloat16(b));
            w = k == SD_IN ? hi : b - hi;
        }
        *(__nv_bfloat16*)(img + SD_IMG_W1 + sd_tile_off(o, k, SD_K1)) = __float2bfloat16(w);
    }
    for (int i = tid; i < 64 * 128; i += T) {
        const int o = i / 128, k = i % 128;
        *(__nv_bfloat16*)(img + SD_IMG_W2 + sd_tile_off(o, k, 128)) = __float2bfloat16(net[SD_W2 + o * 128 + k]);
    }
    for (int i = tid; i < 16 * 64; i += T) {
        const int o = i / 64, k = i % 64;
        *(__nv_bfloat16*)(img + SD_IMG_W3 + sd_tile_off(o, k, 64)) = __float2bfloat16(net[SD_W3 + o * 64 + k]);
    }
    for (int i = tid; i < 128 * 16; i += T) {
        const int m = i / 16, k = i % 16;
        *(__nv_bfloat16*)(img + SD_IMG_ONES + sd_tile_off(m, k, 16)) = __float2bfloat16(k < 2 ? 1.f : 0.f);
    }
    for (int i = tid; i < (64 + 16) * 16; i += T) {
        const int r = i / 16, k = i % 16;
        const int layer = r < 64 ? 1 : 2, o = layer == 1 ? r : r - 64;
        const float b = net[(layer == 1 ? SD_B2 : SD_B3) + o];
        const __nv_bfloat16 hi = __float2bfloat16(b);
        const __nv_bfloat16 lo = __float2bfloat16(b - __bfloat162float(hi));
        *(__nv_bfloat16*)(img + (layer == 1 ? SD_IMG_BT2 : SD_IMG_BT3) + sd_tile_off(o, k, 16)) =
            k == 0 ? hi : (k == 1 ? lo : __float2bfloat16(0.f));
    }
}

__device__ __forceinline__ SdSmemTc sd_carve_tc(unsigned char* raw, const unsigned char* img) {
    SdSmemTc sm;
    unsigned char* p = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    sm.wg = threadIdx.x >> 7;
    sm.tm_off = (uint32_t)sm.wg * SD_TM_COLS;
    sm.a = (__nv_bfloat16*)(p + sm.wg * 2 * 128 * 128); p += SD_TC_WG * 2 * 128 * 128;
    unsigned char* wimg = p;
    sm.w1 = (__nv_bfloat16*)(p + SD_IMG_W1);
    sm.w2 = (__nv_bfloat16*)(p + SD_IMG_W2);
    sm.w3 = (__nv_bfloat16*)(p + SD_IMG_W3);
    sm.ones = (__nv_bfloat16*)(p + SD_IMG_ONES);
    sm.bt2 = (__nv_bfloat16*)(p + SD_IMG_BT2); sm.bt3 = (__nv_bfloat16*)(p + SD_IMG_BT3);
    p += SD_IMG_BYTES;
    unsigned long long* bars = (unsigned long long*)p; p += 8 * SD_TC_WG;
    sm.bar = bars + sm.wg;
    sm.tmem_base = (uint32_t*)p;
    const int tid = threadIdx.x, T = blockDim.x;
    for (int i = tid; i < SD_IMG_BYTES / 16; i += T) ((uint4*)wimg)[i] = ((const uint4*)img)[i];
    if (tid == 0) {
        for (int w = 0; w < SD_TC_WG; w++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"((uint32_t)__cvta_generic_to_shared(bars + w)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < 32) {   // one warp allocates the TMEM columns (128 per warpgroup) and gives up the permit
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n"
                     :: "r"((uint32_t)__cvta_generic_to_shared(sm.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    return sm;
}

__device__ __forceinline__ void sd_release_tc(const SdSmemTc& sm) {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" :: "r"(*sm.tmem_base) : "memory");
}

// one layer: D[128 x N] = b + A[128 x K] * W[N x K]^T on the tensor cores (the bias through ones x bt, see the image
// layout; bt == nullptr: the bias is already inside A W^T -- layer 1, whose padded operand has two spare columns);
// all 128 threads of a warpgroup call it
__device__ __forceinline__ void sd_layer_mma(const SdSmemTc& sm, const __nv_bfloat16* w, const __nv_bfloat16* bt, int K, int N,
                                             uint32_t tm_col, uint32_t& phase) {
    // the A tile was written with ordinary stores: make it visible to the async (tensor core) proxy
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    sd_wg_sync(sm.wg);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if ((threadIdx.x & 127) == 0) {
        const uint32_t tm = *sm.tmem_base + sm.tm_off + tm_col;
        const uint32_t idesc = sd_idesc(N);
        if (bt) sd_mma(tm, sd_smem_desc(sm.ones, 16), sd_smem_desc(bt, 16), idesc, false);
        for (int k = 0; k < K; k += 16) {
            const uint64_t da = sd_smem_desc((const char*)sm.a + (k >> 3) * 128, K);
            const uint64_t db = sd_smem_desc((const char*)w + (k >> 3) * 128, K);
            sd_mma(tm, da, db, idesc, bt != nullptr || k > 0);
        }
        sd_commit(sm.bar);
    }
    sd_wait(sm.bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// 8 consecutive k-elements of one row are 16 contiguous bytes in the canonical layout: one 128-bit store
__device__ __forceinline__ void sd_store8(const SdSmemTc& sm, int row, int k0, int K, const float* v) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]),
                   p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 q;
    q.x = *(uint32_t*)&p0; q.y = *(uint32_t*)&p1; q.z = *(uint32_t*)&p2; q.w = *(uint32_t*)&p3;
    *(uint4*)((char*)sm.a + sd_tile_off(row, k0, K)) = q;
}

// ReLU + bf16 conversion + packing of two activations in ONE instruction (a -> upper half, b -> lower half)
__device__ __forceinline__ uint32_t sd_relu_pack(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// hidden-layer epilogue: `cols` accumulator columns of this thread's row -> ReLU -> bf16 -> next layer's A operand
template <int COLS, int W = 32>
__device__ __forceinline__ void sd_hidden_epilogue(const SdSmemTc& sm, int row, uint32_t taddr) {
#pragma unroll
    for (int c = 0; c < COLS; c += W) {
        float v[W];
        if (W == 64) sd_tmem_ld64(taddr + c, v); else sd_tmem_ld32(taddr + c, v);
#pragma unroll
        for (int i = 0; i < W; i += 8) {
            uint4 q;
            q.x = sd_relu_pack(v[i], v[i + 1]); q.y = sd_relu_pack(v[i + 2], v[i + 3]);
            q.z = sd_relu_pack(v[i + 4], v[i + 5]); q.w = sd_relu_pack(v[i + 6], v[i + 7]);
            *(uint4*)((char*)sm.a + sd_tile_off(row, c + i, COLS)) = q;
        }
    }
}

// layers 1-3 on the A1 tile already in sm.a (K = 48) -> out[16] = the net's raw outputs for this thread's row
template <int VAR = SD_VAR_DEFAULT>
__device__ __forceinline__ void mlp_tc_layers(const SdSmemTc& sm, float* out, uint32_t& phase) {
    const int tid = threadIdx.x & 127;                             // row of this warpgroup's tile
    const uint32_t lane_base = (((uint32_t)(tid & ~31)) << 16) + sm.tm_off;   // TMEM address: lane in bits 31..16
    constexpr int W = (VAR & 2) ? 64 : 32;
    sd_layer_mma(sm, sm.w1, nullptr, SD_K1, SD_H1, SD_TM_D1, phase);
    sd_hidden_epilogue<SD_H1, W>(sm, tid, *sm.tmem_base + lane_base + SD_TM_D1);
    sd_layer_mma(sm, sm.w2, sm.bt2, SD_H1, SD_H2, SD_TM_D2, phase);
    sd_hidden_epilogue<SD_H2, W>(sm, tid, *sm.tmem_base + lane_base + SD_TM_D2);
    sd_layer_mma(sm, sm.w3, sm.bt3, SD_H2, SD_OUT, SD_TM_D3, phase);
    sd_tmem_ld16(*sm.tmem_base + lane_base + SD_TM_D3, out);
}

// general features (ms_mlp_forward): row of A1 from 34 floats
__device__ __forceinline__ void mlp_tc(const SdSmemTc& sm, const float* x, float* out, uint32_t& phase) {
    const int tid = threadIdx.x & 127;
#pragma unroll
    for (int k0 = 0; k0 < SD_K1; k0 += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (k0 + i < SD_IN) ? x[k0 + i] : (k0 + i < SD_IN + 2 ? 1.f : 0.f);     // 34, 35: the bias columns
        sd_store8(sm, tid, k0, SD_K1, v);
    }
    mlp_tc_layers(sm, out, phase);
}

// Features of a game state straight from its bit masks (sd_features without the detour through 34 floats): element k of
// the row is bit k of the hand (k < 16) / of the table set (16 <= k < 32), k = 32 is the constant 1; a pair of bf16
// values is one 32-bit word, 1.0 = 0x3F80.
__device__ __forceinline__ uint32_t sd_bits_pair(uint32_t bits, int k) {
    const uint32_t b = bits >> k;
    return (b & 1u) * 0x3F80u | ((b >> 1) & 1u) * 0x3F800000u;
}
template <int VAR = SD_VAR_DEFAULT>
__device__ __forceinline__ void mlp_tc_state(const SdSmemTc& sm, const MsState& s, int cp, float* out, uint32_t& phase) {
    const int tid = threadIdx.x & 127;
    const uint32_t hand = st_hand(s, cp);
    const uint32_t tset = table_set(s.y, st_table_len(s));
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t bits = h == 0 ? hand : tset;
#pragma unroll
        for (int k0 = 0; k0 < 16; k0 += 8) {
            uint4 q;
            q.x = sd_bits_pair(bits, k0); q.y = sd_bits_pair(bits, k0 + 2); q.z = sd_bits_pair(bits, k0 + 4); q.w = sd_bits_pair(bits, k0 + 6);
            *(uint4*)((char*)sm.a + sd_tile_off(tid, 16 * h + k0, SD_K1)) = q;
        }
    }
    *(uint4*)((char*)sm.a + sd_tile_off(tid, 32, SD_K1)) = make_uint4(0x3F80u, 0x3F803F80u, 0u, 0u);   // [1.0, 0.0 | bias columns 1.0, 1.0 | 0 ..]
    *(uint4*)((char*)sm.a + sd_tile_off(tid, 40, SD_K1)) = make_uint4(0u, 0u, 0u, 0u);
    mlp_tc_layers<VAR>(sm, out, phase);
}

// ---------------------------------------------------------------------------------------------
// AdvantageNetwork.get_advantages masking (deep_cfr.py:66-67) + positive_regret_policy (nets.py:93-101)
__device__ __forceinline__ void sd_policy(const float* raw, uint32_t legal_mask, float* adv, float* pol) {
    float z = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float m = (float)((legal_mask >> i) & 1u);
        adv[i] = raw[i] * m - 1e6f * (1.f - m);
        const float pos = (adv[i] > 0.f ? adv[i] : 0.f) * m;
        pol[i] = pos;
        z += pos;
    }
    if (z < 1e-8f) z = 1e-8f;
    // 0 / z is exactly 0, but a zero numerator sends the IEEE division into its special-case subroutine: with 12+
    // of the 16 slots illegal that subroutine was 15 % of the forward kernel's instructions (profiles/README.md 3)
#pragma unroll
    for (int i = 0; i < 16; i++) pol[i] = ms_div_or_zero(pol[i], z);           // pol >= 0: exactly 0 stays 0
}

// The same policy where the traversal needs it: at the (at most four) legal actions, in legal-list order.  z is summed
// over all 16 slots in index order like sd_policy (the masked terms are exact zeros), but only the legal slots are
// divided: the IEEE division is a subroutine call per warp and slot -- with 16 slots and lanes holding different hands
// a warp ran it 16 times per node, 12 of them for numerators that are zero in every lane (30 % of sd_expand_kernel's
// instructions in profiles/sd_expand_r02m_raw.csv).  `rawp` = the node's 16 raw outputs in global memory (indexed reads).
__device__ __forceinline__ void sd_policy_legal(const float* raw, const float* __restrict__ rawp, uint32_t legal_mask,
                                                uint32_t list, uint32_t nl, float* p4) {
    // z = sum over the 16 slots, in index order, of max(adv, 0) * m with adv = raw * m - 1e6 * (1 - m): for a legal slot
    // (m = 1) that is max(raw, 0) exactly, for an illegal one an exact zero, and adding zeros does not change a float sum
    float z = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float pos = raw[i] > 0.f ? raw[i] : 0.f;
        z += ((legal_mask >> i) & 1u) ? pos : 0.f;
    }
    if (z < 1e-8f) z = 1e-8f;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        float v = 0.f;
        if (k < nl) {
            const float r = rawp[(list >> (4 * k)) & 0xFu];      // adv = r * 1 - 1e6 * 0 = r exactly
            v = ms_div_or_zero(r > 0.f ? r : 0.f, z);
        }
        p4[k] = v;
    }
}

template <int PREC>
__global__ void __launch_bounds__(PREC == 1 ? SD_TC_THREADS : SD_TILE, PREC == 1 ? SD_TC_CTAS_PER_SM : SD_CTAS_PER_SM) sd_mlp_kernel(const float* __restrict__ net, const unsigned char* __restrict__ img,
                                                            const float* __restrict__ feat,
                                                            const float* __restrict__ mask, float* __restrict__ adv_out,
                                                            float* __restrict__ pol_out, long long n) {
    MS_DYN_SMEM(smem_raw);
    SdSmemFp32 s32{};
    SdSmemTc stc{};
    uint32_t phase = 0;
    if (PREC == 0) s32 = sd_carve_fp32(smem_raw, net); else stc = sd_carve_tc(smem_raw, img);
    const int tid = threadIdx.x & (SD_TILE - 1);                 // row inside this warpgroup's tile
    constexpr int NWG = PREC == 1 ? SD_TC_WG : 1;                // tiles a CTA works on side by side
    const long long tiles = (n + SD_TILE - 1) / SD_TILE;
    for (long long tile = (long long)blockIdx.x * NWG + (threadIdx.x >> 7); tile < tiles; tile += (long long)gridDim.x * NWG) {
        const long long g = tile * SD_TILE + tid;
        float x[SD_IN], raw[16];
        uint32_t lm = 0u;
        for (int i = 0; i < SD_IN; i++) x[i] = g < n ? feat[g * SD_IN + i] : 0.f;
        for (int i = 0; i < 16; i++) if (g < n && mask[g * 16 + i] != 0.f) lm |= 1u << i;
        if (PREC == 0) mlp_fp32(s32, x, raw); else mlp_tc(stc, x, raw, phase);
        if (g < n) {
            float adv[16], pol[16];
            sd_policy(raw, lm, adv, pol);
            for (int i = 0; i < 16; i++) {
                if (adv_out) adv_out[g * 16 + i] = adv[i];
                if (pol_out) pol_out[g * 16 + i] = pol[i];
            }
        }
    }
    if (PREC == 1) sd_release_tc(stc);
}

// ---------------------------------------------------------------------------------------------
// Forward level d, part 1: inference for every frontier node -> a.raw[g][16] (the net's outputs before masking).
template <int PREC, int VAR = SD_VAR_DEFAULT>
__global__ void __launch_bounds__(PREC == 1 ? SD_TC_THREADS : SD_TILE, PREC == 1 ? SD_TC_CTAS_PER_SM : SD_CTAS_PER_SM) sd_level_mlp_kernel(SdArgs a, int d) {
    MS_DYN_SMEM(smem_raw);
    const int cp = d & 1;
    SdSmemFp32 s32{};
    SdSmemTc stc{};
    uint32_t phase = 0;
    if (PREC == 0) s32 = sd_carve_fp32(smem_raw, a.net[cp]); else stc = sd_carve_tc(smem_raw, a.img[cp]);
    const int tid = threadIdx.x & (SD_TILE - 1);                 // row inside this warpgroup's tile
    constexpr int NWG = PREC == 1 ? SD_TC_WG : 1;                // tiles a CTA works on side by side
    const long long total = a.n_trav * a.sh.n[d];
    const long long tiles = (total + SD_TILE - 1) / SD_TILE;
    const long long tstep = (long long)gridDim.x * NWG;
    long long tile = (long long)blockIdx.x * NWG + (threadIdx.x >> 7);
    // the state of the NEXT tile is requested before the layers of the current one: the tile loop never waits on HBM
    MsState s_next = (tile < tiles && tile * SD_TILE + tid < total) ? a.lvl[d].state[tile * SD_TILE + tid] : make_uint4(0u, 0u, 0u, 0u);
    for (; tile < tiles; tile += tstep) {
        const long long g = tile * SD_TILE + tid;
        const bool live = g < total;
        MsState s = s_next;
        if (PREC == 1 && (VAR & 1)) {
            const long long gn = g + tstep * SD_TILE;
            s_next = (tile + tstep < tiles && gn < total) ? a.lvl[d].state[gn] : make_uint4(0u, 0u, 0u, 0u);
        } else if (tile > (long long)blockIdx.x * NWG + (threadIdx.x >> 7)) {
            s = live ? a.lvl[d].state[g] : make_uint4(0u, 0u, 0u, 0u);
        }
        float raw[16];
        if (PREC == 0) {
            float x[SD_IN];
            sd_features(s, cp, x);
            mlp_fp32(s32, x, raw);
        } else {
            mlp_tc_state<VAR>(stc, s, cp, raw, phase);
        }
        if (live) {
            float4* o = (float4*)(a.raw + g * 16);
#pragma unroll
            for (int i = 0; i < 4; i++) o[i] = make_float4(raw[4 * i], raw[4 * i + 1], raw[4 * i + 2], raw[4 * i + 3]);
        }
    }
    if (PREC == 1) sd_release_tc(stc);
}

// Forward level d, part 2: masking + regret-matching policy, then expand (traverser: every legal action) or sample
// (opponent: one action from the policy), with the env step.  Plain CUDA at full occupancy.
__global__ void __launch_bounds__(256) sd_expand_kernel(SdArgs a, int d) {
    const int cp = d & 1;
    const bool trav = (cp == a.sh.player);
    const long long total = a.n_trav * a.sh.n[d];
    const int f = a.sh.f[d];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const MsState s = a.lvl[d].state[g];
        float raw[16];
        {
            const float4* r4 = (const float4*)(a.raw + g * 16);
#pragma unroll
            for (int i = 0; i < 4; i++) { const float4 v = r4[i]; raw[4 * i] = v.x; raw[4 * i + 1] = v.y; raw[4 * i + 2] = v.z; raw[4 * i + 3] = v.w; }
        }
        uint32_t list;
        const uint32_t nl = legal_list(s, a.hand_order, cp, list);
        uint32_t lm = 0u;
        for (uint32_t i = 0; i < nl; i++) lm |= 1u << ((list >> (4 * i)) & 0xFu);
        float pl[4];                                             // policy at the legal actions, legal-list order
        sd_policy_legal(raw, a.raw + g * 16, lm, list, nl, pl);
        const uint32_t call = a.lvl[d].call[g];
        if (trav) {
            for (int i = 0; i < f; i++) {
                const uint32_t act = (list >> (4 * i)) & 0xFu;
                MsState c = s;
                step(c, act);
                a.lvl[d + 1].state[g * f + i] = c;
                a.lvl[d + 1].call[g * f + i] = call + 1u + (uint32_t)i * a.sh.size[d + 1];
            }
            a.lvl[d].pol[g] = make_float4(pl[0], pl[1], pl[2], pl[3]);
        } else {
            // opponent: sample one action from the policy restricted to the legal list (deep_cfr.py:347-359)
            float ap[4], sum = 0.f;
            for (uint32_t i = 0; i < 4; i++) ap[i] = pl[i];
            for (uint32_t i = 0; i < nl; i++) sum += ap[i];
            const long long t = g / a.sh.n[d];
            const unsigned long long trav_id = a.first_trav + (unsigned long long)t;
            const uint4 xb = philox4x32_10(make_uint4((uint32_t)trav_id, (uint32_t)(trav_id >> 32), call >> 1,
                                                      MS_TAG_SDCF + (uint32_t)a.sh.player), a.pkey);
            const uint32_t w0 = (call & 1u) ? xb.z : xb.x, w1 = (call & 1u) ? xb.w : xb.y;
            uint32_t ai;
            if (sum == 0.f) ai = __umulhi(w0, nl);                    // np.random.choice(legal_actions)
            else {
                double cdf[4], acc = 0.0;
                for (uint32_t i = 0; i < 4; i++) { if (i < nl) acc = __dadd_rn(acc, (double)ms_div_or_zero(ap[i] > 0.f ? ap[i] : 0.f, sum)); cdf[i] = acc; }
                const double last = acc, u = u53(w0, w1);
                ai = 0u;
                for (uint32_t i = 0; i < nl; i++) if (ms_ddiv_or_zero(cdf[i], last) <= u) ai++;
                if (ai >= nl) ai = nl - 1u;
            }
            MsState c = s;
            step(c, (list >> (4 * ai)) & 0xFu);
            a.lvl[d + 1].state[g] = c;
            a.lvl[d + 1].call[g] = call + 1u;
        }
    }
}

// Opponent level where the mover holds a single card: the move is forced, so the advantage net's
// output cannot influence anything and no inference is run (the reference still calls the net there).
__global__ void __launch_bounds__(256) sd_forced_kernel(SdArgs a, int d) {
    const long long total = a.n_trav * a.sh.n[d];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        MsState s = a.lvl[d].state[g];
        uint32_t list;
        legal_list(s, a.hand_order, d & 1, list);
        step(s, list & 0xFu);
        a.lvl[d + 1].state[g] = s;
        a.lvl[d + 1].call[g] = a.lvl[d].call[g] + 1u;
    }
}

__global__ void __launch_bounds__(256) sd_init_kernel(SdArgs a, uint4 root) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < a.n_trav; t += (long long)gridDim.x * blockDim.x) {
        a.lvl[0].state[t] = root;
        a.lvl[0].call[t] = 0u;
    }
}

__global__ void __launch_bounds__(256) sd_terminal_kernel(SdArgs a) {
    const long long total = a.n_trav * a.sh.n[8];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const MsState s = a.lvl[8].state[g];
        const int r = st_terminal(s) ? reward0_x2(s) : 0;
        a.lvl[8].value[g] = 0.5f * (float)(a.sh.player == 0 ? r : -r);
    }
}

// Backward level d: traverser nodes combine their children's values, form the regret target and emit a
// sample (deep_cfr.py:321-346, :70-75); opponent nodes pass the sampled child's value up.
// A sample is 66 floats (34 features, 16 targets, 16 mask entries) in three row-major arrays.  Written by the
// thread that owns the node, every store instruction of a warp touched 32 different rows (32 partial sectors):
// the launch list showed these kernels at 0.6 TB/s and 44 % of a traversal.  The rows of a warp are staged in
// shared memory instead and written out by the whole warp, consecutive lanes on consecutive floats.
__global__ void __launch_bounds__(256) sd_backward_kernel(SdArgs a, int d) {
    // per warp: 32 feature rows (34 floats, row stride 34: the copy-out below reads them linearly), then the 32 target
    // and mask rows (16 + 16 floats) in the same buffer
    __shared__ float4 stage[8][32 * 34 / 4];
    __shared__ uint32_t rel_s[8][32];
    const int cp = d & 1;
    const bool trav = (cp == a.sh.player);
    const long long total = a.n_trav * a.sh.n[d];
    const int f = a.sh.f[d], nd = a.sh.n[d];
    if (!trav) {
        for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x)
            a.lvl[d].value[g] = a.lvl[d + 1].value[g];
        return;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* st = (float*)stage[wib];
    uint32_t* rel = rel_s[wib];
    for (long long base = blockIdx.x * (long long)blockDim.x + 32 * wib; base < total; base += (long long)gridDim.x * blockDim.x) {
        const long long g = base + lane;
        const bool valid = g < total;
        const int nrows = (int)((total - base) < 32 ? (total - base) : 32);
        // sample slots of the warp's rows: ascending, so a row's slot is the first row's plus a small 32-bit offset
        const long long t0 = base / nd;
        const long long slot0 = t0 * a.sh.samples + a.sh.sample_off[d] + (base - t0 * nd);
        float x[SD_IN], reg[16];
        uint32_t lm = 0u;
        if (valid) {
            const MsState s = a.lvl[d].state[g];
            uint32_t list;
            legal_list(s, a.hand_order, cp, list);
            const float4 p4 = a.lvl[d].pol[g];
            const float pl[4] = {p4.x, p4.y, p4.z, p4.w};
            float value = 0.f, cfv[16];
#pragma unroll
            for (int i = 0; i < 16; i++) cfv[i] = 0.f;
            for (int i = 0; i < f; i++) {
                const uint32_t act = (list >> (4 * i)) & 0xFu;
                const float av = a.lvl[d + 1].value[g * f + i];
                value += pl[i] * av;
                cfv[act] = av;
                lm |= 1u << act;
            }
            a.lvl[d].value[g] = value;
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < 16; i++) { reg[i] = cfv[i] - value; mx = fmaxf(mx, fabsf(reg[i])); }
            if (mx > 0.f) {
                const float dn = mx + 1e-8f;
#pragma unroll
                for (int i = 0; i < 16; i++) reg[i] = ms_div_or_zero(reg[i], dn);
            }
            const long long t = g / nd, j = g % nd;
            rel[lane] = (uint32_t)(t * a.sh.samples + a.sh.sample_off[d] + j - slot0);
            sd_features(s, cp, x);
        }
        __syncwarp();                       // the previous iteration's readers are done with the stage
        if (valid) {
#pragma unroll
            for (int i = 0; i < SD_IN; i++) st[lane * SD_IN + i] = x[i];
        }
        __syncwarp();
        {   // 34 floats per row = 17 float2 (a row starts 136 bytes after the previous one: 8-byte aligned)
            float2* fb = (float2*)(a.out_feat + slot0 * SD_IN);
            const float2* s2 = (const float2*)st;
            for (int e = lane; e < nrows * (SD_IN / 2); e += 32) {
                const int r = e / (SD_IN / 2), c = e - r * (SD_IN / 2);
                fb[rel[r] * (uint32_t)(SD_IN / 2) + (uint32_t)c] = s2[e];
            }
        }
        __syncwarp();
        if (valid) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                st[lane * 16 + i] = reg[i];
                st[512 + lane * 16 + i] = (float)((lm >> i) & 1u);
            }
        }
        __syncwarp();
        {   // 16 floats per row = 4 float4, for the targets and for the masks
            float4* tb = (float4*)(a.out_target + slot0 * 16);
            float4* mb = (float4*)(a.out_mask + slot0 * 16);
            const float4* s4 = (const float4*)st;
            for (int e = lane; e < nrows * 4; e += 32) {
                const uint32_t o = rel[e >> 2] * 4u + (uint32_t)(e & 3);
                tb[o] = s4[e];
                mb[o] = s4[128 + e];
            }
        }
    }
}

__global__ void __launch_bounds__(256) sd_root_value_kernel(SdArgs a) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < a.n_trav; t += (long long)gridDim.x * blockDim.x)
        a.out_value[t] = a.lvl[0].value[t];
}

static void sd_shape(int player, SdShape& sh) {
    sh.player = player;
    sh.n[0] = 1;
    for (int d = 0; d < 8; d++) {
        sh.f[d] = ((d & 1) == player) ? (4 - d / 2) : 1;
        sh.n[d + 1] = sh.n[d] * sh.f[d];
    }
    sh.size[8] = 1;
    for (int d = 7; d >= 0; d--) sh.size[d] = 1u + (uint32_t)sh.f[d] * sh.size[d + 1];
    int off = 0;
    for (int d = 0; d < 8; d++) {
        sh.sample_off[d] = off;
        if ((d & 1) == player) off += sh.n[d];
    }
    sh.samples = off;
}

static size_t sd_workspace(long long n_trav, int player, SdArgs* a, char* base) {
    SdShape sh;
    sd_shape(player, sh);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    for (int d = 0; d <= 8; d++) {
        const size_t nn = (size_t)n_trav * sh.n[d];
        size_t o_state = take(16 * nn), o_call = take(4 * nn), o_val = take(4 * nn), o_pol = take(16 * nn);
        if (a) {
            a->lvl[d].state = (uint4*)(base + o_state); a->lvl[d].call = (uint32_t*)(base + o_call);
            a->lvl[d].value = (float*)(base + o_val); a->lvl[d].pol = (float4*)(base + o_pol);
        }
    }
    for (int p = 0; p < 2; p++) {
        size_t o_img = take(SD_IMG_BYTES);
        if (a) a->img[p] = (const unsigned char*)(base + o_img);
    }
    int widest = 1;
    for (int d = 0; d < 8; d++) widest = sh.n[d] > widest ? sh.n[d] : widest;
    const size_t o_raw = take(sizeof(float) * 16 * (size_t)n_trav * widest);
    if (a) a->raw = (float*)(base + o_raw);
    if (a) a->sh = sh;
    return off;
}

}  // namespace ms

#ifndef MS_HOST_RULES_ONLY   // tests/emu/ms_sdcfr_host.cpp compiles the kernels above for the host's CTA emulator (fp32 path);
                             // below: the library's host side (CUDA runtime calls, launches, C ABI)
using namespace ms;

extern "C" {

int ms_sdcfr_infer_states(const ms_state* d_states, int64_t n, int player_to_move, const float* d_net, int precision,
                          float* d_raw, void* stream) {
    if (n < 0 || player_to_move < 0 || player_to_move > 1 || precision < 0 || precision > 1 ||
        (n > 0 && (!d_states || !d_net || !d_raw)))
        return fail(MS_ERR_ARG, "ms_sdcfr_infer_states: bad argument");
    if (n == 0) return MS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    SdArgs a{};
    const int d = player_to_move;                          // level parity = player to move
    a.sh.n[d] = 1; a.n_trav = n;
    a.lvl[d].state = (uint4*)d_states;
    a.net[d] = d_net; a.raw = d_raw;
    if (precision == 0) {
        MS_CUDA(cudaFuncSetAttribute(sd_level_mlp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM_FP32));
        sd_level_mlp_kernel<0><<<grid_for(n, SD_TILE, 1), SD_TILE, SD_SMEM_FP32, st>>>(a, d);
        MS_LAUNCH_CHECK();
        return MS_OK;
    }
    unsigned char* img = nullptr;
    MS_CUDA(cudaMallocAsync((void**)&img, SD_IMG_BYTES, st));
    sd_prep_kernel<<<8, 256, 0, st>>>(d_net, img);
    MS_LAUNCH_CHECK();
    a.img[d] = img;
    MS_CUDA(cudaFuncSetAttribute(sd_level_mlp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM_TC));
    sd_level_mlp_kernel<1><<<grid_for(n, SD_TC_THREADS, SD_TC_CTAS_PER_SM), SD_TC_THREADS, SD_SMEM_TC, st>>>(a, d);
    MS_LAUNCH_CHECK();
    MS_CUDA(cudaFreeAsync(img, st));
    return MS_OK;
}

int ms_sdcfr_samples_per_traversal(int player) {
    SdShape sh;
    sd_shape(player & 1, sh);
    return sh.samples;
}

size_t ms_sdcfr_workspace_bytes(int64_t n_trav) {
    size_t a = sd_workspace(n_trav, 0, nullptr, nullptr), b = sd_workspace(n_trav, 1, nullptr, nullptr);
    return a > b ? a : b;
}

int ms_mlp_forward(const float* d_net, int precision, const float* d_feat, const float* d_mask, float* d_adv,
                   float* d_pol, int64_t n, void* stream) {
    if (n < 0 || precision < 0 || precision > 1 || (n > 0 && (!d_net || !d_feat || !d_mask)))
        return fail(MS_ERR_ARG, "ms_mlp_forward: bad argument");
    if (n == 0) return MS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == 0) {
        MS_CUDA(cudaFuncSetAttribute(sd_mlp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM_FP32));
        sd_mlp_kernel<0><<<grid_for(n, SD_TILE, 1), SD_TILE, SD_SMEM_FP32, st>>>(d_net, nullptr, d_feat, d_mask, d_adv, d_pol, (long long)n);
        MS_LAUNCH_CHECK();
    } else {
        unsigned char* img = nullptr;
        MS_CUDA(cudaMallocAsync((void**)&img, SD_IMG_BYTES, st));      // stream-ordered scratch for the operand image
        sd_prep_kernel<<<8, 256, 0, st>>>(d_net, img);
        MS_LAUNCH_CHECK();
        MS_CUDA(cudaFuncSetAttribute(sd_mlp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM_TC));
        sd_mlp_kernel<1><<<grid_for(n, SD_TC_THREADS, SD_TC_CTAS_PER_SM), SD_TC_THREADS, SD_SMEM_TC, st>>>(d_net, img, d_feat, d_mask, d_adv, d_pol, (long long)n);
        MS_LAUNCH_CHECK();
        MS_CUDA(cudaFreeAsync(img, st));
    }
    return MS_OK;
}

int ms_sdcfr_traverse(const ms_state* h_root, uint32_t hand_order, int player, const float* d_net0, const float* d_net1,
                      int precision, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav, void* d_workspace,
                      size_t workspace_bytes, float* d_feat, float* d_target, float* d_mask, float* d_root_value,
                      void* stream) {
    if (!h_root || player < 0 || player > 1 || precision < 0 || precision > 1 || n_trav < 0 || !d_net0 || !d_net1)
        return fail(MS_ERR_ARG, "ms_sdcfr_traverse: bad argument");
    if (n_trav == 0) return MS_OK;
    if (!d_workspace || !d_feat || !d_target || !d_mask) return fail(MS_ERR_ARG, "ms_sdcfr_traverse: null buffer");
    // the level shapes assume the reference's root: both hands full, player 0 to move, nothing played yet
    const uint32_t h0 = h_root->hands & 0xFFFFu, h1 = h_root->hands >> 16;
    if (__builtin_popcount(h0) != 4 || __builtin_popcount(h1) != 4 || ((h_root->meta >> 17) & 1u) || ((h_root->meta >> 12) & 0x1Fu))
        return fail(MS_ERR_ARG, "ms_sdcfr_traverse: root must be a fresh 4+4-card deal with player 0 to move");
    SdArgs a{};
    const size_t need = sd_workspace(n_trav, player, &a, (char*)d_workspace);
    if (need > workspace_bytes) return fail(MS_ERR_CAPACITY, "workspace too small: need %zu bytes", need);
    a.net[0] = d_net0; a.net[1] = d_net1;
    a.hand_order = hand_order;
    a.pkey = make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32));
    a.first_trav = first_trav; a.n_trav = n_trav;
    a.out_feat = d_feat; a.out_target = d_target; a.out_mask = d_mask; a.out_value = d_root_value;
    cudaStream_t st = (cudaStream_t)stream;
    const uint4 root = make_uint4(h_root->hands, h_root->table, h_root->captures, h_root->meta);
    sd_init_kernel<<<grid_for(n_trav, 256, 4), 256, 0, st>>>(a, root);
    MS_LAUNCH_CHECK();
    if (precision == 0) MS_CUDA(cudaFuncSetAttribute(sd_level_mlp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM_FP32));
    else {
        MS_CUDA(cudaFuncSetAttribute(sd_level_mlp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM_TC));
        for (int p = 0; p < 2; p++) {
            sd_prep_kernel<<<8, 256, 0, st>>>(a.net[p], (unsigned char*)a.img[p]);
            MS_LAUNCH_CHECK();
        }
    }
    for (int d = 0; d < 8; d++) {
        const bool forced_opp = ((d & 1) != player) && (4 - d / 2 == 1);   // the opponent's last card
        if (forced_opp) {
            sd_forced_kernel<<<grid_for(n_trav * a.sh.n[d], 256, 8), 256, 0, st>>>(a, d);
            MS_LAUNCH_CHECK();
            continue;
        }
        if (precision == 0) {
            sd_level_mlp_kernel<0><<<grid_for(n_trav * a.sh.n[d], SD_TILE, 1), SD_TILE, SD_SMEM_FP32, st>>>(a, d);
        } else {
            sd_level_mlp_kernel<1><<<grid_for(n_trav * a.sh.n[d], SD_TC_THREADS, SD_TC_CTAS_PER_SM), SD_TC_THREADS, SD_SMEM_TC, st>>>(a, d);
        }
        MS_LAUNCH_CHECK();
        sd_expand_kernel<<<grid_for(n_trav * a.sh.n[d], 256, 8), 256, 0, st>>>(a, d);
        MS_LAUNCH_CHECK();
    }
    sd_terminal_kernel<<<grid_for(n_trav * a.sh.n[8], 256, 8), 256, 0, st>>>(a);
    MS_LAUNCH_CHECK();
    for (int d = 7; d >= 0; d--) {
        sd_backward_kernel<<<grid_for(n_trav * a.sh.n[d], 256, 8), 256, 0, st>>>(a, d);
        MS_LAUNCH_CHECK();
    }
    if (d_root_value) {
        sd_root_value_kernel<<<grid_for(n_trav, 256, 4), 256, 0, st>>>(a);
        MS_LAUNCH_CHECK();
    }
    return MS_OK;
}

}  // extern "C"
#endif  // MS_HOST_RULES_ONLY
