// scopa_b200/csrc/ms_sd_train_cluster.cuh -- sd_train_cluster_kernel: the optimiser of ms_sd_train.cuh spread over a
// thread-block CLUSTER of 8 CTAs (8 SMs) that exchange through distributed shared memory.
//
// Work split.  CTA c of the cluster owns minibatch rows 16c .. 16c+15 for the forward and backward passes and
// parameter slice c (1722 of the 13 776 floats) for the optimiser:
//   1. every CTA keeps a full copy of the (padded) parameters in its own shared memory, gathers its 16 rows, runs
//      forward + backward on them and leaves its PARTIAL gradient (all 13 776 entries, summed over its rows) in its
//      shared memory; it also publishes its share of the loss and a bad-row flag;                     cluster barrier 1
//   2. CTA c adds slice c of the eight partial gradients (remote shared-memory reads, rank order), keeps the sum and
//      publishes the slice's sum of squares;                                                          cluster barrier 2
//   3. every CTA adds the eight sums of squares (rank order) -> global norm -> clip factor; CTA c applies Adam to its
//      slice (moments of the slice live in its shared memory for the whole launch) and stores the new parameter
//      values into the parameter images of ALL eight CTAs (remote shared-memory writes);              cluster barrier 3
// Three cluster barriers per optimiser step, no global-memory traffic between steps.  Arithmetic is fp32 fmaf like the
// one-CTA kernel, but gradients are summed per CTA first and then across CTAs, so the two kernels agree to rounding, not
// bit for bit; each is bit-identical to its own host emulation (tests/emu), which is what pins them to torch.
//
// Shared memory per CTA (floats): parameters 13 984, partial gradient 13 776, x 16x35, h1 16x129, h2 16x65, d 16x17,
// Adam slice 2 x 1722, reduction scratch 256 + 32, exchange words 2 x 4  ->  35 452 floats = 141 808 B.
//
// No static __shared__ variables (see tests/emu/cta_emu.h); the only cluster-specific constructs are
// ms_cluster_rank / ms_cluster_sync / ms_cluster_map.
#pragma once
#include "ms_sd_train.cuh"

#ifndef MS_CTA_EMU
#include <cooperative_groups.h>
__device__ __forceinline__ unsigned ms_cluster_rank() { return cooperative_groups::this_cluster().block_rank(); }
__device__ __forceinline__ void ms_cluster_sync() { cooperative_groups::this_cluster().sync(); }
template <class T>
__device__ __forceinline__ T* ms_cluster_map(T* p, unsigned rank) {
    return cooperative_groups::this_cluster().map_shared_rank(p, rank);
}
#endif

namespace ms {
namespace sdc {
using namespace sdt;
constexpr int kCluster = 8, kCThreads = 256, kRowsPer = kMaxBatch / kCluster;          // 16 rows per CTA
constexpr int kSlice = (kNetFloats + kCluster - 1) / kCluster;                        // 1722 parameters per CTA
// shared-memory offsets (floats); the parameter image sits at 0 like in the one-CTA kernel (smem_of)
constexpr int CG = SB3 + kOut, CX = CG + kNetFloats, CH1 = CX + kRowsPer * LDX, CH2 = CH1 + kRowsPer * LD1,
              CD = CH2 + kRowsPer * LD2, CM = CD + kRowsPer * LDO, CV = CM + kSlice, CRED = CV + kSlice,
              CXCH = CRED + kCThreads + 32, kCSmemFloats = CXCH + 8;
constexpr int kCSmemBytes = kCSmemFloats * 4;
static_assert(kCSmemBytes <= 227 * 1024, "shared memory per CTA");
// exchange words, double-buffered by epoch parity: [parity][0] bad-row flag, [1] loss share, [2] slice sum of squares
}  // namespace sdc

__global__ void __cluster_dims__(sdc::kCluster, 1, 1) __launch_bounds__(sdc::kCThreads, 1) sd_train_cluster_kernel(SdTrainArgs a) {
    using namespace sdc;
    MS_DYN_SMEM(sd_train_cluster_smem);
    float* S = reinterpret_cast<float*>(sd_train_cluster_smem);
    const int tid = (int)threadIdx.x, T = (int)blockDim.x;
    const int rank = (int)ms_cluster_rank();
    const int B = a.batch;
    const int row0 = rank * kRowsPer;                                   // first minibatch row of this CTA
    const int R = B - row0 < 0 ? 0 : (B - row0 < kRowsPer ? B - row0 : kRowsPer);      // rows it really has
    const int Rp = R == 0 ? 4 : ((R + 3) & ~3);                         // rows processed (zero rows carry no gradient)
    const float inv_n = 1.0f / (float)(B * kOut);
    const float w1 = (float)(1.0 - a.beta1), fb2 = (float)a.beta2, w2 = (float)(1.0 - a.beta2);
    const float feps = (float)a.eps, fmax_norm = (float)a.max_norm;
    double b1pow = a.b1pow, b2pow = a.b2pow;
    const int e0 = rank * kSlice, e1 = e0 + kSlice < kNetFloats ? e0 + kSlice : kNetFloats;   // this CTA's parameter slice

    for (int e = tid; e < kNetFloats; e += T) S[smem_of(e)] = a.net[e];
    for (int e = e0 + tid; e < e1; e += T) {
        S[CM + e - e0] = a.adam_m[e];
        S[CV + e - e0] = a.adam_v[e];
    }
    ms_cluster_sync();                                                  // every CTA of the cluster is resident

    for (int ep = 0; ep < a.epochs; ++ep) {
        const int* idx = a.idx + (long long)ep * B + row0;
        float* xch = S + CXCH + 4 * (ep & 1);
        if (tid == 0) xch[0] = 0.f;
        __syncthreads();
        // ---- 1. gather, forward, backward on this CTA's rows
        for (int t = tid; t < Rp * kIn; t += T) {
            int m = t / kIn, k = t % kIn;
            float v = 0.f;
            if (m < R) {
                long long row = idx[m];
                if (row < 0 || row >= a.n_rows) {
                    xch[0] = 1.f;
                    row = 0;
                }
                v = a.feat[row * kIn + k];
            }
            S[CX + m * LDX + k] = v;
        }
        __syncthreads();
        cta_gemm<2, 4>(Rp, kH1, kIn, S + CX, LDX, 1, S + SW1, LDX, 1,
                       [&](int i, int j, float v) { S[CH1 + i * LD1 + j] = fmaxf(v + S[SB1 + j], 0.f); });
        __syncthreads();
        cta_gemm<2, 2>(Rp, kH2, kH1, S + CH1, LD1, 1, S + SW2, LD1, 1,
                       [&](int i, int j, float v) { S[CH2 + i * LD2 + j] = fmaxf(v + S[SB2 + j], 0.f); });
        __syncthreads();
        float sq = 0.f;
        cta_gemm<1, 1>(Rp, kOut, kH2, S + CH2, LD2, 1, S + SW3, LD2, 1, [&](int i, int j, float v) {
            float d = 0.f;
            if (i < R) {
                long long row = idx[i];
                if (row < 0 || row >= a.n_rows) row = 0;                // flagged above; the step will be skipped
                float mk = a.mask[row * kOut + j], tg = a.target[row * kOut + j];
                float diff = (v + S[SB3 + j]) * mk - tg * mk;
                sq += diff * diff;
                d = (2.0f * diff * inv_n) * mk;
            }
            S[CD + i * LDO + j] = d;
        });
        const float loss_share = cta_sum<kCThreads>(sq, S + CRED);      // barriers inside: d is complete afterwards
        if (tid == 0) xch[1] = loss_share;
        float* G = S + CG;
        cta_gemm<2, 2>(kOut, kH2, Rp, S + CD, 1, LDO, S + CH2, 1, LD2,
                       [&](int i, int j, float v) { G[OW3 + i * kH2 + j] = v; });
        col_sums(S + CD, LDO, Rp, kOut, G + OB3);
        __syncthreads();                                                // h2 is about to be overwritten
        cta_gemm<2, 2>(Rp, kH2, kOut, S + CD, LDO, 1, S + SW3, 1, LD2, [&](int i, int j, float v) {
            float& h = S[CH2 + i * LD2 + j];
            h = h > 0.f ? v : 0.f;
        });
        __syncthreads();
        cta_gemm<4, 4>(kH2, kH1, Rp, S + CH2, 1, LD2, S + CH1, 1, LD1,
                       [&](int i, int j, float v) { G[OW2 + i * kH1 + j] = v; });
        col_sums(S + CH2, LD2, Rp, kH2, G + OB2);
        __syncthreads();                                                // h1 is about to be overwritten
        cta_gemm<2, 4>(Rp, kH1, kH2, S + CH2, LD2, 1, S + SW2, 1, LD1, [&](int i, int j, float v) {
            float& h = S[CH1 + i * LD1 + j];
            h = h > 0.f ? v : 0.f;
        });
        __syncthreads();
        cta_gemm<4, 4>(kH1, kIn, Rp, S + CH1, 1, LD1, S + CX, 1, LDX,
                       [&](int i, int j, float v) { G[OW1 + i * kIn + j] = v; });
        col_sums(S + CH1, LD1, Rp, kH1, G + OB1);
        ms_cluster_sync();                                              // barrier 1: partial gradients, flags, loss shares
        // ---- 2. this CTA's slice of the summed gradient
        float bad = 0.f, loss = 0.f;
        for (int r = 0; r < kCluster; ++r) {
            const float* x = ms_cluster_map(xch, (unsigned)r);
            bad += x[0];
            loss += x[1];
        }
        if (bad != 0.f) {                                               // the same verdict in every CTA of the cluster
            if (rank == 0 && tid == 0) a.loss[ep] = __int_as_float(0x7fc00000);
            continue;                                                   // (exchange words of the next epoch: other parity)
        }
        if (rank == 0 && tid == 0) a.loss[ep] = loss * inv_n;
        float part = 0.f;
        for (int e = e0 + tid; e < e1; e += T) {
            float g = 0.f;
            for (int r = 0; r < kCluster; ++r) g += ms_cluster_map(G, (unsigned)r)[e];
            G[e] = g;                                                   // nobody else reads slice `rank` of this CTA's G
            part += g * g;
        }
        const float slice_sq = cta_sum<kCThreads>(part, S + CRED);
        if (tid == 0) xch[2] = slice_sq;
        ms_cluster_sync();                                              // barrier 2: slice sums of squares
        // ---- 3. clip factor, Adam on the slice, new parameters to every CTA
        float total = 0.f;
        for (int r = 0; r < kCluster; ++r) total += ms_cluster_map(xch, (unsigned)r)[2];
        const float coef = fminf(fmax_norm / (sqrtf(total) + 1e-6f), 1.0f);
        b1pow *= a.beta1;
        b2pow *= a.beta2;
        const float step_size = (float)(a.lr / (1.0 - b1pow));
        const float bc2_sqrt = (float)sqrt(1.0 - b2pow);
        for (int e = e0 + tid; e < e1; e += T) {
            const float g = G[e] * coef;
            float m = S[CM + e - e0], v = S[CV + e - e0];
            m = m + (g - m) * w1;
            v = v * fb2 + (w2 * g) * g;
            S[CM + e - e0] = m;
            S[CV + e - e0] = v;
            const float denom = sqrtf(v) / bc2_sqrt + feps;
            const int at = smem_of(e);
            const float p = S[at] - step_size * (m / denom);
            for (int r = 0; r < kCluster; ++r) ms_cluster_map(S, (unsigned)r)[at] = p;
        }
        ms_cluster_sync();                                              // barrier 3: parameters in place everywhere
    }
    __syncthreads();
    for (int e = e0 + tid; e < e1; e += T) {
        a.net[e] = S[smem_of(e)];
        a.adam_m[e] = S[CM + e - e0];
        a.adam_v[e] = S[CV + e - e0];
    }
    ms_cluster_sync();                                                  // nobody leaves while its shared memory may be read
}

}  // namespace ms
