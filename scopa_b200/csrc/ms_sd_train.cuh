// scopa_b200/csrc/ms_sd_train.cuh -- sd_train_kernel: `epochs` optimiser steps of AdvantageNetwork.train
// (/root/reference/src/algorithms/deep_cfr/deep_cfr.py:77-110) in ONE launch of ONE CTA.
//
// The reference does, per epoch: sample a minibatch (<= 128 rows) from the replay buffer, forward the
// 34 -> 128 -> 64 -> 16 ReLU MLP, loss = MSELoss(pred * mask, target * mask), backward,
// clip_grad_norm_(max_norm = 1.0), Adam(lr = 5e-4).  In PyTorch that is ~40 launches and one device->host
// read (loss.item()) per epoch for 10 MFLOP of arithmetic, i.e. pure launch latency.  Here the weights
// (13 776 floats), the minibatch and every activation stay in the shared memory of one SM for all epochs:
//
//   shared memory (floats, rows padded to ODD strides so that column-strided reads are conflict-free):
//     params  W1[128][35] b1[128] W2[64][129] b2[64] W3[16][65] b3[16]      13 984
//     x       [128][35]   minibatch features                                  4 480
//     h1      [128][129]  layer-1 activations, later d(loss)/d(h1) in place  16 512
//     h2      [128][65]   layer-2 activations, later d(loss)/d(h2) in place   8 320
//     d       [128][17]   d(loss)/d(out)                                      2 176
//     red     [512 + 32]  block-reduction scratch                               544
//                                                                    total   46 016 floats = 184 064 B
//   global memory: the fp32 net blob (read once, written once), Adam's exp_avg / exp_avg_sq (one read-modify-
//   write per step, L2 resident), the gradient scratch (written and re-read by this CTA only), the gathered
//   replay rows (<= 128 x 66 floats per epoch) and one loss per epoch.
//
// Arithmetic: fp32 with explicit fmaf in k-ascending order (the reference's precision; torch's own summation
// order inside sgemm is unspecified, so parity with torch is to tolerance: tests/test_sd_train_emu.py).
//
// This file has no CUDA-only constructs besides threadIdx / blockDim / __syncthreads / __shared__: it also compiles
// as host C++ behind tests/emu/cta_emu.h (one pthread per CUDA thread, pthread barrier = __syncthreads), which is
// how the kernel's logic is checked against torch on machines without a GPU.  Same source, fmaf everywhere and no
// contraction elsewhere (nvcc --fmad=false / g++ -ffp-contract=off), so the two builds agree bit for bit.
#pragma once
#include "ms_div.cuh"

#ifndef MS_CTA_EMU
#define MS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace ms {

struct SdTrainArgs {
    float* net;            // [13776] blob, nn.Linear order; updated in place
    float* adam_m;         // [13776] exp_avg
    float* adam_v;         // [13776] exp_avg_sq
    const float* feat;     // [n_rows][34] replay buffer
    const float* target;   // [n_rows][16]
    const float* mask;     // [n_rows][16]
    long long n_rows;
    const int* idx;        // [epochs][batch] rows of each minibatch
    int batch;             // 1 .. 128
    int epochs;
    double lr, beta1, beta2, eps, max_norm;
    double b1pow, b2pow;   // beta1 ** steps_done, beta2 ** steps_done
    float* loss;           // [epochs] MSE of each minibatch (NaN when an index was out of range; that step is skipped)
    float* grad;           // [13776] scratch
};

namespace sdt {
constexpr int kThreads = 512;
constexpr int kIn = 34, kH1 = 128, kH2 = 64, kOut = 16, kMaxBatch = 128;
constexpr int LDX = 35, LD1 = 129, LD2 = 65, LDO = 17;
// blob offsets
constexpr int OW1 = 0, OB1 = OW1 + kH1 * kIn, OW2 = OB1 + kH1, OB2 = OW2 + kH2 * kH1, OW3 = OB2 + kH2,
              OB3 = OW3 + kOut * kH2, kNetFloats = OB3 + kOut;
static_assert(kNetFloats == 13776, "net blob size");
// shared-memory offsets (floats)
constexpr int SW1 = 0, SB1 = SW1 + kH1 * LDX, SW2 = SB1 + kH1, SB2 = SW2 + kH2 * LD1, SW3 = SB2 + kH2,
              SB3 = SW3 + kOut * LD2, SX = SB3 + kOut, SH1 = SX + kMaxBatch * LDX, SH2 = SH1 + kMaxBatch * LD1,
              SD = SH2 + kMaxBatch * LD2, SRED = SD + kMaxBatch * LDO, kSmemFloats = SRED + kThreads + 32;
constexpr int kSmemBytes = kSmemFloats * 4;
static_assert(kSmemBytes <= 227 * 1024, "shared memory per CTA");

// blob index -> index of the same parameter in the padded shared-memory image
__device__ __forceinline__ int smem_of(int e) {
    if (e < OB1) return SW1 + (e / kIn) * LDX + e % kIn;
    if (e < OW2) return SB1 + (e - OB1);
    if (e < OB2) return SW2 + ((e - OW2) / kH1) * LD1 + (e - OW2) % kH1;
    if (e < OW3) return SB2 + (e - OB2);
    if (e < OB3) return SW3 + ((e - OW3) / kH2) * LD2 + (e - OW3) % kH2;
    return SB3 + (e - OB3);
}

// C(i, j) = sum_k A(i, k) * B(j, k), A(i, k) = A[i * a_i + k * a_k], B(j, k) = B[j * b_j + k * b_k], k ascending,
// handed to epi(i, j, value) for every i < M, j < N.  A thread owns a TM x TN register tile: rows gi*TM .. +TM-1,
// columns gj, gj + GJ, ... (interleaved, so that the lanes of a warp read consecutive columns of B).
template <int TM, int TN, class Epi>
__device__ __forceinline__ void cta_gemm(int M, int N, int K, const float* A, int a_i, int a_k, const float* B, int b_j,
                                         int b_k, Epi epi) {
    const int GI = (M + TM - 1) / TM, GJ = (N + TN - 1) / TN;
    for (int t = (int)threadIdx.x; t < GI * GJ; t += (int)blockDim.x) {
        const int gi = t / GJ, gj = t % GJ;
        const float* ap[TM];
        const float* bp[TN];
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            int i = gi * TM + r;
            ap[r] = A + (i < M ? i : M - 1) * a_i;      // clamped: out-of-range rows compute, but are not stored
        }
#pragma unroll
        for (int c = 0; c < TN; ++c) {
            int j = gj + GJ * c;
            bp[c] = B + (j < N ? j : N - 1) * b_j;
        }
        float acc[TM][TN];
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[r][c] = 0.f;
        for (int k = 0; k < K; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) a[r] = ap[r][k * a_k];
#pragma unroll
            for (int c = 0; c < TN; ++c) b[c] = bp[c][k * b_k];
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < TM; ++r)
#pragma unroll
            for (int c = 0; c < TN; ++c) {
                int i = gi * TM + r, j = gj + GJ * c;
                if (i < M && j < N) epi(i, j, acc[r][c]);
            }
    }
}

// sum of v over a CTA of NT threads in a fixed order; every thread gets the same float.  red: NT + 32 floats.
template <int NT>
__device__ __forceinline__ float cta_sum(float v, float* red) {
    static_assert(NT % 32 == 0, "whole warps");
    const int tid = (int)threadIdx.x, per = NT / 32;
    red[tid] = v;
    __syncthreads();
    if (tid < 32) {
        float s = 0.f;
        for (int j = 0; j < per; ++j) s += red[tid * per + j];
        red[NT + tid] = s;
    }
    __syncthreads();
    float s = 0.f;
    for (int j = 0; j < 32; ++j) s += red[NT + j];
    __syncthreads();
    return s;
}

// column sums of a [M][ld] shared-memory matrix -> g[0..N) (bias gradients)
__device__ __forceinline__ void col_sums(const float* D, int ld, int M, int N, float* g) {
    for (int n = (int)threadIdx.x; n < N; n += (int)blockDim.x) {
        float s = 0.f;
        for (int m = 0; m < M; ++m) s += D[m * ld + n];
        g[n] = s;
    }
}
}  // namespace sdt

__global__ void __launch_bounds__(sdt::kThreads, 1) sd_train_kernel(SdTrainArgs a) {
    using namespace sdt;
    MS_DYN_SMEM(sd_train_smem);
    float* S = reinterpret_cast<float*>(sd_train_smem);
    __shared__ int bad_index;
    const int tid = (int)threadIdx.x, T = (int)blockDim.x;
    const int B = a.batch;
    const int Mp = (B + 3) & ~3;                         // rows processed (padding rows are zero and carry no gradient)
    const float inv_n = 1.0f / (float)(B * kOut);         // MSELoss: mean over batch x 16
    const float w1 = (float)(1.0 - a.beta1), fb2 = (float)a.beta2, w2 = (float)(1.0 - a.beta2);
    const float feps = (float)a.eps, fmax_norm = (float)a.max_norm;
    double b1pow = a.b1pow, b2pow = a.b2pow;

    for (int e = tid; e < kNetFloats; e += T) S[smem_of(e)] = a.net[e];

    for (int ep = 0; ep < a.epochs; ++ep) {
        const int* idx = a.idx + (long long)ep * B;
        if (tid == 0) bad_index = 0;
        __syncthreads();                                 // also: parameters of the previous step are in place
        // ---- gather the minibatch features
        for (int t = tid; t < Mp * kIn; t += T) {
            int m = t / kIn, k = t % kIn;
            float v = 0.f;
            if (m < B) {
                long long row = idx[m];
                if (row < 0 || row >= a.n_rows) {
                    bad_index = 1;
                    row = 0;
                }
                v = a.feat[row * kIn + k];
            }
            S[SX + m * LDX + k] = v;
        }
        __syncthreads();
        if (bad_index) {                                 // uniform: read after the barrier
            if (tid == 0) a.loss[ep] = __int_as_float(0x7fc00000);
            __syncthreads();                             // everyone has read the flag before it is reset
            continue;
        }
        // ---- forward
        cta_gemm<4, 4>(Mp, kH1, kIn, S + SX, LDX, 1, S + SW1, LDX, 1,
                       [&](int i, int j, float v) { S[SH1 + i * LD1 + j] = fmaxf(v + S[SB1 + j], 0.f); });
        __syncthreads();
        cta_gemm<4, 4>(Mp, kH2, kH1, S + SH1, LD1, 1, S + SW2, LD1, 1,
                       [&](int i, int j, float v) { S[SH2 + i * LD2 + j] = fmaxf(v + S[SB2 + j], 0.f); });
        __syncthreads();
        float sq = 0.f;
        cta_gemm<4, 4>(Mp, kOut, kH2, S + SH2, LD2, 1, S + SW3, LD2, 1, [&](int i, int j, float v) {
            float d = 0.f;
            if (i < B) {
                long long row = idx[i];
                float mk = a.mask[row * kOut + j], tg = a.target[row * kOut + j];
                float diff = (v + S[SB3 + j]) * mk - tg * mk;
                sq += diff * diff;
                d = (2.0f * diff * inv_n) * mk;
            }
            S[SD + i * LDO + j] = d;
        });
        const float loss = cta_sum<kThreads>(sq, S + SRED) * inv_n;    // (barriers inside: d is complete afterwards)
        if (tid == 0) a.loss[ep] = loss;
        // ---- backward: layer 3
        cta_gemm<4, 4>(kOut, kH2, Mp, S + SD, 1, LDO, S + SH2, 1, LD2,
                       [&](int i, int j, float v) { a.grad[OW3 + i * kH2 + j] = v; });
        col_sums(S + SD, LDO, Mp, kOut, a.grad + OB3);
        __syncthreads();                                 // h2 is about to be overwritten
        cta_gemm<4, 4>(Mp, kH2, kOut, S + SD, LDO, 1, S + SW3, 1, LD2, [&](int i, int j, float v) {
            float& h = S[SH2 + i * LD2 + j];
            h = h > 0.f ? v : 0.f;
        });
        __syncthreads();
        // ---- layer 2
        cta_gemm<4, 4>(kH2, kH1, Mp, S + SH2, 1, LD2, S + SH1, 1, LD1,
                       [&](int i, int j, float v) { a.grad[OW2 + i * kH1 + j] = v; });
        col_sums(S + SH2, LD2, Mp, kH2, a.grad + OB2);
        __syncthreads();                                 // h1 is about to be overwritten
        cta_gemm<4, 4>(Mp, kH1, kH2, S + SH2, LD2, 1, S + SW2, 1, LD1, [&](int i, int j, float v) {
            float& h = S[SH1 + i * LD1 + j];
            h = h > 0.f ? v : 0.f;
        });
        __syncthreads();
        // ---- layer 1
        cta_gemm<4, 4>(kH1, kIn, Mp, S + SH1, 1, LD1, S + SX, 1, LDX,
                       [&](int i, int j, float v) { a.grad[OW1 + i * kIn + j] = v; });
        col_sums(S + SH1, LD1, Mp, kH1, a.grad + OB1);
        __syncthreads();                                 // the CTA's own global writes are visible to it from here
        // ---- clip_grad_norm_(max_norm) and Adam
        float part = 0.f;
        for (int e = tid; e < kNetFloats; e += T) {
            float g = a.grad[e];
            part += g * g;
        }
        const float norm = sqrtf(cta_sum<kThreads>(part, S + SRED));
        const float coef = fminf(fmax_norm / (norm + 1e-6f), 1.0f);
        b1pow *= a.beta1;
        b2pow *= a.beta2;
        const float step_size = (float)(a.lr / (1.0 - b1pow));
        const float bc2_sqrt = (float)sqrt(1.0 - b2pow);
        for (int e = tid; e < kNetFloats; e += T) {
            const float g = a.grad[e] * coef;
            float m = a.adam_m[e], v = a.adam_v[e];
            m = m + (g - m) * w1;                        // exp_avg.lerp_(grad, 1 - beta1)
            v = v * fb2 + (w2 * g) * g;                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
            a.adam_m[e] = m;
            a.adam_v[e] = v;
            const float denom = ms_div_or_zero(sqrtf(v), bc2_sqrt) + feps;      // v and m are exactly 0 for weights no
            float& p = S[smem_of(e)];                                           // gradient has reached yet
            p = p - step_size * ms_div_or_zero(m, denom);
        }
    }
    __syncthreads();
    for (int e = tid; e < kNetFloats; e += T) a.net[e] = S[smem_of(e)];
}

}  // namespace ms
