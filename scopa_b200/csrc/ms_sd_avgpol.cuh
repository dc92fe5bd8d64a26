// scopa_b200/csrc/ms_sd_avgpol.cuh -- StrategyBuffer.get_average_policy
// (/root/reference/src/algorithms/deep_cfr/deep_cfr.py:136-160) for a batch of states and ALL stored strategy nets in
// two launches:   policy[row] = sum_k  positive_regret_policy(net_k(feat[row]), mask[row]) * (weight_k / total_weight)
// with positive_regret_policy = relu(adv) * mask / max(sum, 1e-8) (nets.py:93-101).  The reference (and a PyTorch
// restatement of it) runs one batch-1 forward per stored net: up to 100 nets x ~10 launches per decision of
// evaluate_vs_random.  Here CTA (k, y) keeps net k (13 776 floats) in shared memory, forwards its share of the rows (64-row
// chunks y, y + gridDim.y, ...) with
// the same fp32 fmaf GEMM routine as sd_train_kernel, and writes its weighted policy to scratch[k][row][16];
// sd_avgpol_reduce_kernel then adds the K layers in k order (the reference's `policy +=` order), so the result does not
// depend on scheduling.  Same emulation-compatible subset of CUDA as ms_sd_train.cuh (tests/emu).
#pragma once
#include "ms_sd_train.cuh"

namespace ms {

struct SdAvgPolArgs {
    const float* nets;      // [n_nets][13776] blobs, nn.Linear order
    const float* weights;   // [n_nets] weight_k / total_weight, already in fp32
    int n_nets;
    const float* feat;      // [n_rows][34]
    const float* mask;      // [n_rows][16]
    long long n_rows;
    float* scratch;         // [n_nets][n_rows][16]
    float* policy;          // [n_rows][16]
};

namespace sda {
using namespace sdt;
constexpr int kRows = 64, kPolThreads = 512;
// shared-memory offsets (floats): parameters as in sd_train_kernel, then one 64-row chunk
constexpr int PX = SB3 + kOut, PH1 = PX + kRows * LDX, PH2 = PH1 + kRows * LD1, PO = PH2 + kRows * LD2,
              kPolSmemFloats = PO + kRows * LDO;
constexpr int kPolSmemBytes = kPolSmemFloats * 4;
static_assert(kPolSmemBytes <= 227 * 1024, "shared memory per CTA");
}  // namespace sda

__global__ void __launch_bounds__(sda::kPolThreads, 1) sd_avgpol_kernel(SdAvgPolArgs a) {
    using namespace sda;
    MS_DYN_SMEM(sd_avgpol_smem);
    float* S = reinterpret_cast<float*>(sd_avgpol_smem);
    const int tid = (int)threadIdx.x, T = (int)blockDim.x;
    for (int k = (int)blockIdx.x; k < a.n_nets; k += (int)gridDim.x) {
        __syncthreads();                                  // the previous net's last chunk is done with the parameters
        const float* net = a.nets + (long long)k * kNetFloats;
        for (int e = tid; e < kNetFloats; e += T) S[smem_of(e)] = net[e];
        const float wk = a.weights[k];
        float* out = a.scratch + (long long)k * a.n_rows * kOut;
        // blockIdx.y / gridDim.y split the 64-row chunks of one net over several CTAs (few nets, many rows)
        for (long long r0 = (long long)blockIdx.y * kRows; r0 < a.n_rows; r0 += (long long)gridDim.y * kRows) {
            const int R = (int)((a.n_rows - r0) < kRows ? (a.n_rows - r0) : kRows);
            const int Rp = (R + 3) & ~3;
            __syncthreads();                              // parameters loaded / previous chunk consumed
            for (int t = tid; t < Rp * kIn; t += T) {
                int m = t / kIn, c = t % kIn;
                S[PX + m * LDX + c] = m < R ? a.feat[(r0 + m) * kIn + c] : 0.f;
            }
            __syncthreads();
            cta_gemm<4, 4>(Rp, kH1, kIn, S + PX, LDX, 1, S + SW1, LDX, 1,
                           [&](int i, int j, float v) { S[PH1 + i * LD1 + j] = fmaxf(v + S[SB1 + j], 0.f); });
            __syncthreads();
            cta_gemm<4, 4>(Rp, kH2, kH1, S + PH1, LD1, 1, S + SW2, LD1, 1,
                           [&](int i, int j, float v) { S[PH2 + i * LD2 + j] = fmaxf(v + S[SB2 + j], 0.f); });
            __syncthreads();
            cta_gemm<4, 4>(Rp, kOut, kH2, S + PH2, LD2, 1, S + SW3, LD2, 1,
                           [&](int i, int j, float v) { S[PO + i * LDO + j] = v + S[SB3 + j]; });
            __syncthreads();
            // regret matching, one thread per row: pos = relu(adv) * mask, z = max(sum, 1e-8), policy = pos / z
            for (int m = tid; m < R; m += T) {
                const float* mk = a.mask + (r0 + m) * kOut;
                float pos[kOut], z = 0.f;
#pragma unroll
                for (int j = 0; j < kOut; ++j) {
                    pos[j] = fmaxf(S[PO + m * LDO + j], 0.f) * mk[j];
                    z += pos[j];
                }
                z = fmaxf(z, 1e-8f);
#pragma unroll
                for (int j = 0; j < kOut; ++j) out[(r0 + m) * kOut + j] = ms_div_or_zero(pos[j], z) * wk;
            }
        }
    }
}

// policy[row][j] = sum over k (ascending) of scratch[k][row][j]
__global__ void __launch_bounds__(256) sd_avgpol_reduce_kernel(SdAvgPolArgs a) {
    const long long n = a.n_rows * sdt::kOut;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < a.n_nets; ++k) acc += a.scratch[(long long)k * n + e];
        a.policy[e] = acc;
    }
}

}  // namespace ms
