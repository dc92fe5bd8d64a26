// scopa_b200/csrc/ms_sd_sample.cuh -- sd_sample_rows_kernel: the minibatches of AdvantageNetwork.train
// (`random.sample(self.buffer, batch_size)` once per epoch, /root/reference/src/algorithms/deep_cfr/deep_cfr.py:88) for
// all epochs of a train() call in ONE launch: block = epoch, thread m = position m of that epoch's minibatch.
//
// Sampling without replacement by parallel rejection.  Every position draws a row from the counter-based Philox4x32-10
// stream  x = philox(key = seed, ctr = (epoch lo, epoch hi, m | attempt << 8, "SDTR")),  row = mulhi32(x.x, n_rows);
// a position whose row equals the row of a LOWER position draws again (attempt + 1), until no two positions agree.
// The procedure commutes with every relabelling of the rows, so all batches of distinct rows are equally likely, and
// its outcome is a function of (seed, epoch, n_rows, batch) alone: runs are repeatable and do not depend on scheduling
// (the host restatement in tests/test_sd_train_emu.py reproduces it draw for draw).  `epoch` is the global optimiser
// step (steps done before the call + epoch of the call), so consecutive calls continue the stream.
// Expected rounds: 1 + batch^2 / (2 n_rows) for a large buffer; batch == n_rows (a permutation of a tiny buffer) needs
// about n log n redraws.  A round limit (4096) marks the epoch unusable (rows = -1: the optimiser kernels skip such an
// epoch and report NaN); the probability of reaching it is below 1e-50 for every legal (batch, n_rows).
//
// Emulation-compatible subset of CUDA (tests/emu/cta_emu.h): 64-bit products instead of __umulhi.
#pragma once
#include <stdint.h>

namespace ms {

struct SdSampleArgs {
    int* idx;                        // [epochs][batch] out
    int batch;                       // 1 .. 128
    int epochs;
    long long n_rows;                // batch <= n_rows < 2^31
    unsigned long long seed;
    unsigned long long first_epoch;  // optimiser steps done before this call
};

namespace sds {
constexpr int kSampleThreads = 128, kMaxRounds = 4096;
constexpr uint32_t kTagSdTrain = 0x52544453u;   // "SDTR"

__device__ __forceinline__ uint32_t philox_first_word(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                      uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1,
                       n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}
}  // namespace sds

__global__ void __launch_bounds__(sds::kSampleThreads) sd_sample_rows_kernel(SdSampleArgs a) {
    using namespace sds;
    __shared__ int row[kSampleThreads];
    __shared__ int n_dirty;
    const int m = (int)threadIdx.x, B = a.batch;
    for (int ep = (int)blockIdx.x; ep < a.epochs; ep += (int)gridDim.x) {
        const unsigned long long E = a.first_epoch + (unsigned long long)ep;
        uint32_t attempt = 0;
        bool dirty = m < B;
        int rounds = 0;
        for (;;) {
            if (dirty) {
                const uint32_t x = philox_first_word((uint32_t)E, (uint32_t)(E >> 32), (uint32_t)m | (attempt << 8), kTagSdTrain,
                                                     (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                row[m] = (int)(((uint64_t)x * (uint64_t)a.n_rows) >> 32);
                ++attempt;
            }
            if (m == 0) n_dirty = 0;
            __syncthreads();
            dirty = false;
            if (m < B)
                for (int j = 0; j < m; ++j) dirty = dirty || row[j] == row[m];
            if (dirty) n_dirty = 1;                      // (benign race: every writer stores 1)
            __syncthreads();
            const bool again = n_dirty != 0;
            __syncthreads();                             // everyone has read the flag before the next round resets it
            if (!again || ++rounds >= kMaxRounds) break;
        }
        if (m < B) a.idx[(long long)ep * B + m] = dirty ? -1 : row[m];
        __syncthreads();
    }
}

}  // namespace ms
