// scopa_b200/csrc/ms_common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include <mutex>

#include "../../include/scopa_b200.h"

namespace ms {

char* last_error_buf();                      // thread-local, defined in ms_env.cu
extern std::atomic<uint64_t> g_launches;     // kernels launched by this library

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define MS_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ms::fail(MS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                              \
    } while (0)

#define MS_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        ms::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return ms::fail(MS_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                  \
                            cudaGetErrorString(_e), __FILE__, __LINE__);                      \
    } while (0)

// Plumbing of the *_host entry points (defined in ms_env.cu): one grow-only device scratch buffer per device,
// guarded by g_scratch_mu for the duration of a call, three non-blocking streams so that the H2D copy of stage
// c+1, the kernels of stage c and the D2H copy of stage c-1 overlap, and the stage size in games.
extern std::mutex g_scratch_mu;
int scratch_get(size_t bytes, char** out, cudaStream_t* stream);
int host_pipe_streams(cudaStream_t out[3]);
int64_t host_chunk();
// Size of the pipeline stage that starts at game `lo` of `n` (ms_env.cu): full stages of host_chunk() games between a SHORT
// first stage (the GPU starts after a quarter-size copy instead of idling through a full one) and a short last stage (the
// device->host copy nothing overlaps with is a quarter size too).
int64_t host_stage_size(int64_t lo, int64_t n);
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

constexpr int kNumSMs = 148;   // B200

// grid for a grid-stride kernel: enough CTAs to cover n, capped at a multiple of the SM count
inline int grid_for(int64_t n, int block, int ctas_per_sm) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace ms
