// tests/emu/cta_emu.h -- runs ONE CUDA thread block on the host: one pthread per CUDA thread, a pthread barrier
// for __syncthreads(), function-local statics for __shared__.  Test infrastructure: lets kernels written with
// nothing but threadIdx / blockDim / __syncthreads / __shared__ (no warp intrinsics, no atomics) be executed and
// checked on machines without a GPU.  Not a performance model and not a race detector.
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
#define MS_DYN_SMEM(name) alignas(16) static unsigned char name[232448]

struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 threadIdx;
static emu_dim3 blockDim = {1, 1, 1}, blockIdx = {0, 0, 0}, gridDim = {1, 1, 1};
static pthread_barrier_t emu_barrier;
static inline void __syncthreads() { pthread_barrier_wait(&emu_barrier); }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }

template <class Kernel, class Args>
struct emu_launch_ctx { Kernel k; const Args* a; unsigned tid; };

template <class Kernel, class Args>
static void* emu_thread_main(void* p) {
    auto* c = static_cast<emu_launch_ctx<Kernel, Args>*>(p);
    threadIdx = {c->tid, 0, 0};
    c->k(*c->a);
    return nullptr;
}

// kernel<<<1, threads>>>(args)
template <class Kernel, class Args>
static int emu_launch_cta(Kernel k, const Args& args, unsigned threads) {
    blockDim = {threads, 1, 1};
    if (pthread_barrier_init(&emu_barrier, nullptr, threads)) return -1;
    std::vector<pthread_t> th(threads);
    std::vector<emu_launch_ctx<Kernel, Args>> ctx(threads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    int rc = 0;
    unsigned started = 0;
    for (; started < threads; ++started) {
        ctx[started] = {k, &args, started};
        if (pthread_create(&th[started], &attr, emu_thread_main<Kernel, Args>, &ctx[started])) { rc = -2; break; }
    }
    if (rc) return rc;   // (a partial start would dead-lock on the barrier; the caller treats it as fatal)
    for (unsigned i = 0; i < started; ++i) pthread_join(th[i], nullptr);
    pthread_attr_destroy(&attr);
    pthread_barrier_destroy(&emu_barrier);
    return 0;
}

// kernel<<<blocks, threads>>>(args) for kernels whose blocks do not communicate: the blocks run one after another
template <class Kernel, class Args>
static int emu_launch_grid(Kernel k, const Args& args, unsigned blocks, unsigned threads) {
    gridDim = {blocks, 1, 1};
    for (unsigned b = 0; b < blocks; ++b) {
        blockIdx = {b, 0, 0};
        int rc = emu_launch_cta(k, args, threads);
        if (rc) return rc;
    }
    blockIdx = {0, 0, 0};
    gridDim = {1, 1, 1};
    return 0;
}
