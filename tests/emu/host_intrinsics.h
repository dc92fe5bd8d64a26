// tests/emu/host_intrinsics.h -- host meanings of the CUDA intrinsics the product's kernels use, shared by the host builds
// of tests/emu (ms_*_host.cpp).  Test infrastructure.
//   * integer / bit intrinsics: the compiler builtins;
//   * __dadd_rn / __dmul_rn / __ddiv_rn: IEEE double add / mul / div (the files are compiled with -ffp-contract=off);
//   * atomics: real atomics (a CAS loop for the fp64 add, like the shared-memory form on the device);
//   * warp votes that the kernels use ONLY to share a loop bound between the lanes of a warp (__reduce_max_sync,
//     __any_sync over __activemask()): the lane's own value -- a lane never needs more iterations than its own bound.
#pragma once
#include <cstdint>
#include <cstring>

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __activemask() { return 1u; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return v; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcv(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void*) { return 0; }   // feeds the tcgen05 / mbarrier PTX only, never run

static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return cmp;                                   // the value found: `cmp` itself on success, the other owner's key otherwise
}
static inline double atomicAdd(double* p, double v) {
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = __atomic_load_n(q, __ATOMIC_RELAXED), want;
    double cur;
    do {
        std::memcpy(&cur, &old, 8);
        cur += v;
        std::memcpy(&want, &cur, 8);
    } while (!__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    std::memcpy(&cur, &old, 8);
    return cur;
}

// system-scope forms (peer memory over NVLink on the device): the same host atomics
static inline unsigned atomicAdd_system(unsigned* p, unsigned v) { return atomicAdd(p, v); }
static inline double atomicAdd_system(double* p, double v) { return atomicAdd(p, v); }
static inline unsigned atomicOr_system(unsigned* p, unsigned v) { return atomicOr(p, v); }
static inline unsigned long long atomicCAS_system(unsigned long long* p, unsigned long long cmp, unsigned long long val) { return atomicCAS(p, cmp, val); }

// CUDA's function / variable qualifiers (cuda_runtime.h defines some of them for the host compiler): cleared here; the
// one-thread builds then define the few they need, the CTA-emulator builds include cta_emu.h
#undef __device__
#undef __global__
#undef __host__
#undef __shared__
#undef __constant__
#undef __forceinline__
#undef __noinline__
#undef __launch_bounds__
#undef __align__
#define __host__
#define __constant__
#define __noinline__
#define __align__(n) alignas(n)

#ifdef MS_HOST_ONE_THREAD
// "a grid of one block of one thread": every kernel of these files is a grid-stride loop without block-wide steps other
// than staging a table, so this runs every row in order
#define __device__
#define __global__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
static inline void __syncthreads() {}
struct host_idx { unsigned x; };
static const host_idx blockIdx = {0}, threadIdx = {0}, blockDim = {1}, gridDim = {1};
#endif
