// Microbenchmark behind the deal kernel's seeding chain (profiles/README.md, section 2): where do the
// ~1870 dependent MT19937 init_by_array steps per seed spend their time?  Variants of the same arithmetic:
//   V0  seed-independent table read from the constant bank with a register index (LDC per step)
//   V1  table staged in shared memory, one 128-bit broadcast load per four steps
//   V2  V1 with 8 CTAs of 256 threads per SM forced (32 registers)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/deal_variants profiles/deal_variants.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__constant__ uint32_t c_init[624];
constexpr int WIN = 40;

__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    return y;
}

__device__ __forceinline__ unsigned long long shuffle16(const uint32_t* lo, const uint32_t* hi) {
    unsigned long long perm = 0xFEDCBA9876543210ull;
    int kk = 0;
    for (int i = 15; i >= 1; i--) {
        const uint32_t nn = (uint32_t)i + 1u;
        const int kbits = 32 - __clz(nn);
        uint32_t r;
        do {
            if (kk >= WIN) return 0ull;
            uint32_t y = (lo[kk] & 0x80000000u) | (lo[kk + 1] & 0x7fffffffu);
            y = hi[kk] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            kk++;
            r = temper(y) >> (32 - kbits);
        } while (r >= nn);
        unsigned long long ci = (perm >> (4 * i)) & 0xFull, cr = (perm >> (4 * r)) & 0xFull;
        perm &= ~((0xFull << (4 * i)) | (0xFull << (4 * r)));
        perm |= (cr << (4 * i)) | (ci << (4 * r));
    }
    return perm;
}

// ---------------------------------------------------------------- V0: constant bank, register index
__device__ __forceinline__ void seed_v0(uint32_t key0, uint32_t key1, uint32_t* lo, uint32_t* hi) {
    const bool two = key1 != 0u;
    const uint32_t kodd = two ? key1 + 1u : key0;
    uint32_t prev = (c_init[1] ^ ((c_init[0] ^ (c_init[0] >> 30)) * 1664525u)) + key0;
    const uint32_t first1 = prev;
#pragma unroll 8
    for (int k = 1; k < 623; k += 2) {
        prev = (c_init[k + 1] ^ ((prev ^ (prev >> 30)) * 1664525u)) + kodd;
        prev = (c_init[k + 2] ^ ((prev ^ (prev >> 30)) * 1664525u)) + key0;
    }
    const uint32_t m1w = (first1 ^ ((prev ^ (prev >> 30)) * 1664525u)) + kodd;
    uint32_t p1 = first1, p2 = m1w;
    auto lock_step = [&](int i, uint32_t kw) {
        p1 = (c_init[i] ^ ((p1 ^ (p1 >> 30)) * 1664525u)) + kw;
        p2 = (p1 ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - (uint32_t)i;
    };
#pragma unroll 1
    for (int i = 2; i <= WIN; i += 2) { lock_step(i, kodd); lo[i] = p2; lock_step(i + 1, key0); lo[i + 1] = p2; }
    lock_step(WIN + 2, kodd);
#pragma unroll 8
    for (int i = WIN + 3; i < 397; i += 2) { lock_step(i, key0); lock_step(i + 1, kodd); }
#pragma unroll 1
    for (int i = 397; i < 397 + WIN; i += 2) { lock_step(i, key0); hi[i - 397] = p2; lock_step(i + 1, kodd); hi[i - 396] = p2; }
#pragma unroll 8
    for (int i = 397 + WIN; i < 623; i += 2) { lock_step(i, key0); lock_step(i + 1, kodd); }
    lock_step(623, key0);
    lo[1] = (m1w ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - 1u;
    lo[0] = 0x80000000u;
}

__global__ void __launch_bounds__(256) k_v0(const long long* seeds, long long n, unsigned long long* out) {
    uint32_t lo[WIN + 2], hi[WIN];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        unsigned long long a = (unsigned long long)seeds[g];
        seed_v0((uint32_t)a, (uint32_t)(a >> 32), lo, hi);
        out[g] = shuffle16(lo, hi);
    }
}

// ---------------------------------------------------------------- V1: shared-memory table, 128-bit broadcast loads
// word i is written by step k = i - 1 with key word j = k % len: even i -> kodd, odd i -> key0
#define P1STEP(x, t, kw) (x) = ((t) ^ (((x) ^ ((x) >> 30)) * 1664525u)) + (kw)
#define P2STEP(y, p, i) (y) = ((p) ^ (((y) ^ ((y) >> 30)) * 1566083941u)) - (uint32_t)(i)

__device__ __forceinline__ void seed_v1(const uint4* __restrict__ T4, uint32_t key0, uint32_t key1, uint32_t* lo, uint32_t* hi) {
    const bool two = key1 != 0u;
    const uint32_t kodd = two ? key1 + 1u : key0;
    uint4 t = T4[0];
    uint32_t prev = (t.y ^ ((t.x ^ (t.x >> 30)) * 1664525u)) + key0;   // word 1
    const uint32_t first1 = prev;
    P1STEP(prev, t.z, kodd);
    P1STEP(prev, t.w, key0);
#pragma unroll 4
    for (int q = 1; q < 156; q++) {
        t = T4[q];
        P1STEP(prev, t.x, kodd); P1STEP(prev, t.y, key0); P1STEP(prev, t.z, kodd); P1STEP(prev, t.w, key0);
    }
    const uint32_t m1w = (first1 ^ ((prev ^ (prev >> 30)) * 1664525u)) + kodd;
    uint32_t p1 = first1, p2 = m1w;
    t = T4[0];
    P1STEP(p1, t.z, kodd); P2STEP(p2, p1, 2); lo[2] = p2;
    P1STEP(p1, t.w, key0); P2STEP(p2, p1, 3); lo[3] = p2;
    // groups 1 .. 10 hold words 4 .. 43: words 4 .. WIN + 1 = 41 are kept
#pragma unroll 1
    for (int q = 1; q < (WIN + 4) / 4; q++) {
        t = T4[q];
        const int i = 4 * q;
        P1STEP(p1, t.x, kodd); P2STEP(p2, p1, i);     lo[i] = p2;
        P1STEP(p1, t.y, key0); P2STEP(p2, p1, i + 1); lo[i + 1] = p2;
        P1STEP(p1, t.z, kodd); P2STEP(p2, p1, i + 2); if (i + 2 <= WIN + 1) lo[i + 2] = p2;
        P1STEP(p1, t.w, key0); P2STEP(p2, p1, i + 3); if (i + 3 <= WIN + 1) lo[i + 3] = p2;
    }
#pragma unroll 4
    for (int q = (WIN + 4) / 4; q < 99; q++) {   // words 44 .. 395
        t = T4[q];
        const int i = 4 * q;
        P1STEP(p1, t.x, kodd); P2STEP(p2, p1, i);
        P1STEP(p1, t.y, key0); P2STEP(p2, p1, i + 1);
        P1STEP(p1, t.z, kodd); P2STEP(p2, p1, i + 2);
        P1STEP(p1, t.w, key0); P2STEP(p2, p1, i + 3);
    }
    // groups 99 .. 109 hold words 396 .. 439: words 397 .. 397 + WIN - 1 = 436 are kept
#pragma unroll 1
    for (int q = 99; q < 110; q++) {
        t = T4[q];
        const int i = 4 * q;
        P1STEP(p1, t.x, kodd); P2STEP(p2, p1, i);     if (i >= 397 && i < 397 + WIN) hi[i - 397] = p2;
        P1STEP(p1, t.y, key0); P2STEP(p2, p1, i + 1); if (i + 1 < 397 + WIN) hi[i + 1 - 397] = p2;
        P1STEP(p1, t.z, kodd); P2STEP(p2, p1, i + 2); if (i + 2 < 397 + WIN) hi[i + 2 - 397] = p2;
        P1STEP(p1, t.w, key0); P2STEP(p2, p1, i + 3); if (i + 3 < 397 + WIN) hi[i + 3 - 397] = p2;
    }
#pragma unroll 4
    for (int q = 110; q < 156; q++) {   // words 440 .. 623
        t = T4[q];
        const int i = 4 * q;
        P1STEP(p1, t.x, kodd); P2STEP(p2, p1, i);
        P1STEP(p1, t.y, key0); P2STEP(p2, p1, i + 1);
        P1STEP(p1, t.z, kodd); P2STEP(p2, p1, i + 2);
        P1STEP(p1, t.w, key0); P2STEP(p2, p1, i + 3);
    }
    lo[1] = (m1w ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - 1u;
    lo[0] = 0x80000000u;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_v1(const long long* seeds, long long n, unsigned long long* out) {
    __shared__ uint4 T4[156];
    for (int i = threadIdx.x; i < 624; i += blockDim.x) ((uint32_t*)T4)[i] = c_init[i];
    __syncthreads();
    uint32_t lo[WIN + 2], hi[WIN];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        unsigned long long a = (unsigned long long)seeds[g];
        seed_v1(T4, (uint32_t)a, (uint32_t)(a >> 32), lo, hi);
        out[g] = shuffle16(lo, hi);
    }
}


// ---------------------------------------------------------------- V3: V2 + a shuffle whose draw loop is uniform over the warp
// The reference's loop is "for each position: draw until accepted"; run as written, a warp pays the worst lane's
// rejections at EVERY position and the lanes read different outputs.  Same draws, other loop order: every lane
// looks at output kk in the same iteration and either accepts it for its current position or not.
__device__ __forceinline__ unsigned long long shuffle16_uniform(const uint32_t* lo, const uint32_t* hi) {
    unsigned long long perm = 0xFEDCBA9876543210ull;
    int i = 15;
    uint32_t a = lo[0];
    for (int kk = 0; kk < WIN && i >= 1; kk++) {
        const uint32_t b = lo[kk + 1];
        uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
        y = hi[kk] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        a = b;
        const uint32_t nn = (uint32_t)i + 1u;
        const uint32_t r = temper(y) >> __clz(nn);          // 32 - bit_length(nn) = clz(nn)
        if (r < nn) {
            const unsigned long long d = ((perm >> (4 * i)) ^ (perm >> (4 * r))) & 0xFull;
            perm ^= (d << (4 * i)) | (d << (4 * r));
            i--;
        }
    }
    return i >= 1 ? 0ull : perm;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_v3(const long long* seeds, long long n, unsigned long long* out) {
    __shared__ uint4 T4[156];
    for (int i = threadIdx.x; i < 624; i += blockDim.x) ((uint32_t*)T4)[i] = c_init[i];
    __syncthreads();
    uint32_t lo[WIN + 2], hi[WIN];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        unsigned long long a = (unsigned long long)seeds[g];
        seed_v1(T4, (uint32_t)a, (uint32_t)(a >> 32), lo, hi);
        out[g] = shuffle16_uniform(lo, hi);
    }
}

// ---------------------------------------------------------------- V4: V3 + the pass-1 shift moved to the multiplier pipe
// x >> 30 == mulhi(x, 4); `four` comes from a kernel argument so that ptxas keeps the IMAD.HI.
#define P1STEP_H(x, t, kw) (x) = ((t) ^ (((x) ^ __umulhi((x), four)) * 1664525u)) + (kw)
__device__ __forceinline__ void seed_v4(const uint4* __restrict__ T4, uint32_t key0, uint32_t key1, uint32_t* lo, uint32_t* hi, uint32_t four) {
    const bool two = key1 != 0u;
    const uint32_t kodd = two ? key1 + 1u : key0;
    uint4 t = T4[0];
    uint32_t prev = (t.y ^ ((t.x ^ (t.x >> 30)) * 1664525u)) + key0;
    const uint32_t first1 = prev;
    P1STEP(prev, t.z, kodd);
    P1STEP_H(prev, t.w, key0);
#pragma unroll 4
    for (int q = 1; q < 156; q++) {
        t = T4[q];
        P1STEP(prev, t.x, kodd); P1STEP_H(prev, t.y, key0); P1STEP(prev, t.z, kodd); P1STEP_H(prev, t.w, key0);
    }
    const uint32_t m1w = (first1 ^ ((prev ^ (prev >> 30)) * 1664525u)) + kodd;
    uint32_t p1 = first1, p2 = m1w;
    t = T4[0];
    P1STEP_H(p1, t.z, kodd); P2STEP(p2, p1, 2); lo[2] = p2;
    P1STEP_H(p1, t.w, key0); P2STEP(p2, p1, 3); lo[3] = p2;
#pragma unroll 1
    for (int q = 1; q < (WIN + 4) / 4; q++) {
        t = T4[q];
        const int i = 4 * q;
        P1STEP_H(p1, t.x, kodd); P2STEP(p2, p1, i);     lo[i] = p2;
        P1STEP_H(p1, t.y, key0); P2STEP(p2, p1, i + 1); lo[i + 1] = p2;
        P1STEP_H(p1, t.z, kodd); P2STEP(p2, p1, i + 2); if (i + 2 <= WIN + 1) lo[i + 2] = p2;
        P1STEP_H(p1, t.w, key0); P2STEP(p2, p1, i + 3); if (i + 3 <= WIN + 1) lo[i + 3] = p2;
    }
#pragma unroll 4
    for (int q = (WIN + 4) / 4; q < 99; q++) {
        t = T4[q];
        const int i = 4 * q;
        P1STEP_H(p1, t.x, kodd); P2STEP(p2, p1, i);
        P1STEP_H(p1, t.y, key0); P2STEP(p2, p1, i + 1);
        P1STEP_H(p1, t.z, kodd); P2STEP(p2, p1, i + 2);
        P1STEP_H(p1, t.w, key0); P2STEP(p2, p1, i + 3);
    }
#pragma unroll 1
    for (int q = 99; q < 110; q++) {
        t = T4[q];
        const int i = 4 * q;
        P1STEP_H(p1, t.x, kodd); P2STEP(p2, p1, i);     if (i >= 397 && i < 397 + WIN) hi[i - 397] = p2;
        P1STEP_H(p1, t.y, key0); P2STEP(p2, p1, i + 1); if (i + 1 < 397 + WIN) hi[i + 1 - 397] = p2;
        P1STEP_H(p1, t.z, kodd); P2STEP(p2, p1, i + 2); if (i + 2 < 397 + WIN) hi[i + 2 - 397] = p2;
        P1STEP_H(p1, t.w, key0); P2STEP(p2, p1, i + 3); if (i + 3 < 397 + WIN) hi[i + 3 - 397] = p2;
    }
#pragma unroll 4
    for (int q = 110; q < 156; q++) {
        t = T4[q];
        const int i = 4 * q;
        P1STEP_H(p1, t.x, kodd); P2STEP(p2, p1, i);
        P1STEP_H(p1, t.y, key0); P2STEP(p2, p1, i + 1);
        P1STEP_H(p1, t.z, kodd); P2STEP(p2, p1, i + 2);
        P1STEP_H(p1, t.w, key0); P2STEP(p2, p1, i + 3);
    }
    lo[1] = (m1w ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - 1u;
    lo[0] = 0x80000000u;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_v4(const long long* seeds, long long n, unsigned long long* out, uint32_t four) {
    __shared__ uint4 T4[156];
    for (int i = threadIdx.x; i < 624; i += blockDim.x) ((uint32_t*)T4)[i] = c_init[i];
    __syncthreads();
    uint32_t lo[WIN + 2], hi[WIN];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        unsigned long long a = (unsigned long long)seeds[g];
        seed_v4(T4, (uint32_t)a, (uint32_t)(a >> 32), lo, hi, four);
        out[g] = shuffle16_uniform(lo, hi);
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main() {
    uint32_t t[624];
    t[0] = 19650218u;
    for (int i = 1; i < 624; i++) t[i] = 1812433253u * (t[i - 1] ^ (t[i - 1] >> 30)) + (uint32_t)i;
    CK(cudaMemcpyToSymbol(c_init, t, sizeof(t)));
    const long long n = 1000000;
    std::vector<long long> hs(n);
    for (long long i = 0; i < n; i++) hs[i] = i + 1;
    hs[10] = (1ll << 33) + 7;
    long long* ds; unsigned long long *o0, *o1;
    CK(cudaMalloc(&ds, 8 * n)); CK(cudaMalloc(&o0, 8 * n)); CK(cudaMalloc(&o1, 8 * n));
    CK(cudaMemcpy(ds, hs.data(), 8 * n, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char* name, auto launch) {
        for (int i = 0; i < 2; i++) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 10; i++) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaGetLastError());
        printf("%-28s %.4f ms per 1 M seeds\n", name, ms / 10);
    };
    for (int ctas : {6, 8, 12, 16, 32}) {
        char nm[64];
        snprintf(nm, 64, "V0 const/LDC grid 148x%d", ctas);
        timeit(nm, [&] { k_v0<<<148 * ctas, 256>>>(ds, n, o0); });
    }
    for (int ctas : {6, 8, 12, 16, 32}) {
        char nm[64];
        snprintf(nm, 64, "V1 smem grid 148x%d", ctas);
        timeit(nm, [&] { k_v1<1><<<148 * ctas, 256>>>(ds, n, o1); });
    }
    for (int ctas : {8, 16, 32}) {
        char nm[64];
        snprintf(nm, 64, "V2 smem 32-reg grid 148x%d", ctas);
        timeit(nm, [&] { k_v1<8><<<148 * ctas, 256>>>(ds, n, o1); });
    }
    timeit("V1 one thread per seed", [&] { k_v1<1><<<(n + 255) / 256, 256>>>(ds, n, o1); });
    timeit("V2 one thread per seed", [&] { k_v1<8><<<(n + 255) / 256, 256>>>(ds, n, o1); });
    unsigned long long *o3, *o4;
    CK(cudaMalloc(&o3, 8 * n)); CK(cudaMalloc(&o4, 8 * n));
    for (int ctas : {8, 16}) {
        char nm[64];
        snprintf(nm, 64, "V3 uniform shuffle 148x%d", ctas);
        timeit(nm, [&] { k_v3<8><<<148 * ctas, 256>>>(ds, n, o3); });
        snprintf(nm, 64, "V3 (min 1 CTA) 148x%d", ctas);
        timeit(nm, [&] { k_v3<1><<<148 * ctas, 256>>>(ds, n, o3); });
        snprintf(nm, 64, "V4 + mulhi 148x%d", ctas);
        timeit(nm, [&] { k_v4<8><<<148 * ctas, 256>>>(ds, n, o4, 4u); });
        snprintf(nm, 64, "V4 (min 1 CTA) 148x%d", ctas);
        timeit(nm, [&] { k_v4<1><<<148 * ctas, 256>>>(ds, n, o4, 4u); });
    }
    std::vector<unsigned long long> h0(n), h1(n), h3(n), h4(n);
    CK(cudaMemcpy(h3.data(), o3, 8 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h4.data(), o4, 8 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h0.data(), o0, 8 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1.data(), o1, 8 * n, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (long long i = 0; i < n; i++) bad += (h0[i] != h1[i]) + (h0[i] != h3[i]) + (h0[i] != h4[i]);
    printf("mismatches V0 vs V1..V4: %lld; seed 42 deck %016llx (expect hand order of the reference deal)\n", bad, h0[41]);
    return bad != 0;
}
