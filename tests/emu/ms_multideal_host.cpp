// tests/emu/ms_multideal_host.cpp -- the PRODUCT's multi-deal sampled-CFR kernels (scopa_b200/csrc/ms_multideal.cu:
// md_mccfr_kernel<768> with the lock-free open-addressing infoset table, md_apply_kernel, md_export_kernel,
// md_lookup_kernel, md_build_kernel + md_blocked_kernel) executed on the host by the CTA emulator of
// tests/emu/cta_emu.h: one pthread per CUDA thread, the "HBM" table in host memory, atomicCAS / atomicAdd / atomicOr as
// real atomics, __ldcg as a plain load, IEEE double without contraction.  The host side of the library (allocation,
// deal of the roots, launches) is the few lines of ms_md_create / ms_md_mccfr_batch / ms_md_apply restated below;
// the dealt roots are passed in by the test.  Test infrastructure.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <vector>

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __activemask() { return 1u; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return v; }   // only ever a shared loop bound
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline long long __double_as_longlong(double d) { long long v; __builtin_memcpy(&v, &d, 8); return v; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
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
        __builtin_memcpy(&cur, &old, 8);
        cur += v;
        __builtin_memcpy(&want, &cur, 8);
    } while (!__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    __builtin_memcpy(&cur, &old, 8);
    return cur;
}
#undef __device__
#undef __global__
#undef __host__
#undef __shared__
#undef __constant__
#undef __forceinline__
#undef __launch_bounds__
#undef __align__
#include "cta_emu.h"
#define __host__
#define __constant__
#define __align__(n) alignas(n)
#define md_smem emu_dyn_smem[emu_block_slot]
static inline void __syncwarp() { __syncthreads(); }
// full-mask shuffles of the final counter reductions (every thread of the block executes them equally often)
static unsigned long long emu_shfl_buf[EMU_MAX_CLUSTER][2048];
static inline unsigned long long __shfl_down_sync(unsigned, unsigned long long v, int off) {
    unsigned long long* b = emu_shfl_buf[emu_block_slot];
    const unsigned t = threadIdx.x, lane = t & 31u;
    b[t] = v;
    __syncthreads();
    const unsigned long long r = (lane + (unsigned)off < 32u && t + (unsigned)off < blockDim.x) ? b[t + off] : v;
    __syncthreads();
    return r;
}

#define MS_HOST_RULES_ONLY
#include "../../scopa_b200/csrc/ms_multideal.cu"

namespace ms {   // declared in ms_common.cuh, defined in ms_env.cu in the library
std::atomic<uint64_t> g_launches{0};
char* last_error_buf() { static thread_local char buf[512]; return buf; }
}

namespace {
using namespace ms;
struct HostMd {
    MdDev dev{};
    int log2cap = 0;
    std::vector<uint4> roots; std::vector<uint32_t> hand_order;
    std::vector<unsigned long long> counters;
    std::vector<unsigned int> dirty;
    MdSlot* slots = nullptr;
    std::vector<MdDealInfo> info; bool have_info = false;
} M;

unsigned grid_of(long long n, int block, int per_sm) {
    long long need = (n + block - 1) / block, cap = 148LL * per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

struct BatchArgs { MdDev t; int player; long long n_trav; uint2 key; unsigned long long first_trav; };
void batch_entry(BatchArgs a) { md_mccfr_kernel<MD_THREADS>(a.t, a.player, a.n_trav, a.key, a.first_trav); }
void apply_entry(MdDev t) { md_apply_kernel(t); }
struct ExportArgs { MdDev t; unsigned long long* keys; double* reg; double* strat; long long max_n; unsigned long long* n_out; };
void export_entry(ExportArgs a) { md_export_kernel(a.t, a.keys, a.reg, a.strat, a.max_n, a.n_out); }
struct LookupArgs { MdDev t; const unsigned long long* keys; long long n; double* reg; double* strat; uint8_t* found; };
struct BuildArgs { MdDev t; MdDealInfo* info; };
void build_entry(BuildArgs a) { md_build_kernel(a.t, a.info); }
struct BlockedArgs { MdDev t; const MdDealInfo* info; int player; unsigned long long first_visit; long long n_visits; int pairs; uint2 key; };
void blocked_entry(BlockedArgs a) { md_blocked_kernel(a.t, a.info, a.player, a.first_visit, a.n_visits, a.pairs, a.key); }
void lookup_entry(LookupArgs a) { md_lookup_kernel(a.t, a.keys, a.n, a.reg, a.strat, a.found); }
}  // namespace

extern "C" {

// ms_md_create + ms_md_reset with the roots already dealt (states [n][4] u32, hand_order [n])
int host_md_create(const uint32_t* roots4, const uint32_t* hand_order, long long n_deals, int log2_capacity) {
    if (M.slots) std::free(M.slots);
    M = HostMd();
    M.log2cap = log2_capacity;
    const size_t cap = (size_t)1 << log2_capacity;
    M.slots = (MdSlot*)std::aligned_alloc(128, cap * sizeof(MdSlot));
    if (!M.slots) return -1;
    std::memset(M.slots, 0, cap * sizeof(MdSlot));
    M.dirty.assign(cap >> 5, 0u);
    M.counters.assign(8, 0ull);
    for (long long i = 0; i < n_deals; i++) {
        M.roots.push_back(make_uint4(roots4[4 * i], roots4[4 * i + 1], roots4[4 * i + 2], roots4[4 * i + 3]));
        M.hand_order.push_back(hand_order[i]);
    }
    M.dev.slots = M.slots; M.dev.mask = (unsigned long long)cap - 1ull; M.dev.shift = 64 - log2_capacity;
    M.dev.max_probe = cap < 8192 ? (unsigned int)cap : 8192u;
    M.dev.roots = M.roots.data(); M.dev.hand_order = M.hand_order.data(); M.dev.n_deals = (unsigned int)n_deals;
    M.dev.dirty = M.dirty.data(); M.dev.counters = M.counters.data();
    return 0;
}

// ms_md_mccfr_batch: md_mccfr_kernel<768><<<grid_for(n_trav, 768, 1), 768, MdFrames<768>::kBytes>>>
int host_md_batch(int player, long long n_trav, unsigned long long philox_seed, unsigned long long first_trav) {
    if ((size_t)MdFrames<MD_THREADS>::kBytes > EMU_SMEM_BYTES) return -4;
    BatchArgs a{M.dev, player, n_trav, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)), first_trav};
    return emu_launch_grid(batch_entry, a, grid_of(n_trav, MD_THREADS, 1), MD_THREADS);
}

// ms_md_mccfr_blocked: md_build_kernel once per table (describes every deal's tree, claims the slots), then
// md_blocked_kernel<<<min(n_visits, 148), MDB_THREADS, mdb_smem_bytes()>>>
int host_md_blocked(int player, long long first_visit, long long n_visits, int pairs_per_visit, unsigned long long philox_seed) {
    if (!M.have_info) {
        M.info.resize(M.dev.n_deals);
        BuildArgs b{M.dev, M.info.data()};
        if (emu_launch_grid(build_entry, b, grid_of((long long)M.dev.n_deals * 256, 256, 4), 256)) return -1;
        M.have_info = true;
    }
    if (mdb_smem_bytes() > EMU_SMEM_BYTES) return -4;
    BlockedArgs a{M.dev, M.info.data(), player, (unsigned long long)first_visit, n_visits, pairs_per_visit,
                  make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32))};
    return emu_launch_grid(blocked_entry, a, (unsigned)(n_visits < 148 ? n_visits : 148), MDB_THREADS);
}

int host_md_apply() { return emu_launch_grid(apply_entry, M.dev, 4, 256); }

void host_md_counters(unsigned long long* out5, int reset) {
    for (int i = 0; i < 5; i++) out5[i] = M.counters[i];
    if (reset) for (int i = 0; i < 3; i++) M.counters[i] = 0;
}

long long host_md_export(unsigned long long* keys, double* reg, double* strat, long long max_n) {
    unsigned long long n = 0;
    ExportArgs a{M.dev, keys, reg, strat, max_n, &n};
    if (emu_launch_grid(export_entry, a, 4, 256)) return -1;
    return (long long)n;
}

int host_md_lookup(const unsigned long long* keys, long long n, double* reg, double* strat, uint8_t* found) {
    LookupArgs a{M.dev, keys, n, reg, strat, found};
    return emu_launch_grid(lookup_entry, a, 2, 256);
}

}  // extern "C"
