// tests/emu/ms_multideal_host.cpp -- the PRODUCT's multi-deal sampled-CFR kernels (scopa_b200/csrc/ms_multideal.cu:
// md_mccfr_kernel<768> with the lock-free open-addressing infoset table, md_apply_kernel, md_export_kernel,
// md_lookup_kernel, md_build_kernel + md_blocked_kernel) executed on the host by the CTA emulator of
// tests/emu/cta_emu.h: one pthread per CUDA thread, the "HBM" table in host memory, atomicCAS / atomicAdd / atomicOr as
// real atomics, __ldcg as a plain load, IEEE double without contraction.  The host side of the library (allocation,
// deal of the roots, launches) is the few lines of ms_md_create / ms_md_mccfr_batch / ms_md_apply restated below;
// the dealt roots are passed in by the test.  host_md_world_create emulates several RANKS sharing one sharded table (what
// ms_md_ipc_attach sets up over CUDA IPC), host_md_barrier_all runs md_peer_barrier_kernel for all of them concurrently.
// Test infrastructure.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <vector>

#include "host_intrinsics.h"
#include "cta_emu.h"
#define md_smem emu_dyn_smem[emu_block_slot]
static inline void __syncwarp() { __syncthreads(); }
#include "cta_emu_warp.h"

// the barrier's memory operations (ms_multideal.cu defines the PTX forms for the device)
#include <chrono>
static inline void md_peer_signal(unsigned long long* flag, unsigned long long epoch) { __atomic_store_n(flag, epoch, __ATOMIC_RELEASE); }
static inline unsigned long long md_peer_poll(const unsigned long long* flag) { return __atomic_load_n(flag, __ATOMIC_ACQUIRE); }
static inline unsigned long long md_peer_clock_ns() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline void md_peer_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return cmp;
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
    std::vector<unsigned long long> flags;      // barrier flags of this "rank"
    unsigned int peer_err = 0;
};
// Several emulated RANKS in one process (host_md_world_create): each owns a shard of the table, a dirty bitmap, counters
// and deal descriptions, and sees the others' shards through MdDev::peer_slots / peer_dirty -- exactly what CUDA IPC
// gives the device build.  host_md_select picks the rank the calls below act on.
HostMd RANKS[MD_MAX_PEERS];
int CUR = 0, WORLD = 1;
unsigned long long EPOCH = 0;
#define M RANKS[CUR]

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
static int create_rank(const uint32_t* roots4, const uint32_t* hand_order, long long n_deals, int log2_capacity);

int host_md_create(const uint32_t* roots4, const uint32_t* hand_order, long long n_deals, int log2_capacity) {
    for (int r = 0; r < MD_MAX_PEERS; r++) { if (RANKS[r].slots) std::free(RANKS[r].slots); RANKS[r] = HostMd(); }
    CUR = 0; WORLD = 1; EPOCH = 0;
    return create_rank(roots4, hand_order, n_deals, log2_capacity);
}

// `world` ranks with the same deals, every shard 2^log2_capacity slots, wired to each other like ms_md_ipc_attach does
int host_md_world_create(const uint32_t* roots4, const uint32_t* hand_order, long long n_deals, int log2_capacity, int world) {
    if (world < 1 || world > MD_MAX_PEERS) return -2;
    for (int r = 0; r < MD_MAX_PEERS; r++) { if (RANKS[r].slots) std::free(RANKS[r].slots); RANKS[r] = HostMd(); }
    WORLD = world; EPOCH = 0;
    for (int r = 0; r < world; r++) { CUR = r; if (create_rank(roots4, hand_order, n_deals, log2_capacity)) return -1; }
    for (int r = 0; r < world; r++) {
        RANKS[r].dev.world = world; RANKS[r].dev.rank = r;
        for (int q = 0; q < world; q++) { RANKS[r].dev.peer_slots[q] = RANKS[q].slots; RANKS[r].dev.peer_dirty[q] = RANKS[q].dirty.data(); }
    }
    CUR = 0;
    return 0;
}

int host_md_select(int rank) { if (rank < 0 || rank >= WORLD) return -1; CUR = rank; return 0; }

static int create_rank(const uint32_t* roots4, const uint32_t* hand_order, long long n_deals, int log2_capacity) {
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
    M.dev.world = 1; M.dev.rank = 0; M.dev.lshift = log2_capacity;
    M.dev.peer_slots[0] = M.slots; M.dev.peer_dirty[0] = M.dirty.data();
    M.flags.assign(MD_MAX_PEERS, 0ull);
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

// ms_md_peer_barrier on every rank at once: md_peer_barrier_kernel, one block per emulated rank, running concurrently.
// `absent` >= 0: that rank does not show up (its block returns at once) -- the others must time out and report it.
struct BarrierArgs { MdPeerSync ps[MD_MAX_PEERS]; unsigned long long epoch; unsigned int* err[MD_MAX_PEERS]; int absent; };
static void barrier_entry(BarrierArgs a) {
    if ((int)blockIdx.x == a.absent) return;
    md_peer_barrier_kernel(a.ps[blockIdx.x], a.epoch, a.err[blockIdx.x]);
}
int host_md_barrier_all(int absent) {
    BarrierArgs a{};
    a.epoch = ++EPOCH; a.absent = absent;
    for (int r = 0; r < WORLD; r++) {
        for (int q = 0; q < WORLD; q++) a.ps[r].flags[q] = RANKS[q].flags.data();
        a.ps[r].my_flags = RANKS[r].flags.data(); a.ps[r].rank = r; a.ps[r].world = WORLD;
        a.err[r] = &RANKS[r].peer_err;
    }
    return emu_launch_cluster(barrier_entry, a, (unsigned)WORLD, 32);
}
unsigned host_md_peer_error(int rank) { return RANKS[rank].peer_err; }

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
