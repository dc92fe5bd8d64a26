// tests/emu/ms_solver_host.cpp -- the PRODUCT's game-tree enumeration and vanilla-CFR kernels (scopa_b200/csrc/
// ms_solver.cu: tree_expand_kernel <<<1, 256>>>, cfr_kernel <<<1, 512>>> with cfr_run / cfr_traversal, and regret_match of
// ms_tree_walk.cuh) executed on the host by the CTA emulator of tests/emu/cta_emu.h: one pthread per CUDA thread,
// pthread barriers for __syncthreads, one buffer for the dynamic shared memory.  __dadd_rn / __dmul_rn / __ddiv_rn are
// IEEE double add / mul / div (this file is compiled with -ffp-contract=off), so the emulated kernel must produce the
// same float64 bits as the device -- and as numpy in the reference (tests/test_solver_host.py checks the latter against
// tables recorded from the unmodified reference).
// Between the two kernels the library indexes the infosets on the host (solver_build in ms_solver.cu: "pure
// bookkeeping, no game rules"); that step is restated here from its description: slots in breadth-first
// first-occurrence order, per-slot chains of node ids in ascending order, slots grouped by tree level.
// Test infrastructure.
#include <cstdint>
#include <cuda_runtime.h>   // vector types for the host compiler
#include <unordered_map>
#include <vector>

#include "host_intrinsics.h"
#include "cta_emu.h"
static inline void __syncwarp() { __syncthreads(); }   // only used by the one-warp kernels (<<<1, 32>>>): the warp is the block
#include "cta_emu_warp.h"

// the peer exchange's two memory operations + clock + load (device: PTX in ms_solver.cu), for two emulated ranks in one process
#include <atomic>
#include <chrono>
static inline void peer_signal(unsigned long long* flag, unsigned long long epoch) { __atomic_store_n(flag, epoch, __ATOMIC_RELEASE); }
static inline unsigned long long peer_poll(const unsigned long long* flag) { return __atomic_load_n(flag, __ATOMIC_ACQUIRE); }
static inline unsigned long long peer_clock_ns() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline double peer_load(const double* p) { return *(const volatile double*)p; }
static inline void peer_store(double* p, double v) { *(volatile double*)p = v; }
static inline void peer_fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> static inline T __ldcg_host(const T* p) { return *p; }
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
    return cmp;
}

#define MS_HOST_RULES_ONLY
#include "../../scopa_b200/csrc/ms_solver.cu"

namespace ms {   // declared in ms_common.cuh, defined in ms_env.cu in the library
std::atomic<uint64_t> g_launches{0};
char* last_error_buf() { static thread_local char buf[512]; return buf; }
}

namespace {
using namespace ms;

struct ExpandArgs { uint4 root; uint32_t hand_order; TreeOut t; };
void expand_entry(ExpandArgs a) { tree_expand_kernel(a.root, a.hand_order, a.t); }
struct CfrArgs { SolverDev d; int n_dec, iters, only_player; double r0, r1; double* out; };
void cfr_entry(CfrArgs a) { cfr_kernel(a.d, a.n_dec, a.iters, a.only_player, a.r0, a.r1, a.out); }
struct InplaceArgs { SolverDev d; long long iters; uint2 key; unsigned long long first_iter; int nframes; };
void inplace_entry(InplaceArgs a) { mccfr_inplace_kernel(a.d, a.iters, a.key, a.first_iter, a.nframes); }
void inplace_tree_entry(InplaceArgs a) { mccfr_inplace_tree_kernel(a.d, a.iters, a.key, a.first_iter, a.nframes); }
struct BatchArgs { SolverDev d; int player; long long n_trav; uint2 key; unsigned long long first_trav; int nframes, ncopy; };
struct StaticArgs { SolverDev d; int player; long long n_trav; uint2 key; unsigned long long first_trav; StaticDims dm; };
void static_entry(StaticArgs a) { mccfr_static_kernel(a.d, a.player, a.n_trav, a.key, a.first_trav, a.dm); }
struct ManyRunArgs { SolverDev d; ManyRuns m; long long iters; unsigned long long seed0, first_iter; int nframes, warps; };
void inplace_many_entry(ManyRunArgs a) { mccfr_inplace_many_kernel(a.d, a.m, a.iters, a.seed0, a.first_iter, a.nframes, a.warps); }
struct PeersArgs { SolverDev d[2]; PeerView pv[2]; unsigned long long epoch; unsigned int* err[2]; };
void peers_entry(PeersArgs a) { const unsigned r = emu_block_slot; mccfr_apply_peers_kernel(a.d[r], a.pv[r], a.epoch, a.err[r]); }
struct FusedArgs { SolverDev d[2]; PeerView pv[2]; unsigned int* err[2]; unsigned int* ticket[2]; unsigned long long epoch, first_trav[2];
                   long long n_trav; uint2 key; StaticDims dm; int force_rank; };
void fused_entry(FusedArgs a) {
    const unsigned r = a.force_rank >= 0 ? (unsigned)a.force_rank : emu_block_slot;
    mccfr_static_peers_kernel(a.d[r], 2, a.n_trav, a.key, a.first_trav[r], a.dm, a.pv[r], a.epoch, a.err[r], a.ticket[r]);
}
void tree_entry(BatchArgs a) { mccfr_tree_kernel<TREE_THREADS>(a.d, a.player, a.n_trav, a.key, a.first_trav, a.nframes, a.ncopy); }
void restep_entry(BatchArgs a) { mccfr_batch_kernel(a.d, a.player, a.n_trav, a.key, a.first_trav, a.nframes); }
void es_tree_entry(BatchArgs a) { mccfr_es_tree_kernel<TREE_THREADS>(a.d, a.player, a.n_trav, a.key, a.first_trav, a.nframes, a.ncopy); }
void os_entry(BatchArgs a) { mccfr_os_kernel(a.d, a.player, a.n_trav, a.key, a.first_trav); }
void apply_entry(SolverDev d) { mccfr_apply_kernel(d); }
struct ManyArgs { const CfrJob* jobs; int iters; };
void cfr_many_entry(ManyArgs a) { cfr_many_kernel(a.jobs, a.iters); }
struct BrArgs { SolverDev d; int n_dec, kind; double* out2; };
void br_entry(BrArgs a) { best_response_kernel(a.d, a.n_dec, a.kind, a.out2); }
struct PolArgs { SolverDev d; int kind; double* out; };
void policy_entry(PolArgs a) { policy_kernel(a.d, a.kind, a.out); }
struct EvalArgs { SolverDev d; const double* pol0; const double* pol1; long long n; uint2 key; unsigned long long first; float* rew; uchar2* scopas; };
void eval_entry(EvalArgs a) { eval_kernel(a.d, a.pol0, a.pol1, a.n, a.key, a.first, a.rew, a.scopas); }

struct HostSolver {
    std::vector<uint4> state; std::vector<int> parent, child_begin32, level_begin, slot_level_begin;
    std::vector<uint8_t> nchild, slot_nlegal, slot_player, slot_legal, touched;
    std::vector<unsigned long long> key, slot_key, counters;
    std::vector<uint16_t> legal, child_begin, chain_begin, chain_nodes;
    std::vector<int8_t> rx2;
    std::vector<int16_t> node_slot;
    std::vector<double> regret, strategy, delta;
    std::vector<unsigned long long> hkeys; std::vector<int16_t> hslots;
    int N = 0, L = 0, S = 0, n_dec = 0, hcap = 0, nframes = 1, nframes_tree = 1, nframes_es = 1;
    SolverDev dev{};
} H;
}  // namespace

extern "C" {

int host_solver_build(const uint32_t* root4, uint32_t hand_order, int* n_nodes, int* n_slots, int* n_levels) {
    H = HostSolver();
    H.state.resize(MAXN); H.parent.resize(MAXN); H.child_begin32.resize(MAXN); H.nchild.resize(MAXN); H.key.resize(MAXN);
    H.legal.resize(MAXN); H.rx2.resize(MAXN); H.level_begin.assign(MAXL + 2, 0);
    int counts[4] = {0, 0, 0, 0};
    ExpandArgs ea;
    ea.root = make_uint4(root4[0], root4[1], root4[2], root4[3]);
    ea.hand_order = hand_order;
    ea.t.state = H.state.data(); ea.t.parent = H.parent.data(); ea.t.child_begin = H.child_begin32.data();
    ea.t.nchild = H.nchild.data(); ea.t.key = H.key.data(); ea.t.legal = H.legal.data(); ea.t.rx2 = H.rx2.data();
    ea.t.level_begin = H.level_begin.data(); ea.t.counts = counts;
    if (emu_launch_cta(expand_entry, ea, 256)) return -1;
    if (counts[2]) return -2;
    const int N = counts[0], L = counts[1];
    H.N = N; H.L = L;
    // infoset slots: breadth-first first occurrence; chains: the slot's nodes in ascending index order
    std::vector<int> level(N, 0), slot_level;
    for (int l = 0; l < L; l++) for (int v = H.level_begin[l]; v < H.level_begin[l + 1]; v++) level[v] = l;
    std::unordered_map<unsigned long long, int> slot_of;
    std::vector<std::vector<int>> chains;
    H.node_slot.assign(N, -1);
    for (int v = 0; v < N; v++) {
        if (!H.nchild[v]) continue;
        auto it = slot_of.find(H.key[v]);
        int s;
        if (it == slot_of.end()) {
            s = (int)chains.size();
            slot_of.emplace(H.key[v], s);
            chains.emplace_back();
            slot_level.push_back(level[v]);
            H.slot_key.push_back(H.key[v]);
            H.slot_nlegal.push_back(H.nchild[v]);
            H.slot_player.push_back((uint8_t)((H.key[v] >> 52) & 1ull));
            for (int i = 0; i < 4; i++) H.slot_legal.push_back(i < H.nchild[v] ? (uint8_t)((H.legal[v] >> (4 * i)) & 0xF) : (uint8_t)0xFF);
        } else {
            s = it->second;
            if (slot_level[s] != level[v] || H.slot_nlegal[s] != H.nchild[v]) return -3;
        }
        H.node_slot[v] = (int16_t)s;
        chains[s].push_back(v);
    }
    const int S = (int)chains.size();
    H.S = S;
    H.slot_level_begin.assign(L + 1, S);
    for (int l = 0, s = 0; l <= L; l++) {
        while (s < S && slot_level[s] < l) s++;
        H.slot_level_begin[l] = s;
    }
    for (int s = 0; s < S; s++) {
        H.chain_begin.push_back((uint16_t)H.chain_nodes.size());
        for (int v : chains[s]) H.chain_nodes.push_back((uint16_t)v);
    }
    H.chain_begin.push_back((uint16_t)H.chain_nodes.size());
    H.n_dec = (int)H.chain_nodes.size();
    H.child_begin.resize(N);
    for (int v = 0; v < N; v++) H.child_begin[v] = (uint16_t)H.child_begin32[v];
    H.regret.assign(4 * (size_t)S, 0.0); H.strategy.assign(4 * (size_t)S, 0.0); H.delta.assign(6 * (size_t)S, 0.0);
    H.touched.assign(S, 0); H.counters.assign(4, 0);
    SolverDev& d = H.dev;
    d.n_nodes = N; d.n_levels = L; d.n_slots = S; d.root_cur = (int)((root4[3] >> 17) & 1u);
    d.root = ea.root; d.hand_order = hand_order;
    d.level_begin = H.level_begin.data(); d.child_begin = H.child_begin.data(); d.nchild = H.nchild.data();
    d.node_slot = H.node_slot.data(); d.rx2 = H.rx2.data();
    d.chain_begin = H.chain_begin.data(); d.chain_nodes = H.chain_nodes.data(); d.slot_level_begin = H.slot_level_begin.data();
    d.slot_nlegal = H.slot_nlegal.data(); d.slot_player = H.slot_player.data();
    // key -> slot index probed by lookup_slot (open addressing, linear probing; a wrong restatement of the hash makes
    // every lookup miss), and the frames the sampled traversal needs: decision levels of one player
    int hcap = 1024;
    while (hcap < 2 * S + 2) hcap *= 2;
    H.hcap = hcap;
    H.hkeys.assign(hcap, 0xFFFFFFFFFFFFFFFFull); H.hslots.assign(hcap, -1);
    for (int s = 0; s < S; s++) {
        uint32_t h = (uint32_t)((H.slot_key[s] * 0x9E3779B97F4A7C15ull) >> 40) & (uint32_t)(hcap - 1);
        while (H.hkeys[h] != 0xFFFFFFFFFFFFFFFFull) h = (h + 1) & (uint32_t)(hcap - 1);
        H.hkeys[h] = H.slot_key[s]; H.hslots[h] = (int16_t)s;
    }
    int dl[2] = {0, 0};
    for (int l = 0; l < L; l++) {
        bool any = false;
        for (int v = H.level_begin[l]; v < H.level_begin[l + 1]; v++) any |= H.nchild[v] > 0;
        if (any) dl[(d.root_cur + l) & 1]++;
    }
    H.nframes = dl[0] > dl[1] ? dl[0] : dl[1];
    if (H.nframes < 1) H.nframes = 1;
    d.hkeys = H.hkeys.data(); d.hslots = H.hslots.data(); d.hcap = hcap;
    // frames of the tree-walking kernels (solver_build, step 2): the reference estimator pushes no frame at a traverser
    // node whose single move leads to the end of the game through forced moves only; external sampling pushes one at
    // every traverser node with more than one action.  Longest such chain of one player, bottom-up over the tree.
    {
        std::vector<int> a0(N, 0), a1(N, 0), e0(N, 0), e1(N, 0);
        for (int v = N - 1; v >= 0; v--) {
            const int nc = H.nchild[v];
            if (!nc) continue;
            const int cb = H.child_begin32[v];
            int m0 = 0, m1 = 0, x0 = 0, x1 = 0;
            for (int c = cb; c < cb + nc; c++) {
                if (a0[c] > m0) m0 = a0[c];
                if (a1[c] > m1) m1 = a1[c];
                if (e0[c] > x0) x0 = e0[c];
                if (e1[c] > x1) x1 = e1[c];
            }
            const bool forced = nc == 1 && (H.nchild[cb] == 0 || (H.nchild[cb] == 1 && H.nchild[H.child_begin32[cb]] == 0));
            const int p = (int)((H.state[v].w >> 17) & 1u), push = forced ? 0 : 1;
            a0[v] = m0 + (p == 0 ? push : 0); a1[v] = m1 + (p == 1 ? push : 0);
            e0[v] = x0 + ((p == 0 && nc > 1) ? 1 : 0); e1[v] = x1 + ((p == 1 && nc > 1) ? 1 : 0);
        }
        H.nframes_tree = std::max(1, std::max(a0[0], a1[0]));
        H.nframes_es = std::max(1, std::max(e0[0], e1[0]));
    }
    d.regret = H.regret.data(); d.strategy = H.strategy.data(); d.delta = H.delta.data();
    d.touched = H.touched.data(); d.counters = H.counters.data();
    *n_nodes = N; *n_slots = S; *n_levels = L;
    return 0;
}

void host_solver_tree(uint32_t* states, int* parent, int* child_begin, uint8_t* nchild, int* slot, int* level_begin) {
    for (int v = 0; v < H.N; v++) {
        states[4 * v] = H.state[v].x; states[4 * v + 1] = H.state[v].y; states[4 * v + 2] = H.state[v].z; states[4 * v + 3] = H.state[v].w;
        parent[v] = H.parent[v]; child_begin[v] = H.child_begin32[v]; nchild[v] = H.nchild[v]; slot[v] = H.node_slot[v];
    }
    for (int l = 0; l <= H.L; l++) level_begin[l] = H.level_begin[l];
}

void host_solver_export(unsigned long long* keys, uint8_t* nlegal, uint8_t* legal4, double* regret, double* strategy) {
    for (int s = 0; s < H.S; s++) {
        if (keys) keys[s] = H.slot_key[s];
        if (nlegal) nlegal[s] = H.slot_nlegal[s];
        for (int i = 0; i < 4; i++) {
            if (legal4) legal4[4 * s + i] = H.slot_legal[4 * s + i];
            if (regret) regret[4 * s + i] = H.regret[4 * s + i];
            if (strategy) strategy[4 * s + i] = H.strategy[4 * s + i];
        }
    }
}

// ms_cfr_iterate / ms_cfr_traverse: cfr_kernel<<<1, 512, cfr_smem_bytes(...)>>>
int host_cfr(int iters, int only_player, double r0, double r1, double* out_value) {
    if (cfr_smem_bytes(H.N, H.S, H.n_dec) > EMU_SMEM_BYTES) return -4;
    CfrArgs a{H.dev, H.n_dec, iters, only_player, r0, r1, out_value};
    return emu_launch_cta(cfr_entry, a, 512);
}

// ms_mccfr_inplace (re-stepping form): mccfr_inplace_kernel<<<1, 32, smem>>>
int host_mccfr_inplace(long long iters, unsigned long long philox_seed, unsigned long long first_iter) {
    const size_t smem = 64 * (size_t)H.S + 8 * (size_t)H.hcap + (size_t)H.nframes * 44 + 2 * (size_t)H.hcap + H.S + 64;
    if (smem > EMU_SMEM_BYTES) return -4;
    InplaceArgs a{H.dev, iters, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)), first_iter, H.nframes};
    return emu_launch_cta(inplace_entry, a, 32);
}

// ms_mccfr_inplace (default form): mccfr_inplace_tree_kernel<<<1, 32, smem>>>
int host_mccfr_inplace_tree(long long iters, unsigned long long philox_seed, unsigned long long first_iter) {
    const size_t smem = 64 * (size_t)H.S + (size_t)H.nframes_tree * 26 + 4 * (size_t)H.N + H.S + 64;
    if (smem > EMU_SMEM_BYTES) return -4;
    InplaceArgs a{H.dev, iters, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)), first_iter, H.nframes_tree};
    return emu_launch_cta(inplace_tree_entry, a, 32);
}

// does the enumerated tree have the shape of a fresh deal (solver_build's static_shape test, restated)?
static bool host_static_dims(StaticDims& dm) {
    if (H.L != STATIC_PLIES + 1 || H.dev.root_cur != 0) return false;
    for (int l = 0; l < H.L; l++)
        for (int v = H.level_begin[l]; v < H.level_begin[l + 1]; v++) {
            const int want = l < STATIC_PLIES ? 4 - l / 2 : 0;
            if (H.nchild[v] != want || (want && (int)((H.state[v].w >> 17) & 1u) != (l & 1))) return false;
        }
    dm = static_dims_from(H.level_begin.data(), H.slot_level_begin.data());
    return mccfr_static_smem(H.S, dm) + 16 <= 227 * 1024;
}

// ms_mccfr_batch_mode: 0 = mccfr_static_kernel on a fresh deal's tree (the headline kernel; mccfr_tree_kernel otherwise),
// 4 = mccfr_tree_kernel, 3 = mccfr_batch_kernel (re-stepping), 1 = external sampling on the tree, 2 = outcome sampling;
// grids as grid_for(n_trav, threads, 1), blocks one after another
int host_mccfr_batch(int mode, int player, long long n_trav, unsigned long long philox_seed, unsigned long long first_trav) {
    BatchArgs a{H.dev, player, n_trav, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)), first_trav, 0, 1};
    auto grid = [&](int threads) { long long g = (n_trav + threads - 1) / threads; return (unsigned)(g < 1 ? 1 : (g > 148 ? 148 : g)); };
    StaticDims dm{};
    if (mode == 0 && host_static_dims(dm)) {
        StaticArgs sa{H.dev, player, n_trav, a.key, first_trav, dm};
        return emu_launch_grid(static_entry, sa, grid(STATIC_THREADS), STATIC_THREADS);
    }
    if (mode == 0 || mode == 4 || mode == 1) {
        a.nframes = mode != 1 ? H.nframes_tree : H.nframes_es;
        for (int ncopy : {4, 2, 1}) {
            const size_t smem = mode != 1 ? mccfr_tree_smem(H.S, H.N, a.nframes, TREE_THREADS, ncopy)
                                          : es_tree_smem(H.S, H.N, a.nframes, TREE_THREADS, ncopy);
            if (smem > 227 * 1024) continue;
            a.ncopy = ncopy;
            return emu_launch_grid(mode != 1 ? tree_entry : es_tree_entry, a, grid(TREE_THREADS), TREE_THREADS);
        }
        return -4;
    }
    if (mode == 3) {
        a.nframes = H.nframes;
        int threads = MCCFR_THREADS;
        while (threads > 128 && mccfr_batch_smem(H.S, H.hcap, H.nframes, threads) > 227 * 1024) threads -= 128;
        return emu_launch_grid(restep_entry, a, grid(threads), threads);
    }
    if (mode == 2) return emu_launch_grid(os_entry, a, grid(256), 256);
    return -5;
}

int host_mccfr_apply() { return emu_launch_grid(apply_entry, H.dev, (unsigned)((H.S + 255) / 256), 256); }

// ms_mccfr_inplace_many: mccfr_inplace_many_kernel<<<ceil(n_runs / warps), 256, smem>>>, one warp per run
int host_mccfr_inplace_many(int n_runs, long long iters, unsigned long long seed0, unsigned long long first_iter,
                            double* regret, double* strategy, uint8_t* touched) {
    const size_t tree_b = (4 * (size_t)H.N + 15) & ~(size_t)15, per_warp = inplace_many_warp_bytes(H.S, H.nframes_tree);
    int warps = (int)((227 * 1024 - tree_b) / per_warp);
    if (warps > 8) warps = 8;
    if (warps < 1) return -4;
    if (warps > n_runs) warps = n_runs;
    ManyRunArgs a{H.dev, ManyRuns{regret, strategy, touched, n_runs}, iters, seed0, first_iter, H.nframes_tree, warps};
    return emu_launch_grid(inplace_many_entry, a, (unsigned)((n_runs + warps - 1) / warps), 256);
}

// ms_mccfr_apply_peers with TWO emulated ranks in this process: rank 0 is the solver H, rank 1 a second replica of its
// table (`regret1` / `strategy1` / `touched1`, caller-owned) with its own delta buffer `delta1` [6S]; both ranks' kernels
// (one block each) run CONCURRENTLY as an emulated cluster launch, meet at the flag barrier, sum the two delta buffers
// in rank order and update their replicas.  `absent` = 1: rank 1 never arrives (its kernel is not launched), so rank 0
// must time out, set its error word and leave its table alone.  -> rank 0's error word (0 = fine), or < 0.
static std::vector<double> g_inbox[2];                // one inbox per emulated rank: [2 ranks][6 S]
static unsigned long long g_flags[2][MS_MAX_PEERS];
static unsigned int g_err[2][4];
static unsigned long long g_epoch = 0;
int host_apply_peers(double* regret1, double* strategy1, uint8_t* touched1, double* delta1, int absent) {
    const size_t n = 6 * (size_t)H.S;
    for (int r = 0; r < 2; r++) g_inbox[r].assign(2 * n, 0.0);
    PeersArgs a{};
    a.d[0] = H.dev;
    a.d[1] = H.dev; a.d[1].regret = regret1; a.d[1].strategy = strategy1; a.d[1].touched = touched1; a.d[1].delta = delta1;
    for (int r = 0; r < 2; r++) {
        PeerView& pv = a.pv[r];
        pv.inbox[0] = g_inbox[0].data(); pv.inbox[1] = g_inbox[1].data();
        pv.flags[0] = g_flags[0]; pv.flags[1] = g_flags[1];
        pv.my_flags = g_flags[r]; pv.rank = r; pv.world = 2;
        a.err[r] = g_err[r];
    }
    a.epoch = ++g_epoch;
    int rc;
    if (absent) rc = emu_launch_cluster(peers_entry, a, 1, 1024);     // only block 0 = rank 0 runs
    else rc = emu_launch_cluster(peers_entry, a, 2, 1024);
    if (rc) return -1;
    return (int)g_err[0][0];          // (the kernel itself clears each rank's delta buffer after pushing it)
}
// ms_mccfr_batch_peers (mccfr_static_peers_kernel: traversals + exchange in one launch) with two emulated ranks, G CTAs
// each.  Rank r runs traversal ids [first_r, first_r + n_trav) with n_trav <= (G - 1) * 1024, so that CTAs 0 .. G-2 do all
// the traversal work and can run one after another; the LAST CTA of each rank (no traversals left for it: it only takes
// the final ticket and performs the exchange) runs concurrently with the other rank's last CTA.
static unsigned int g_ticket[2][4];
int host_batch_peers(double* regret1, double* strategy1, uint8_t* touched1, double* delta1, unsigned long long* counters1,
                     long long n_trav, unsigned long long philox_seed, unsigned long long first0, unsigned long long first1, int G) {
    StaticDims dm{};
    if (!host_static_dims(dm) || G < 2 || n_trav > (long long)(G - 1) * STATIC_THREADS) return -5;
    const size_t n = 6 * (size_t)H.S;
    for (int r = 0; r < 2; r++) g_inbox[r].assign(2 * n, 0.0);
    FusedArgs a{};
    a.d[0] = H.dev;
    a.d[1] = H.dev; a.d[1].regret = regret1; a.d[1].strategy = strategy1; a.d[1].touched = touched1; a.d[1].delta = delta1;
    a.d[1].counters = counters1;
    for (int r = 0; r < 2; r++) {
        PeerView& pv = a.pv[r];
        pv.inbox[0] = g_inbox[0].data(); pv.inbox[1] = g_inbox[1].data();
        pv.flags[0] = g_flags[0]; pv.flags[1] = g_flags[1];
        pv.my_flags = g_flags[r]; pv.rank = r; pv.world = 2;
        a.err[r] = g_err[r]; a.ticket[r] = g_ticket[r]; g_ticket[r][0] = 0;
    }
    a.first_trav[0] = first0; a.first_trav[1] = first1;
    a.n_trav = n_trav; a.key = make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32));
    a.epoch = ++g_epoch;
    a.dm = dm;
    gridDim = {(unsigned)G, 1, 1};
    for (int r = 0; r < 2; r++) {
        a.force_rank = r;
        for (int b = 0; b + 1 < G; b++)
            if (emu_run_blocks(fused_entry, a, (unsigned)b, 1, STATIC_THREADS)) return -1;
    }
    a.force_rank = -1;                    // the two last CTAs, one per rank (slot = rank), side by side
    if (emu_run_blocks(fused_entry, a, (unsigned)(G - 1), 2, STATIC_THREADS)) return -1;
    gridDim = {1, 1, 1};
    if (g_ticket[0][0] != 0 || g_ticket[1][0] != 0) return -6;       // the last CTA re-arms the ticket
    return (int)g_err[0][0];
}

void host_peers_reset() { g_epoch = 0; for (int r = 0; r < 2; r++) { for (auto& f : g_flags[r]) f = 0; for (auto& e : g_err[r]) e = 0; } }

// ms_cfr_iterate_many: cfr_many_kernel<<<n_jobs, 512, smem>>>, one CTA per job.  The shim holds one solver, so the jobs
// are `n_jobs` references to it and the emulator runs the CTAs one after another: n_jobs x iters iterations in all.
int host_cfr_many(int n_jobs, int iters) {
    std::vector<CfrJob> jobs((size_t)n_jobs, CfrJob{H.dev, H.n_dec});
    ManyArgs a{jobs.data(), iters};
    return emu_launch_grid(cfr_many_entry, a, (unsigned)n_jobs, 512);
}

// ms_best_response: best_response_kernel<<<1, 512, cfr_smem_bytes(...)>>>
int host_best_response(int kind, double* out2) {
    if (cfr_smem_bytes(H.N, H.S, H.n_dec) > EMU_SMEM_BYTES) return -4;
    BrArgs a{H.dev, H.n_dec, kind, out2};
    return emu_launch_cta(br_entry, a, 512);
}

// ms_solver_policy / ms_eval_policies (a few blocks of 256 threads instead of the device's grid: same grid-stride loops)
int host_policy(int kind, double* out) {
    PolArgs a{H.dev, kind, out};
    return emu_launch_grid(policy_entry, a, (unsigned)((H.S + 255) / 256), 256);
}
int host_eval(const double* pol0, const double* pol1, long long n, unsigned long long philox_seed, unsigned long long first,
              float* rew, uint8_t* scopas) {
    EvalArgs a{H.dev, pol0, pol1, n, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)), first, rew, (uchar2*)scopas};
    return emu_launch_grid(eval_entry, a, 4, 256);
}

// the slot-aligned delta buffer the multi-GPU exchange sums over ranks: [S][4] regret deltas, [S] update counts, [S] first-touch marks
void host_solver_delta(double* out6S) { for (size_t i = 0; i < H.delta.size(); i++) out6S[i] = H.delta[i]; }

void host_solver_set_delta(const double* in6S) { for (size_t i = 0; i < H.delta.size(); i++) H.delta[i] = in6S[i]; }

double host_solver_delta_abs_sum() {
    double t = 0.0;
    for (double v : H.delta) t += v < 0 ? -v : v;
    return t;
}

void host_solver_counters(unsigned long long* out3, uint8_t* touched, int reset) {
    for (int i = 0; i < 3; i++) { out3[i] = H.counters[i]; if (reset) H.counters[i] = 0; }
    if (touched) for (int s = 0; s < H.S; s++) touched[s] = H.touched[s];
}

}  // extern "C"
