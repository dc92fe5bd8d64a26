// scopa_b200/csrc/ms_solver.cu -- one-deal solver state on the device: game-tree enumeration, the
// slot-aligned infoset table, vanilla CFR (order-exact level-synchronous sweep), the reference's
// sampled-CFR estimator (in-place and batched), and the best-response sweep.  sm_100a only.
//
// Replaces (paths relative to /root/reference/):
//   CFRTrainer._cfr_recursive / train      src/algorithms/vanilla_cfr.py:56-120
//   MCCFRTrainer._sample / iteration       src/algorithms/mc_cfr.py:37-92
//   the dict-of-InfoNode tables            vanilla_cfr.py:49-54, mc_cfr.py:28-35
//   exploitability.exploitability(...)     vanilla_cfr.py:112-118 (third-party open_spiel, restated)
#include <algorithm>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "ms_common.cuh"
#include "ms_state.cuh"
#include "ms_tree_walk.cuh"
#include "ms_static_walk.cuh"

#ifndef MS_DYN_SMEM   // the host emulation of tests/emu/ms_solver_host.cpp supplies its own (one buffer per block)
#define MS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace ms {

constexpr int MAXN = 4096;   // tree nodes (a 4+4-card deal has 2229)
constexpr int MAXL = 34;     // levels
constexpr int MAXS = 1664;   // infoset slots

// ------------------------------------------------------------------------------------------------
// Tree enumeration: level-synchronous expansion with the same step()/legal_list() as the env
// kernels.  Children are appended in hand (= legal_actions) order, so inside one level the node
// index order equals the reference's depth-first visiting order restricted to that level.
struct TreeOut {
    uint4* state; int* parent; int* child_begin; uint8_t* nchild; unsigned long long* key;
    uint16_t* legal; int8_t* rx2; int* level_begin; int* counts;   // counts: [0] nodes [1] levels [2] overflow
};

__global__ void __launch_bounds__(256) tree_expand_kernel(uint4 root, uint32_t hand_order, TreeOut t) {
    __shared__ int s_begin, s_end, s_total, s_stop;
    const int tid = threadIdx.x, bd = blockDim.x;
    if (tid == 0) {
        t.state[0] = root; t.parent[0] = -1; t.level_begin[0] = 0;
        s_begin = 0; s_end = 1; s_stop = 0; t.counts[2] = 0;
    }
    __syncthreads();
    int lvl = 0;
    for (;; lvl++) {
        const int begin = s_begin, end = s_end;
        for (int i = begin + tid; i < end; i += bd) {
            const MsState s = t.state[i];
            const bool term = st_terminal(s);
            uint32_t list = 0u;
            const uint32_t nl = term ? 0u : legal_list(s, hand_order, st_cur(s), list);
            t.nchild[i] = (uint8_t)nl;
            t.legal[i] = (uint16_t)list;
            t.key[i] = term ? 0xFFFFFFFFFFFFFFFFull : infoset_key(s, st_cur(s));
            t.rx2[i] = term ? (int8_t)reward0_x2(s) : (int8_t)0;
        }
        __syncthreads();
        if (tid == 0) {
            int acc = end;
            for (int i = begin; i < end; i++) { t.child_begin[i] = acc; acc += t.nchild[i]; }
            if (acc > MAXN || lvl + 2 > MAXL) { t.counts[2] = 1; acc = end; }
            s_total = acc;
            if (acc == end) s_stop = 1;
        }
        __syncthreads();
        if (s_stop) break;
        for (int i = begin + tid; i < end; i += bd) {
            const int nc = t.nchild[i], cb = t.child_begin[i];
            const uint32_t list = t.legal[i];
            for (int k = 0; k < nc; k++) {
                MsState c = t.state[i];
                step(c, (list >> (4 * k)) & 0xFu);
                t.state[cb + k] = c;
                t.parent[cb + k] = i;
            }
        }
        __syncthreads();
        if (tid == 0) { t.level_begin[lvl + 1] = end; s_begin = end; s_end = s_total; }
        __syncthreads();
    }
    if (tid == 0) { t.level_begin[lvl + 1] = s_end; t.counts[0] = s_end; t.counts[1] = lvl + 1; }
}

// ------------------------------------------------------------------------------------------------
// Device-side view of a solver (plain pointers; passed to kernels by value).
struct SolverDev {
    // tree
    int n_nodes, n_levels, n_slots, root_cur;
    uint4 root; uint32_t hand_order;
    const int* level_begin;        // [n_levels + 1]
    const uint16_t* child_begin;   // [n_nodes]
    const uint8_t* nchild;         // [n_nodes]
    const int16_t* node_slot;      // [n_nodes] (-1 terminal)
    const int8_t* rx2;             // [n_nodes] 2 * reward of player 0 at terminals
    // slots
    const uint16_t* chain_begin;   // [n_slots + 1]
    const uint16_t* chain_nodes;   // [n_decision] node ids, ascending inside a chain
    const int* slot_level_begin;   // [n_levels + 1]
    const uint8_t* slot_nlegal;    // [n_slots]
    const uint8_t* slot_player;    // [n_slots]
    // hash index key -> slot (open addressing, linear probing)
    const unsigned long long* hkeys; const int16_t* hslots; int hcap;
    // table
    double* regret; double* strategy;   // [n_slots][4]
    double* delta;                      // [n_slots][4] regret deltas, then [n_slots] update counts, then [n_slots] first-touch marks
    uint8_t* touched;                   // [n_slots]
    unsigned long long* counters;       // [0] updates [1] visits [2] env steps
};


// ------------------------------------------------------------------------------------------------
// K4  vanilla CFR.  One CTA owns the whole iteration; tree values, reaches, regrets and current
// strategies live in shared memory for all iterations of the launch.
//
// The reference refreshes an infoset's strategy after EVERY visit (vanilla_cfr.py:97), so one
// traversal is a Gauss-Seidel sweep whose result depends on the depth-first visiting order
// (SURVEY.md H3).  Exact level-synchronous restatement, per traversal for traverser tp:
//   A  top-down : opponent reach of every node (the opponent's strategy cannot change during the
//                 traversal: only tp's regrets are written);
//   B  bottom-up: opponent levels take the expectation under the opponent's strategy; at traverser
//                 levels each infoset walks ITS nodes in depth-first order (= ascending node index
//                 within the level) as a sequential chain: u = sigma.u_children, regret += opp*(u_a-u),
//                 sigma <- RM(regret); the sigma each node used is parked in its children's (now dead)
//                 value slots;
//   C  top-down : own reach from the parked sigmas, then strategy_sum += reach * sigma_used along
//                 the same chains.
// All arithmetic is float64 with explicit round-to-nearest mul/add/div (no FMA contraction) in the
// reference's operation order -> results are bit-identical to numpy's.
struct CfrSmem {
    double* u; double* r; double* reg; double* sig;
    uint16_t* child_begin; uint16_t* chain_nodes; uint16_t* chain_begin; int16_t* node_slot;
    uint8_t* nchild; int8_t* rx2; uint8_t* nlegal;
};

__host__ __device__ inline size_t cfr_smem_bytes(int n_nodes, int n_slots, int n_dec) {
    size_t b = 0;
    b += sizeof(double) * (size_t)n_nodes * 2;
    b += sizeof(double) * (size_t)n_slots * 8;
    b += sizeof(uint16_t) * ((size_t)n_nodes + n_dec + n_slots + 1) + sizeof(int16_t) * (size_t)n_nodes;
    b += (size_t)n_nodes * 2 + n_slots;
    return b + 64;
}

__device__ __forceinline__ CfrSmem cfr_carve(unsigned char* base, int N, int S, int D) {
    CfrSmem m;
    m.u = (double*)base; m.r = m.u + N; m.reg = m.r + N; m.sig = m.reg + 4 * S;
    m.child_begin = (uint16_t*)(m.sig + 4 * S);
    m.chain_nodes = m.child_begin + N;
    m.chain_begin = m.chain_nodes + D;
    m.node_slot = (int16_t*)(m.chain_begin + S + 1);
    m.nchild = (uint8_t*)(m.node_slot + N);
    m.rx2 = (int8_t*)(m.nchild + N);
    m.nlegal = (uint8_t*)(m.rx2 + N);
    return m;
}

__device__ void cfr_traversal(const SolverDev& d, const CfrSmem& m, int tp, double r0, double r1,
                              const int* s_lvl, const int* s_slvl) {
    const int tid = threadIdx.x, bd = blockDim.x;
    const int L = d.n_levels;
    // ---- A: opponent reach
    if (tid == 0) m.r[0] = (tp == 0) ? r1 : r0;
    __syncthreads();
    for (int l = 0; l + 1 < L; l++) {
        const int cp = (d.root_cur + l) & 1;
        for (int v = s_lvl[l] + tid; v < s_lvl[l + 1]; v += bd) {
            const int nc = m.nchild[v];
            if (nc == 0) continue;
            const int cb = m.child_begin[v];
            const double ro = m.r[v];
            if (cp != tp) {
                const double* sg = m.sig + 4 * m.node_slot[v];
                for (int i = 0; i < nc; i++) m.r[cb + i] = __dmul_rn(ro, sg[i]);
            } else {
                for (int i = 0; i < nc; i++) m.r[cb + i] = ro;
            }
        }
        __syncthreads();
    }
    // ---- B: values bottom-up, regret chains at traverser levels
    for (int l = L - 1; l >= 0; l--) {
        const int cp = (d.root_cur + l) & 1;
        for (int v = s_lvl[l] + tid; v < s_lvl[l + 1]; v += bd) {
            const int nc = m.nchild[v];
            if (nc == 0) {
                const int rr = m.rx2[v];
                m.u[v] = 0.5 * (double)(tp == 0 ? rr : -rr);     // +0.0 on ties for both players
            } else if (cp != tp) {
                const int cb = m.child_begin[v];
                const double* sg = m.sig + 4 * m.node_slot[v];
                double acc = 0.0;
                for (int i = 0; i < nc; i++) acc = __dadd_rn(acc, __dmul_rn(sg[i], m.u[cb + i]));
                m.u[v] = acc;
            } else if (nc == 1) {
                // traverser node with one legal action: sigma = [1.0] before and after the visit and the
                // regret delta is opp * (u - 1.0 * u) = 0, so there is no chain dependency: per-node work
                const int cb = m.child_begin[v];
                m.u[v] = __dadd_rn(0.0, __dmul_rn(1.0, m.u[cb]));
                m.u[cb] = 1.0;                            // sigma_used, parked like the chain does
            }
        }
        if (cp == tp) {
            for (int s = s_slvl[l] + tid; s < s_slvl[l + 1]; s += bd) {
                const int n = m.nlegal[s];
                if (n == 1) continue;                     // handled per node above
                double reg[4], sg[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { reg[i] = m.reg[4 * s + i]; sg[i] = m.sig[4 * s + i]; }
                for (int k = m.chain_begin[s]; k < m.chain_begin[s + 1]; k++) {
                    const int v = m.chain_nodes[k];
                    const int cb = m.child_begin[v];
                    double au[4];
                    double util = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (i < n) { au[i] = m.u[cb + i]; util = __dadd_rn(util, __dmul_rn(sg[i], au[i])); }
                    const double opp = m.r[v];
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (i < n) {
                            reg[i] = __dadd_rn(reg[i], __dmul_rn(opp, __dadd_rn(au[i], -util)));
                            m.u[cb + i] = sg[i];          // park sigma_used in the dead child slot
                        }
                    m.u[v] = util;
                    regret_match(reg, n, sg);             // refreshed after every visit (:97)
                }
#pragma unroll
                for (int i = 0; i < 4; i++) { m.reg[4 * s + i] = reg[i]; m.sig[4 * s + i] = sg[i]; }
            }
        }
        __syncthreads();
    }
    // ---- C: own reach top-down, then strategy sums along the chains
    if (tid == 0) m.r[0] = (tp == 0) ? r0 : r1;
    __syncthreads();
    for (int l = 0; l + 1 < L; l++) {
        const int cp = (d.root_cur + l) & 1;
        for (int v = s_lvl[l] + tid; v < s_lvl[l + 1]; v += bd) {
            const int nc = m.nchild[v];
            if (nc == 0) continue;
            const int cb = m.child_begin[v];
            const double rt = m.r[v];
            if (cp == tp) for (int i = 0; i < nc; i++) m.r[cb + i] = __dmul_rn(rt, m.u[cb + i]);
            else for (int i = 0; i < nc; i++) m.r[cb + i] = rt;
        }
        __syncthreads();
    }
    for (int s = tid; s < d.n_slots; s += bd) {
        if (d.slot_player[s] != tp) continue;
        const int n = m.nlegal[s];
        double acc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = d.strategy[4 * s + i];
        for (int k = m.chain_begin[s]; k < m.chain_begin[s + 1]; k++) {
            const int v = m.chain_nodes[k];
            const int cb = m.child_begin[v];
            const double rt = m.r[v];
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < n) acc[i] = __dadd_rn(acc[i], __dmul_rn(rt, m.u[cb + i]));
        }
#pragma unroll
        for (int i = 0; i < 4; i++) d.strategy[4 * s + i] = acc[i];
    }
    __syncthreads();
}

__device__ void cfr_run(const SolverDev& d, int n_dec, int iters, int only_player, double r0, double r1, double* out_value,
                        unsigned char* smem_raw, int* s_lvl, int* s_slvl);

__global__ void __launch_bounds__(512, 1) cfr_kernel(SolverDev d, int n_dec, int iters, int only_player, double r0,
                                                     double r1, double* out_value) {
    MS_DYN_SMEM(smem_raw);
    __shared__ int s_lvl[MAXL + 1], s_slvl[MAXL + 1];
    cfr_run(d, n_dec, iters, only_player, r0, r1, out_value, smem_raw, s_lvl, s_slvl);
}

// Throughput mode: independent deals solved side by side, one CTA per deal (one CTA per SM).
struct CfrJob { SolverDev d; int n_dec; };
__global__ void __launch_bounds__(512, 1) cfr_many_kernel(const CfrJob* __restrict__ jobs, int iters) {
    MS_DYN_SMEM(smem_raw);
    __shared__ int s_lvl[MAXL + 1], s_slvl[MAXL + 1];
    const CfrJob job = jobs[blockIdx.x];
    cfr_run(job.d, job.n_dec, iters, -1, 1.0, 1.0, nullptr, smem_raw, s_lvl, s_slvl);
}

__device__ void cfr_run(const SolverDev& d, int n_dec, int iters, int only_player, double r0, double r1, double* out_value,
                        unsigned char* smem_raw, int* s_lvl, int* s_slvl) {
    const int tid = threadIdx.x, bd = blockDim.x;
    const int N = d.n_nodes, S = d.n_slots;
    CfrSmem m = cfr_carve(smem_raw, N, S, n_dec);
    for (int i = tid; i <= d.n_levels; i += bd) { s_lvl[i] = d.level_begin[i]; s_slvl[i] = d.slot_level_begin[i]; }
    for (int i = tid; i < N; i += bd) {
        m.child_begin[i] = d.child_begin[i]; m.nchild[i] = d.nchild[i];
        m.node_slot[i] = d.node_slot[i]; m.rx2[i] = d.rx2[i];
    }
    for (int i = tid; i < n_dec; i += bd) m.chain_nodes[i] = d.chain_nodes[i];
    for (int i = tid; i <= S; i += bd) m.chain_begin[i] = d.chain_begin[i];
    for (int s = tid; s < S; s += bd) {
        const int n = d.slot_nlegal[s];
        m.nlegal[s] = (uint8_t)n;
        double reg[4], sg[4];
        for (int i = 0; i < 4; i++) { reg[i] = d.regret[4 * s + i]; m.reg[4 * s + i] = reg[i]; }
        regret_match(reg, n, sg);     // invariant: local_strategy == RM(regret_sum) between visits
        for (int i = 0; i < 4; i++) m.sig[4 * s + i] = sg[i];
    }
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        for (int tp = 0; tp < 2; tp++) {
            if (only_player >= 0 && tp != only_player) continue;
            cfr_traversal(d, m, tp, r0, r1, s_lvl, s_slvl);
        }
    }
    for (int i = tid; i < 4 * S; i += bd) d.regret[i] = m.reg[i];
    if (tid == 0 && out_value) *out_value = m.u[0];
}

// ------------------------------------------------------------------------------------------------
// K3  the reference's sampled-CFR estimator (mc_cfr.py:37-86).
//
// One thread = one traversal, run as an explicit depth-first search whose stack lives in shared
// memory.  The recursion shape is the same for every traversal (1 + |hand| recursive calls at a
// traverser node, 1 at an opponent node), so the lanes of a warp stay convergent.  Opponent nodes
// are tail calls and need no frame; only traverser nodes push one:
//   state (16 B) | opp reach f64 | own sampling prob f64 | slot, n_legal, legal list, child cursor |
//   returned child values as exact bytes (a returned utility is always a terminal reward = k/2).
// Infosets are found by probing the open-addressing key->slot index with the packed 64-bit key.
// Sampling: u = u53(Philox4x32-10(key = seed, ctr = (traversal id lo, hi, call index, "MCCF"+tp)))
// and numpy's cumsum / normalise / searchsorted-right rule, so that the CPU oracle can follow the
// same stream (call index = order of _sample invocations inside the traversal).
struct Frames {            // SoA: [frame][thread]
    uint4* st; double* ro; double* sp; uint2* meta;
    int stride;            // threads per CTA
};
// meta.x: slot (bits 0-11) | n_legal (12-14) | child cursor (15-17) | legal list (16 bits at 18..)  -> needs 34 bits;
// so: meta.x = slot | n_legal << 12 | cursor << 16 ;  meta.y = legal list (16) | util byte << 16 ;  cfv bytes
// are kept in a third word.
struct Frames3 { uint32_t* cfv; };

__device__ __forceinline__ int lookup_slot(const unsigned long long* hk, const int16_t* hs, int hcap,
                                           unsigned long long key) {
    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 40) & (uint32_t)(hcap - 1);
    while (true) {
        const unsigned long long k = hk[h];
        if (k == key) return hs[h];
        if (k == 0xFFFFFFFFFFFFFFFFull) return -1;
        h = (h + 1) & (uint32_t)(hcap - 1);
    }
}


// searchsorted(cdf, u, side='right') on a precomputed normalised cdf (entries beyond n are 2.0)
__device__ __forceinline__ int sample_cdf(const double* cdf, int n, double u) {
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) idx += (cdf[i] <= u) ? 1 : 0;
    return idx < n ? idx : n - 1;
}


template <bool INPLACE>
__device__ void mccfr_traverse(const SolverDev& d, const MccfrShared& sh, int tp, unsigned long long trav,
                               uint2 pkey, uint4* f_st, double* f_ro, double* f_sp, uint2* f_meta, uint32_t* f_cfv,
                               int fstride, unsigned long long& n_upd, unsigned long long& n_vis,
                               unsigned long long& n_step) {
    MsState s = d.root;
    const uint32_t dealt = dealt_set(d.root);
    double ro = 1.0, sp = 1.0;
    int fi = -1;
    uint4 xblk = make_uint4(0u, 0u, 0u, 0u);   // cached Philox block: serves call indices 2b and 2b+1
    uint32_t xblk_id = 0xFFFFFFFFu;
    // Descents are deferred to the top of the loop so that step() -- the largest piece of code here -- is
    // instantiated once instead of at every descent site (the kernel was 57 KB of SASS, beyond the 32 KB
    // L1.5 instruction cache; ncu showed no_instruction stalls).
    bool pend = false;
    uint32_t pend_a = 0u;
    uint32_t call = 0;
    int ret_x2 = 0;
    bool returning = false;
    const uint32_t tag = MS_TAG_MCCF + (uint32_t)tp;
    while (true) {
        if (!returning) {
            if (pend) { step(s, pend_a, table_set_from_dealt(s, dealt)); n_step++; pend = false; }
            const uint32_t my_call = call++;
            n_vis++;
            if (st_terminal(s)) {
                const int r = reward0_x2(s);
                ret_x2 = (tp == 0) ? r : -r;
                returning = true;
                continue;
            }
            const int p = st_cur(s);
            uint32_t list;
            const uint32_t nl = legal_list(s, d.hand_order, p, list);
            const int slot = lookup_slot(sh.hk, sh.hs, sh.hcap, infoset_key(s, p));
            sh.touched[slot] = 1;     // node created on first touch, for both players (mc_cfr.py:52)
            if (nl == 1u) {
                // Forced move (one legal action): sigma = [1.0], so the draw cannot change anything and no
                // random word is generated (the call index still advances: the oracle's draw is a no-op too).
                const uint32_t a1 = list & 0xFu;
                if (p != tp) {            // opponent: reach *= 1.0
                    pend_a = a1; pend = true;
                    continue;
                }
                // Traverser's last card.  If the rest of the game is forced as well (the opponent holds at
                // most one card) both recursive calls of the reference (:58-67 and :71-78) walk the same
                // deterministic line, so it is played once and accounted twice.
                MsState t2 = s;
                int below = 0, slot2 = -1;
                bool forced = false;
                uint32_t act = a1;
#pragma unroll 1
                for (int k = 0; k < 2; k++) {
                    step(t2, act, table_set_from_dealt(t2, dealt));
                    below++;
                    if (st_terminal(t2)) { forced = true; break; }
                    if (k == 1 || __popc(st_hand(t2, p ^ 1)) != 1) break;
                    uint32_t l2;
                    legal_list(t2, d.hand_order, p ^ 1, l2);
                    slot2 = lookup_slot(sh.hk, sh.hs, sh.hcap, infoset_key(t2, p ^ 1));
                    act = l2 & 0xFu;
                }
                if (forced) {
                    if (slot2 >= 0) sh.touched[slot2] = 1;
                    const int r = reward0_x2(t2);
                    ret_x2 = (tp == 0) ? r : -r;
                    // regret delta = w * (cfv - sigma.cfv) = w * 0 exactly; strategy_sum += 1.0 * sigma = 1.0
                    if (INPLACE) sh.str[4 * slot] = __dadd_rn(sh.str[4 * slot], 1.0);
                    else atomicAdd(&sh.dcnt[slot], 1u);
                    n_upd++;
                    n_vis += 2 * below; call += 2u * (uint32_t)below; n_step += below;   // visits: reference-equivalent; steps: executed
                    returning = true;
                    continue;
                }
            }
            double sg[4];
            if (INPLACE) regret_match(sh.reg + 4 * slot, (int)nl, sg);
            else {
#pragma unroll
                for (int i = 0; i < 4; i++) sg[i] = sh.sig[4 * slot + i];
            }
            int ai = 0;
            if (nl > 1u) {
                if ((my_call >> 1) != xblk_id) {
                    xblk_id = my_call >> 1;
                    xblk = philox4x32_10(make_uint4((uint32_t)trav, (uint32_t)(trav >> 32), xblk_id, tag), pkey);
                }
                const double u = (my_call & 1u) ? u53(xblk.z, xblk.w) : u53(xblk.x, xblk.y);
                ai = INPLACE ? sample_action(sg, (int)nl, u) : sample_cdf(sh.cdf + 4 * slot, (int)nl, u);
            }
            const uint32_t a = (list >> (4 * ai)) & 0xFu;
            if (p != tp) {            // opponent: reach *= sigma[a]; tail call (mc_cfr.py:63-65)
                ro = __dmul_rn(ro, sg[ai]);
                pend_a = a; pend = true;
                continue;
            }
            // traverser: push a frame, descend into the sampled action first (:58-67)
            fi++;
            const int o = fi * fstride;
            f_st[o] = s; f_ro[o] = ro; f_sp[o] = sp;
            f_meta[o] = make_uint2((uint32_t)slot | (nl << 12), list);
            f_cfv[o] = 0u;
            sp = __dmul_rn(sp, sg[ai]);
            pend_a = a; pend = true;
            continue;
        }
        // ---- a child returned ret_x2 to the top frame
        if (fi < 0) break;
        const int o = fi * fstride;
        uint2 meta = f_meta[o];
        const int slot = (int)(meta.x & 0xFFFu);
        const int nl = (int)((meta.x >> 12) & 0x7u);
        int cur = (int)((meta.x >> 16) & 0x7u);
        uint32_t cfvb = f_cfv[o];
        if (cur == 0) meta.y = (meta.y & 0xFFFFu) | (((uint32_t)ret_x2 & 0xFFu) << 16);   // util of the sampled action
        else cfvb |= ((uint32_t)ret_x2 & 0xFFu) << (8 * (cur - 1));
        cur++;
        double sg[4];
        if (INPLACE) regret_match(sh.reg + 4 * slot, nl, sg);   // unchanged since entry: an infoset cannot recur below itself
        else {
#pragma unroll
            for (int i = 0; i < 4; i++) sg[i] = sh.sig[4 * slot + i];
        }
        if (cur <= nl) {              // evaluate action i = cur-1 with a fresh sampled continuation (:71-78)
            const int i = cur - 1;
            meta.x = (meta.x & 0xFFFFu) | ((uint32_t)cur << 16);
            f_meta[o] = meta; f_cfv[o] = cfvb;
            s = f_st[o];
            ro = f_ro[o];
            sp = __dmul_rn(f_sp[o], sg[i]);
            pend_a = (meta.y >> (4 * i)) & 0xFu; pend = true;
            returning = false;
            continue;
        }
        // ---- all actions evaluated: regret / strategy update (:79-84)
        double cfv[4];
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cfv[i] = 0.5 * (double)(int)(int8_t)((cfvb >> (8 * i)) & 0xFFu);
            if (i < nl) v = __dadd_rn(v, __dmul_rn(sg[i], cfv[i]));
        }
        const double fro = f_ro[o], fsp = f_sp[o];
        const double w = fsp > 0.0 ? __ddiv_rn(fro, fsp) : 0.0;
        if (INPLACE) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < nl) {
                    sh.reg[4 * slot + i] = __dadd_rn(sh.reg[4 * slot + i], __dmul_rn(w, __dadd_rn(cfv[i], -v)));
                    sh.str[4 * slot + i] = __dadd_rn(sh.str[4 * slot + i], __dmul_rn(1.0, sg[i]));  // reach_probs[tp] is always 1.0
                }
        } else {
            if (nl > 1) {             // |A| = 1: cfv - v == 0 exactly
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (i < nl) atomicAdd(&sh.dreg[4 * slot + i], __dmul_rn(w, __dadd_rn(cfv[i], -v)));
            }
            atomicAdd(&sh.dcnt[slot], 1u);   // strategy delta = count * sigma (sigma is frozen for the batch)
        }
        n_upd++;
        ret_x2 = (int)(int8_t)((meta.y >> 16) & 0xFFu);
        fi--;
        returning = true;
    }
}

// in-place mode: one thread, table in shared memory, reference semantics (every update is visible to
// the next node visit).  Used for parity / curve validation, not for throughput.
__global__ void __launch_bounds__(32, 1) mccfr_inplace_kernel(SolverDev d, long long iters, uint2 pkey,
                                                             unsigned long long first_iter, int nframes) {
    MS_DYN_SMEM(smem_raw);
    const int S = d.n_slots;
    double* reg = (double*)smem_raw;
    double* str = reg + 4 * S;
    unsigned long long* hk = (unsigned long long*)(str + 4 * S);
    uint4* f_st = (uint4*)(hk + d.hcap);
    double* f_ro = (double*)(f_st + nframes);
    double* f_sp = f_ro + nframes;
    uint2* f_meta = (uint2*)(f_sp + nframes);
    uint32_t* f_cfv = (uint32_t*)(f_meta + nframes);
    int16_t* hs = (int16_t*)(f_cfv + nframes);
    uint8_t* touched = (uint8_t*)(hs + d.hcap);
    const int tid = threadIdx.x;
    for (int i = tid; i < 4 * S; i += 32) { reg[i] = d.regret[i]; str[i] = d.strategy[i]; }
    for (int i = tid; i < d.hcap; i += 32) { hk[i] = d.hkeys[i]; hs[i] = d.hslots[i]; }
    for (int i = tid; i < S; i += 32) touched[i] = d.touched[i];
    __syncwarp();
    if (tid == 0) {
        MccfrShared sh{hk, hs, d.hcap, nullptr, nullptr, nullptr, nullptr, touched, reg, str};
        unsigned long long nu = 0, nv = 0, ns = 0;
        for (long long it = 0; it < iters; it++)
            for (int tp = 0; tp < 2; tp++)
                mccfr_traverse<true>(d, sh, tp, first_iter + (unsigned long long)it, pkey, f_st, f_ro, f_sp, f_meta,
                                     f_cfv, 1, nu, nv, ns);
        atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns);
    }
    __syncwarp();
    for (int i = tid; i < 4 * S; i += 32) { d.regret[i] = reg[i]; d.strategy[i] = str[i]; }
    for (int i = tid; i < S; i += 32) d.touched[i] = touched[i];
}

constexpr int MCCFR_THREADS = 768;

__host__ __device__ inline size_t mccfr_batch_smem(int S, int hcap, int nframes, int threads) {
    size_t b = 0;
    b += sizeof(double) * 12 * (size_t)S;                // sigma + cdf + regret deltas
    b += sizeof(unsigned long long) * (size_t)hcap;      // hash keys
    b += (size_t)threads * nframes * (16 + 8 + 8 + 8 + 4);   // frames
    b += sizeof(uint32_t) * (size_t)S;                   // counts
    b += sizeof(int16_t) * (size_t)hcap;                 // hash slots
    b += (size_t)S;                                      // touched
    return b + 64;
}

// batch mode: sigma frozen for the launch; deltas accumulate in per-CTA shared tables and are
// flushed to the global delta array with one fp64 RED per non-zero entry.
__global__ void __launch_bounds__(MCCFR_THREADS, 1) mccfr_batch_kernel(SolverDev d, int player, long long n_trav,
                                                                    uint2 pkey, unsigned long long first_trav,
                                                                    int nframes) {
    MS_DYN_SMEM(smem_raw);
    const int S = d.n_slots, T = blockDim.x, tid = threadIdx.x;
    double* sig = (double*)smem_raw;
    double* cdf = sig + 4 * S;
    double* dreg = cdf + 4 * S;
    unsigned long long* hk = (unsigned long long*)(dreg + 4 * S);
    uint4* f_st = (uint4*)(hk + d.hcap);
    double* f_ro = (double*)(f_st + (size_t)T * nframes);
    double* f_sp = f_ro + (size_t)T * nframes;
    uint2* f_meta = (uint2*)(f_sp + (size_t)T * nframes);
    uint32_t* f_cfv = (uint32_t*)(f_meta + (size_t)T * nframes);
    uint32_t* dcnt = f_cfv + (size_t)T * nframes;
    int16_t* hs = (int16_t*)(dcnt + S);
    uint8_t* touched = (uint8_t*)(hs + d.hcap);

    for (int s = tid; s < S; s += T) {
        double reg[4], sg[4];
        for (int i = 0; i < 4; i++) reg[i] = d.regret[4 * s + i];
        regret_match(reg, d.slot_nlegal[s], sg);
        double cd[4];
        strategy_cdf(sg, d.slot_nlegal[s], cd);
        for (int i = 0; i < 4; i++) { sig[4 * s + i] = sg[i]; cdf[4 * s + i] = cd[i]; dreg[4 * s + i] = 0.0; }
        dcnt[s] = 0u; touched[s] = 0;
    }
    for (int i = tid; i < d.hcap; i += T) { hk[i] = d.hkeys[i]; hs[i] = d.hslots[i]; }
    __syncthreads();

    MccfrShared sh{hk, hs, d.hcap, sig, cdf, dreg, dcnt, touched, nullptr, nullptr};
    unsigned long long nu = 0, nv = 0, ns = 0;
    const long long gstride = (long long)gridDim.x * T;
    // Threads walk the tree in lock-step, so at any moment the whole CTA updates the few infosets of one depth of
    // one player's tree and the shared-memory fp64 adds (compare-and-swap loops) collide.  Odd warps therefore run
    // player 1's traversal first: the two halves of the CTA work on disjoint infosets (+1 %).  Strategies are frozen
    // for the launch and every draw is addressed by (traversal, call index), so the order cannot change a result.
    // Also measured and dropped: summing the lanes that share a slot with match_any + shuffles before one atomic
    // per group (27.5 -> 21.4 G updates/s) and fire-and-forget global REDs into per-CTA tables in L2 (23.2 G).
    const int flip = (tid >> 5) & 1;
    for (long long k = blockIdx.x * (long long)T + tid; k < n_trav; k += gstride) {
        for (int j = 0; j < 2; j++) {
            const int tp = j ^ flip;
            if (player < 2 && tp != player) continue;
            mccfr_traverse<false>(d, sh, tp, first_trav + (unsigned long long)k, pkey, f_st + tid, f_ro + tid, f_sp + tid,
                                  f_meta + tid, f_cfv + tid, T, nu, nv, ns);
        }
    }
    __syncthreads();
    for (int i = tid; i < 4 * S; i += T) {
        const double v = dreg[i];
        if (v != 0.0) atomicAdd(&d.delta[i], v);
    }
    for (int s = tid; s < S; s += T) {
        if (dcnt[s]) atomicAdd(&d.delta[4 * S + s], (double)dcnt[s]);
        if (touched[s] && !d.touched[s]) atomicAdd(&d.delta[5 * S + s], 1.0);   // first touch: travels with the delta
    }
    // counters: warp reduce, one atomic per warp
    for (int off = 16; off > 0; off >>= 1) {
        nu += __shfl_down_sync(0xffffffffu, nu, off);
        nv += __shfl_down_sync(0xffffffffu, nv, off);
        ns += __shfl_down_sync(0xffffffffu, ns, off);
    }
    if ((tid & 31) == 0) { atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns); }
}

// ------------------------------------------------------------------------------------------------
// K3t  the same estimator walking the ENUMERATED tree.  A solver holds one deal, whose whole game tree (2229 nodes)
// was expanded at creation with the env's step() (tree_expand_kernel) -- vanilla CFR already sweeps it.  The
// re-stepping kernel above spends ~60 % of its instructions on re-deriving what that tree already says (step +
// capture resolution, the legal list, the infoset key and its hash probe: ncu source view, profiles/README.md).
// Here a node visit is one 32-bit record from shared memory:
//   bits 0-11 first child (terminal: 2 * reward of player 0, biased by 2048) | 12-22 infoset slot (0x7FF =
//   terminal) | 23-25 number of children = legal actions, in legal_actions() order | 26 player to move
// and a descent is `first child + action index`.  Same recursion, same Philox addressing by call index, same
// forced-endgame shortcut, same frozen-sigma batch semantics and delta layout as mccfr_batch_kernel -- the two
// produce the same tables (tests/test_gpu_solver.py).  Frames shrink from 44 to 26 bytes (no packed state) and the
// code to 64 registers, so a CTA runs 1024 traversals at a time instead of 768.
// in-place mode on the enumerated tree: one thread, the table itself in shared memory, reference semantics (every
// update is visible to the next node visit) -- what MCCFRTrainer.iteration() runs by default
__global__ void __launch_bounds__(32, 1) mccfr_inplace_tree_kernel(SolverDev d, long long iters, uint2 pkey,
                                                                  unsigned long long first_iter, int nframes) {
    MS_DYN_SMEM(smem_raw);
    const int S = d.n_slots, N = d.n_nodes, tid = threadIdx.x;
    double* reg = (double*)smem_raw;
    double* str = reg + 4 * S;
    TreeFrames f;
    f.ro = str + 4 * S;
    f.sp = f.ro + nframes;
    f.meta = (uint32_t*)(f.sp + nframes);
    f.cfv = f.meta + nframes;
    uint32_t* tree = f.cfv + nframes;
    f.cb = (uint16_t*)(tree + N);
    uint8_t* touched = (uint8_t*)(f.cb + nframes);
    for (int i = tid; i < 4 * S; i += 32) { reg[i] = d.regret[i]; str[i] = d.strategy[i]; }
    for (int i = tid; i < S; i += 32) touched[i] = d.touched[i];
    for (int v = tid; v < N; v += 32) {
        const int sl = d.node_slot[v];
        uint32_t rec;
        if (sl < 0) rec = ((uint32_t)((int)d.rx2[v] + 2048) & 0xFFFu) | (TREE_TERMINAL << 12);
        else rec = (uint32_t)d.child_begin[v] | ((uint32_t)sl << 12) | ((uint32_t)d.nchild[v] << 23) | ((uint32_t)d.slot_player[sl] << 26);
        tree[v] = rec;
    }
    __syncwarp();
    if (tid == 0) {
        MccfrShared sh{nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, touched, reg, str};
        unsigned long long nu = 0, nv = 0, ns = 0;
        for (long long it = 0; it < iters; it++)
            for (int tp = 0; tp < 2; tp++)
                mccfr_tree_traverse<true>(tree, sh, tp, first_iter + (unsigned long long)it, pkey, f, 1, nu, nv, ns);
        atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns);
    }
    __syncwarp();
    for (int i = tid; i < 4 * S; i += 32) { d.regret[i] = reg[i]; d.strategy[i] = str[i]; }
    for (int i = tid; i < S; i += 32) d.touched[i] = touched[i];
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) mccfr_tree_kernel(SolverDev d, int player, long long n_trav, uint2 pkey,
                                                                unsigned long long first_trav, int nframes, int ncopy) {
    MS_DYN_SMEM(smem_raw);
    const int S = d.n_slots, N = d.n_nodes, T = THREADS, tid = threadIdx.x;
    double* sig = (double*)smem_raw;
    double* cdf = sig + 4 * S;              // [S][3]: the last entry of a normalised cdf is 1.0 and is never read
    double* dreg = cdf + 3 * S;             // ncopy copies of the delta table, chosen by lane (see below)
    TreeFrames f;
    f.ro = dreg + (size_t)ncopy * 4 * S;
    f.sp = f.ro + (size_t)T * nframes;
    f.meta = (uint32_t*)(f.sp + (size_t)T * nframes);
    f.cfv = f.meta + (size_t)T * nframes;
    uint32_t* tree = f.cfv + (size_t)T * nframes;
    uint32_t* dcnt = tree + N;
    f.cb = (uint16_t*)(dcnt + S);
    uint8_t* touched = (uint8_t*)(f.cb + (size_t)T * nframes);

    for (int s = tid; s < S; s += T) {
        double reg[4], sg[4], cd[4];
        for (int i = 0; i < 4; i++) reg[i] = d.regret[4 * s + i];
        regret_match(reg, d.slot_nlegal[s], sg);
        strategy_cdf(sg, d.slot_nlegal[s], cd);
        for (int i = 0; i < 4; i++) sig[4 * s + i] = sg[i];
        for (int i = 0; i < 3; i++) cdf[3 * s + i] = cd[i];
        dcnt[s] = 0u; touched[s] = 0;
    }
    for (int i = tid; i < ncopy * 4 * S; i += T) dreg[i] = 0.0;
    for (int v = tid; v < N; v += T) {
        const int sl = d.node_slot[v];
        uint32_t rec;
        if (sl < 0) rec = ((uint32_t)((int)d.rx2[v] + 2048) & 0xFFFu) | (TREE_TERMINAL << 12);
        else rec = (uint32_t)d.child_begin[v] | ((uint32_t)sl << 12) | ((uint32_t)d.nchild[v] << 23) | ((uint32_t)d.slot_player[sl] << 26);
        tree[v] = rec;
    }
    __syncthreads();

    // A shared-memory fp64 atomicAdd is a compare-and-swap loop, and the lanes of a warp walk the tree in lock-step: at
    // an update they sit on a handful of infosets and the loops of one warp retry against each other (ncu r01g: 41 %
    // of the kernel's shared-memory wavefronts).  Several copies of the delta table (as many of 4 / 2 / 1 as shared
    // memory holds), chosen by lane id, divide that: 71.8 -> 87.2 G updates/s with two copies.
    MccfrShared sh{nullptr, nullptr, 0, sig, cdf, dreg + (size_t)(tid & (ncopy - 1)) * 4 * S, dcnt, touched, nullptr, nullptr};
    f.ro += tid; f.sp += tid; f.meta += tid; f.cfv += tid; f.cb += tid;
    unsigned long long nu = 0, nv = 0, ns = 0;
    const long long gstride = (long long)gridDim.x * T;
    const int flip = (tid >> 5) & 1;      // odd warps run player 1 first: the halves of the CTA update disjoint infosets
    for (long long k = blockIdx.x * (long long)T + tid; k < n_trav; k += gstride) {
        for (int j = 0; j < 2; j++) {
            const int tp = j ^ flip;
            if (player < 2 && tp != player) continue;
            mccfr_tree_traverse<false>(tree, sh, tp, first_trav + (unsigned long long)k, pkey, f, T, nu, nv, ns);
        }
    }
    __syncthreads();
    for (int i = tid; i < 4 * S; i += T) {
        double v = dreg[i];
        for (int c = 1; c < ncopy; c++) v = __dadd_rn(v, dreg[(size_t)c * 4 * S + i]);
        if (v != 0.0) atomicAdd(&d.delta[i], v);
    }
    for (int s = tid; s < S; s += T) {
        if (dcnt[s]) atomicAdd(&d.delta[4 * S + s], (double)dcnt[s]);
        if (touched[s] && !d.touched[s]) atomicAdd(&d.delta[5 * S + s], 1.0);   // first touch: travels with the delta
    }
    for (int off = 16; off > 0; off >>= 1) {
        nu += __shfl_down_sync(0xffffffffu, nu, off);
        nv += __shfl_down_sync(0xffffffffu, nv, off);
        ns += __shfl_down_sync(0xffffffffu, ns, off);
    }
    if ((tid & 31) == 0) { atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns); }
}

// ------------------------------------------------------------------------------------------------
// K3s  the headline kernel: the same estimator on a fresh deal's STATIC recursion shape (ms_static_walk.cuh).
// One thread = one traversal at a time, recursion state in registers; the CTA's working set (node records with
// integer cdf thresholds, sigma and 1/sigma, the private delta tables) is rebuilt in shared memory from the table at
// the start of every launch -- the frozen strategy of the batch.
__device__ __forceinline__ void mccfr_static_body(const SolverDev& d, int player, long long n_trav, uint2 pkey,
                                                  unsigned long long first_trav, const StaticDims& dm, unsigned char* smem_raw) {
    const int S = d.n_slots, T = STATIC_THREADS, tid = threadIdx.x, lane = tid & 31;
    uint4* node = (uint4*)smem_raw;
    double* sig = (double*)(node + dm.n6);
    double* rsig = sig + 4 * dm.S2;
    double* acc = rsig + 4 * dm.S2;
    uint32_t* dcnt = (uint32_t*)(acc + 32 * (size_t)dm.n_acc);
    uint8_t* touched = (uint8_t*)(dcnt + S);
    int* s_need = (int*)(((uintptr_t)(touched + S) + 15) & ~(uintptr_t)15);
    uint32_t* thr = (uint32_t*)acc;         // [S2][3] staging of the per-slot thresholds (the accumulators are zeroed afterwards)

    if (tid == 0) *s_need = 0;
    __syncthreads();
    for (int s = tid; s < S; s += T) {
        if (s < dm.S2) {
            double reg[4], sg[4], cd[4];
            const int n = d.slot_nlegal[s];
            for (int i = 0; i < 4; i++) reg[i] = d.regret[4 * s + i];
            regret_match(reg, n, sg);
            strategy_cdf(sg, n, cd);
            for (int i = 0; i < 4; i++) {
                sig[4 * s + i] = sg[i];
                rsig[4 * s + i] = sg[i] > 0.0 ? __ddiv_rn(1.0, sg[i]) : 0.0;
            }
            // T_i = ceil(cdf_i * 2^31) (exact scaling; cdf_i in [0, 1]); the last entry of a row is never read
            for (int i = 0; i < 3; i++) thr[3 * s + i] = (i + 1 < n) ? (uint32_t)ceil(cd[i] * 2147483648.0) : 0x80000000u;
        }
        dcnt[s] = 0u; touched[s] = 0;
        if (!d.touched[s]) *s_need = 1;
    }
    __syncthreads();
    for (int v = tid; v < dm.n6; v += T) {
        const int sl = d.node_slot[v];
        const uint32_t link = (uint32_t)d.child_begin[v] | ((uint32_t)sl << 12);
        if (v < dm.n5) node[v] = make_uint4(thr[3 * sl], thr[3 * sl + 1], thr[3 * sl + 2], link);
        else {              // ply 5: both children are ply-6 nodes, whose forced child (ply 7) has the leaf as its forced child
            uint32_t e[2];
            for (int k = 0; k < 2; k++) {
                const int c6 = d.child_begin[v] + k, c7 = d.child_begin[c6], leaf = d.child_begin[c7];
                e[k] = (uint32_t)d.node_slot[c6] | ((uint32_t)d.node_slot[c7] << 11) | ((uint32_t)((int)d.rx2[leaf] + 16) << 22);
            }
            node[v] = make_uint4(thr[3 * sl], link, e[0], e[1]);
        }
    }
    __syncthreads();
    for (int i = tid; i < 32 * dm.n_acc; i += T) acc[i] = 0.0;
    __syncthreads();

    StaticShared c;
    c.node = node; c.sig = sig; c.rsig = rsig;
    c.acc = acc + lane;
    c.dcnt = dcnt; c.touched = touched;
    const bool need_touch = *s_need != 0;
    c.key = pkey; c.blk = make_uint4(0u, 0u, 0u, 0u);
    unsigned long long v0, u0, e0, v1, u1, e1;
    static_shape_counts(0, v0, u0, e0);
    static_shape_counts(1, v1, u1, e1);
    unsigned long long nu = 0, nv = 0, ns = 0;
    const long long gstride = (long long)gridDim.x * T;
    const int flip = (tid >> 5) & 1;      // odd warps run player 1 first: the halves of the CTA update disjoint infosets
    for (long long k = blockIdx.x * (long long)T + tid; k < n_trav; k += gstride) {
        const unsigned long long trav = first_trav + (unsigned long long)k;
        c.t_lo = (uint32_t)trav; c.t_hi = (uint32_t)(trav >> 32);
        for (int j = 0; j < 2; j++) {
            const int tp = j ^ flip;
            if (player < 2 && tp != player) continue;
            c.nd = 0u; c.tag = MS_TAG_MCCF_SEQ + (uint32_t)tp;
            if (tp == 0) {
                if (need_touch) StaticWalk<0, 0, true>::run(0u, 1.0, c, dm); else StaticWalk<0, 0, false>::run(0u, 1.0, c, dm);
                nu += u0; nv += v0; ns += e0;
            } else {
                if (need_touch) StaticWalk<0, 1, true>::run(0u, 1.0, c, dm); else StaticWalk<0, 1, false>::run(0u, 1.0, c, dm);
                nu += u1; nv += v1; ns += e1;
            }
        }
    }
    __syncthreads();
    // flush: per infoset, D_i = sum of its lane columns; delta_a = D_a - sum_j sigma_j D_j (D_last = 0)
    for (int s = tid; s < dm.S2; s += T) {
        int p = 0, sb0 = 0, cnt = 1, ab = 0;
#pragma unroll
        for (int q = 0; q < 6; q++)         // (compile-time indices only: dm stays in the constant bank)
            if (s >= dm.sb[q] && s < dm.sb[q + 1]) { p = q; sb0 = dm.sb[q]; cnt = dm.sb[q + 1] - dm.sb[q]; ab = dm.accbase[q]; }
        const int nl = 4 - p / 2;
        double D0 = 0.0, D1 = 0.0, D2 = 0.0, sum = 0.0;
        for (int i = 0; i < nl - 1; i++) {
            const double* col = acc + 32 * (size_t)(ab + i * cnt + s - sb0);
            double t = 0.0;
            for (int l = 0; l < 32; l++) t = __dadd_rn(t, col[l]);
            if (i == 0) D0 = t; else if (i == 1) D1 = t; else D2 = t;
            sum = __dadd_rn(sum, __dmul_rn(sig[4 * s + i], t));
        }
        for (int a = 0; a < nl; a++) {
            const double Da = a == 0 ? D0 : (a == 1 ? D1 : D2);
            const double v = a < nl - 1 ? __dadd_rn(Da, -sum) : -sum;
            if (v != 0.0) atomicAdd(&d.delta[4 * s + a], v);
        }
    }
    for (int s = tid; s < S; s += T) {
        if (dcnt[s]) atomicAdd(&d.delta[4 * S + s], (double)dcnt[s]);
        if (touched[s] && !d.touched[s]) atomicAdd(&d.delta[5 * S + s], 1.0);
    }
    for (int off = 16; off > 0; off >>= 1) {
        nu += __shfl_down_sync(0xffffffffu, nu, off);
        nv += __shfl_down_sync(0xffffffffu, nv, off);
        ns += __shfl_down_sync(0xffffffffu, ns, off);
    }
    if ((tid & 31) == 0) { atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns); }
}

__global__ void __launch_bounds__(STATIC_THREADS, 1) mccfr_static_kernel(SolverDev d, int player, long long n_trav, uint2 pkey,
                                                                         unsigned long long first_trav, StaticDims dm) {
    MS_DYN_SMEM(smem_raw);
    mccfr_static_body(d, player, n_trav, pkey, first_trav, dm, smem_raw);
}

// ------------------------------------------------------------------------------------------------
// K3m  many INDEPENDENT reference-semantics runs in one launch (the reference's experiment protocol,
// run_mccfr_experiment.py:195-202: R runs from an empty table, each its own random stream).  One warp per run: lane 0
// walks (mccfr_tree_traverse<true>, exactly what mccfr_inplace_tree_kernel does for one run), the warp stages and
// writes back the run's table.  Shared memory holds the tree once per CTA and one (regret, strategy, touched, frames)
// set per warp; run r of the launch is bit-identical to a solo ms_mccfr_inplace with philox seed `seed0 + r`.
struct ManyRuns { double* regret; double* strategy; uint8_t* touched; int n_runs; };   // [n_runs][4S], [n_runs][4S], [n_runs][S]

__host__ __device__ inline size_t inplace_many_warp_bytes(int S, int nframes) {
    return (64 * (size_t)S + (size_t)nframes * 26 + (size_t)S + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(256, 1) mccfr_inplace_many_kernel(SolverDev d, ManyRuns m, long long iters, unsigned long long seed0,
                                                                    unsigned long long first_iter, int nframes, int warps) {
    MS_DYN_SMEM(smem_raw);
    const int S = d.n_slots, N = d.n_nodes, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* tree = (uint32_t*)smem_raw;
    unsigned char* wbase = smem_raw + ((4 * (size_t)N + 15) & ~(size_t)15) + (size_t)wid * inplace_many_warp_bytes(S, nframes);
    for (int v = tid; v < N; v += blockDim.x) {
        const int sl = d.node_slot[v];
        uint32_t rec;
        if (sl < 0) rec = ((uint32_t)((int)d.rx2[v] + 2048) & 0xFFFu) | (TREE_TERMINAL << 12);
        else rec = (uint32_t)d.child_begin[v] | ((uint32_t)sl << 12) | ((uint32_t)d.nchild[v] << 23) | ((uint32_t)d.slot_player[sl] << 26);
        tree[v] = rec;
    }
    __syncthreads();
    const int run = blockIdx.x * warps + wid;
    const bool active = wid < warps && run < m.n_runs;      // (block-wide barriers below: no early exit)
    double* reg = (double*)wbase;
    double* str = reg + 4 * S;
    TreeFrames f;
    f.ro = str + 4 * S;
    f.sp = f.ro + nframes;
    f.meta = (uint32_t*)(f.sp + nframes);
    f.cfv = f.meta + nframes;
    f.cb = (uint16_t*)(f.cfv + nframes);
    uint8_t* touched = (uint8_t*)(f.cb + nframes);
    double* greg = m.regret + (size_t)run * 4 * S;
    double* gstr = m.strategy + (size_t)run * 4 * S;
    uint8_t* gtch = m.touched + (size_t)run * S;
    if (active) {
        for (int i = lane; i < 4 * S; i += 32) { reg[i] = greg[i]; str[i] = gstr[i]; }
        for (int i = lane; i < S; i += 32) touched[i] = gtch[i];
    }
    __syncthreads();
    if (active && lane == 0) {
        MccfrShared sh{nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, touched, reg, str};
        const unsigned long long seed = seed0 + (unsigned long long)run;
        const uint2 pkey = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        unsigned long long nu = 0, nv = 0, ns = 0;
        for (long long it = 0; it < iters; it++)
            for (int tp = 0; tp < 2; tp++)
                mccfr_tree_traverse<true>(tree, sh, tp, first_iter + (unsigned long long)it, pkey, f, 1, nu, nv, ns);
        atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns);
    }
    __syncthreads();
    if (!active) return;
    for (int i = lane; i < 4 * S; i += 32) { greg[i] = reg[i]; gstr[i] = str[i]; }
    for (int i = lane; i < S; i += 32) gtch[i] = touched[i];
}

// ------------------------------------------------------------------------------------------------
// Textbook estimators (opt-in; the reference's MCCFRTrainer is the hybrid estimator above).  Same table, same
// frozen-sigma batch semantics, same delta layout -- regret deltas [S][4] plus one scalar per slot that
// multiplies the frozen sigma in the apply step (a visit count for external sampling, a sum of importance
// weights for outcome sampling) -- so ms_mccfr_apply / ms_mccfr_apply_peers serve all three.
//
// External sampling (Lanctot et al. 2009; update rules of open_spiel's external_sampling_mccfr.py with
// AverageType.SIMPLE): the traverser expands every action, the opponent samples one; at traverser nodes
// regret[a] += u_a - sum_b sigma_b u_b, at opponent nodes strategy_sum += sigma.  Child values are expectations
// (doubles), so a frame holds 4 doubles next to the state.
constexpr int ES_THREADS = 512;
constexpr uint32_t MS_TAG_ES = MS_TAG_MCCF + 16u, MS_TAG_OS = MS_TAG_MCCF + 32u;

__host__ __device__ inline size_t es_smem_bytes(int S, int hcap, int nframes, int threads) {
    return sizeof(double) * 12 * (size_t)S + 8 * (size_t)hcap + (size_t)threads * nframes * (16 + 32 + 8) +
           4 * (size_t)S + 2 * (size_t)hcap + (size_t)S + 64;
}

__device__ void es_traverse(const SolverDev& d, const MccfrShared& sh, int tp, unsigned long long trav, uint2 pkey,
                            uint4* f_st, double* f_cv, uint2* f_meta, int fstride, unsigned long long& n_upd,
                            unsigned long long& n_vis, unsigned long long& n_step) {
    MsState s = d.root;
    const uint32_t dealt = dealt_set(d.root);
    int fi = -1;
    uint32_t call = 0;
    double ret = 0.0;
    bool returning = false, pend = false;
    uint32_t pend_a = 0u;
    uint4 xblk = make_uint4(0u, 0u, 0u, 0u);
    uint32_t xblk_id = 0xFFFFFFFFu;
    const uint32_t tag = MS_TAG_ES + (uint32_t)tp;
    while (true) {
        if (!returning) {
            if (pend) { step(s, pend_a, table_set_from_dealt(s, dealt)); n_step++; pend = false; }
            const uint32_t my_call = call++;
            n_vis++;
            if (st_terminal(s)) {
                const int r = reward0_x2(s);
                ret = 0.5 * (double)(tp == 0 ? r : -r);
                returning = true;
                continue;
            }
            const int p = st_cur(s);
            uint32_t list;
            const uint32_t nl = legal_list(s, d.hand_order, p, list);
            const int slot = lookup_slot(sh.hk, sh.hs, sh.hcap, infoset_key(s, p));
            sh.touched[slot] = 1;
            if (p != tp) {                          // opponent: average strategy += sigma, sample, tail call
                atomicAdd(&sh.dcnt[slot], 1u);
                int ai = 0;
                if (nl > 1u) {
                    if ((my_call >> 1) != xblk_id) {
                        xblk_id = my_call >> 1;
                        xblk = philox4x32_10(make_uint4((uint32_t)trav, (uint32_t)(trav >> 32), xblk_id, tag), pkey);
                    }
                    const double u = (my_call & 1u) ? u53(xblk.z, xblk.w) : u53(xblk.x, xblk.y);
                    ai = sample_cdf(sh.cdf + 4 * slot, (int)nl, u);
                }
                pend_a = (list >> (4 * ai)) & 0xFu; pend = true;
                continue;
            }
            fi++;                                   // traverser: expand every action
            const int o = fi * fstride;
            f_st[o] = s;
            f_meta[o] = make_uint2((uint32_t)slot | (nl << 12), list);
            pend_a = list & 0xFu; pend = true;
            continue;
        }
        if (fi < 0) break;
        const int o = fi * fstride;
        uint2 meta = f_meta[o];
        const int slot = (int)(meta.x & 0xFFFu);
        const int nl = (int)((meta.x >> 12) & 0x7u);
        int cur = (int)((meta.x >> 16) & 0x7u);
        f_cv[(size_t)o * 4 + cur] = ret;
        cur++;
        if (cur < nl) {
            f_meta[o] = make_uint2((meta.x & 0xFFFFu) | ((uint32_t)cur << 16), meta.y);
            s = f_st[o];
            pend_a = (meta.y >> (4 * cur)) & 0xFu; pend = true;
            returning = false;
            continue;
        }
        double cv[4], value = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cv[i] = (i < nl) ? f_cv[(size_t)o * 4 + i] : 0.0;
            if (i < nl) value = __dadd_rn(value, __dmul_rn(sh.sig[4 * slot + i], cv[i]));
        }
        if (nl > 1) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < nl) atomicAdd(&sh.dreg[4 * slot + i], __dadd_rn(cv[i], -value));
        }
        n_upd++;
        ret = value;
        fi--;
        returning = true;
    }
}

__global__ void __launch_bounds__(ES_THREADS, 1) mccfr_es_kernel(SolverDev d, int player, long long n_trav, uint2 pkey,
                                                                 unsigned long long first_trav, int nframes) {
    MS_DYN_SMEM(smem_raw);
    const int S = d.n_slots, T = blockDim.x, tid = threadIdx.x;
    double* sig = (double*)smem_raw;
    double* cdf = sig + 4 * S;
    double* dreg = cdf + 4 * S;
    unsigned long long* hk = (unsigned long long*)(dreg + 4 * S);
    uint4* f_st = (uint4*)(hk + d.hcap);
    double* f_cv = (double*)(f_st + (size_t)T * nframes);
    uint2* f_meta = (uint2*)(f_cv + (size_t)T * nframes * 4);
    uint32_t* dcnt = (uint32_t*)(f_meta + (size_t)T * nframes);
    int16_t* hs = (int16_t*)(dcnt + S);
    uint8_t* touched = (uint8_t*)(hs + d.hcap);
    for (int s = tid; s < S; s += T) {
        double reg[4], sg[4], cd[4];
        for (int i = 0; i < 4; i++) reg[i] = d.regret[4 * s + i];
        regret_match(reg, d.slot_nlegal[s], sg);
        strategy_cdf(sg, d.slot_nlegal[s], cd);
        for (int i = 0; i < 4; i++) { sig[4 * s + i] = sg[i]; cdf[4 * s + i] = cd[i]; dreg[4 * s + i] = 0.0; }
        dcnt[s] = 0u; touched[s] = 0;
    }
    for (int i = tid; i < d.hcap; i += T) { hk[i] = d.hkeys[i]; hs[i] = d.hslots[i]; }
    __syncthreads();
    MccfrShared sh{hk, hs, d.hcap, sig, cdf, dreg, dcnt, touched, nullptr, nullptr};
    unsigned long long nu = 0, nv = 0, ns = 0;
    for (long long k = blockIdx.x * (long long)T + tid; k < n_trav; k += (long long)gridDim.x * T)
        for (int tp = 0; tp < 2; tp++) {
            if (player < 2 && tp != player) continue;
            es_traverse(d, sh, tp, first_trav + (unsigned long long)k, pkey, f_st + tid, f_cv + (size_t)tid * 4, f_meta + tid,
                        T, nu, nv, ns);
        }
    __syncthreads();
    for (int i = tid; i < 4 * S; i += T) { const double v = dreg[i]; if (v != 0.0) atomicAdd(&d.delta[i], v); }
    for (int s = tid; s < S; s += T) {
        if (dcnt[s]) atomicAdd(&d.delta[4 * S + s], (double)dcnt[s]);
        if (touched[s] && !d.touched[s]) atomicAdd(&d.delta[5 * S + s], 1.0);   // first touch: travels with the delta
    }
    for (int off = 16; off > 0; off >>= 1) {
        nu += __shfl_down_sync(0xffffffffu, nu, off);
        nv += __shfl_down_sync(0xffffffffu, nv, off);
        ns += __shfl_down_sync(0xffffffffu, ns, off);
    }
    if ((tid & 31) == 0) { atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns); }
}

// Outcome sampling (update rules of open_spiel's outcome_sampling_mccfr.py: epsilon-on-policy exploration at the
// traverser's nodes with epsilon = 0.6, baseline 0): one trajectory per traversal, no branching, so the per-depth
// records live in registers / local memory and the updates are replayed on the way back.
#define MS_OS_EPSILON 0.6
__device__ void os_traverse(const SolverDev& d, const MccfrShared& sh, double* dwt, int tp, unsigned long long trav,
                            uint2 pkey, unsigned long long& n_upd, unsigned long long& n_vis, unsigned long long& n_step) {
    MsState s = d.root;
    const uint32_t dealt = dealt_set(d.root);
    const uint32_t tag = MS_TAG_OS + (uint32_t)tp;
    double my_reach = 1.0, opp_reach = 1.0, sample_reach = 1.0;
    int rslot[16], rmeta[16];                 // per ply: slot; n_legal | sampled index << 4 | traverser flag << 8
    double rsig[16], rsp[16], rw[16], rmw[16];
    int depth = 0;
    uint32_t call = 0;
    while (!st_terminal(s) && depth < 16) {
        const uint32_t my_call = call++;
        n_vis++;
        const int p = st_cur(s);
        uint32_t list;
        const uint32_t nl = legal_list(s, d.hand_order, p, list);
        const int slot = lookup_slot(sh.hk, sh.hs, sh.hcap, infoset_key(s, p));
        sh.touched[slot] = 1;
        double sg[4], sp[4];
        const double uni = __ddiv_rn(1.0, (double)nl);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            sg[i] = sh.sig[4 * slot + i];
            sp[i] = (p == tp) ? __dadd_rn(__dmul_rn(MS_OS_EPSILON, uni), __dmul_rn(1.0 - MS_OS_EPSILON, sg[i])) : sg[i];
        }
        const uint4 x = philox4x32_10(make_uint4((uint32_t)trav, (uint32_t)(trav >> 32), my_call >> 1, tag), pkey);
        const double u = (my_call & 1u) ? u53(x.z, x.w) : u53(x.x, x.y);
        const int ai = sample_action(sp, (int)nl, u);
        rslot[depth] = slot;
        rmeta[depth] = (int)nl | (ai << 4) | ((p == tp) ? 256 : 0);
        rsig[depth] = sg[ai]; rsp[depth] = sp[ai];
        rw[depth] = __ddiv_rn(opp_reach, sample_reach);
        rmw[depth] = __ddiv_rn(my_reach, sample_reach);
        if (p == tp) my_reach = __dmul_rn(my_reach, sg[ai]); else opp_reach = __dmul_rn(opp_reach, sg[ai]);
        sample_reach = __dmul_rn(sample_reach, sp[ai]);
        step(s, (list >> (4 * ai)) & 0xFu, table_set_from_dealt(s, dealt)); n_step++;
        depth++;
    }
    n_vis++;                                   // the terminal call
    const int r = st_terminal(s) ? reward0_x2(s) : 0;
    double value = 0.5 * (double)(tp == 0 ? r : -r);
    for (int k = depth - 1; k >= 0; k--) {
        const int nl = rmeta[k] & 15, ai = (rmeta[k] >> 4) & 15, slot = rslot[k];
        const double est = __ddiv_rn(value, rsp[k]);
        const double ve = __dmul_rn(rsig[k], est);
        if (rmeta[k] & 256) {
            const double cfv = __dmul_rn(ve, rw[k]);
            for (int i = 0; i < nl; i++) {
                const double cfa = __dmul_rn(i == ai ? est : 0.0, rw[k]);
                atomicAdd(&sh.dreg[4 * slot + i], __dadd_rn(cfa, -cfv));
            }
            atomicAdd(&dwt[slot], rmw[k]);     // strategy_sum += (my_reach / sample_reach) * sigma, sigma applied later
            n_upd++;
        }
        value = ve;
    }
}

// External sampling walking the enumerated tree (the form ms_mccfr_batch_mode(mode = 1) launches; mccfr_es_kernel
// above, which re-steps the env, is kept for trees that do not fit).  Same update rules, same Philox addressing, same
// node records, lane-indexed delta-table copies and frozen-sigma semantics as mccfr_tree_kernel.  A traverser node
// with one legal action needs no frame: its value is its child's (1.0 * x = x exactly) and its regret delta is 0.
struct EsTreeFrames { double* cv; uint32_t* meta; uint16_t* cb; };

__host__ __device__ inline size_t es_tree_smem(int S, int n_nodes, int nframes, int threads, int ncopy) {
    return sizeof(double) * (7 + 4 * (size_t)ncopy) * (size_t)S + 4 * (size_t)n_nodes + 4 * (size_t)S + (size_t)S +
           (size_t)threads * nframes * (32 + 4 + 2) + 64;
}

__device__ void es_tree_traverse(const uint32_t* __restrict__ tree, const MccfrShared& sh, int tp, unsigned long long trav,
                                 uint2 pkey, const EsTreeFrames& f, int fstride, unsigned long long& n_upd,
                                 unsigned long long& n_vis, unsigned long long& n_step) {
    uint32_t node = 0u, call = 0u;
    int fi = -1;
    double ret = 0.0;
    bool returning = false;
    uint4 xblk = make_uint4(0u, 0u, 0u, 0u);
    uint32_t xblk_id = 0xFFFFFFFFu;
    const uint32_t tag = MS_TAG_ES + (uint32_t)tp;
    while (true) {
        if (!returning) {
            const uint32_t rec = tree[node];
            const uint32_t my_call = call++;
            n_vis++;
            const uint32_t slot = (rec >> 12) & 0x7FFu;
            if (slot == TREE_TERMINAL) {
                const int r = (int)(rec & 0xFFFu) - 2048;
                ret = 0.5 * (double)(tp == 0 ? r : -r);
                returning = true;
                continue;
            }
            const uint32_t nl = (rec >> 23) & 0x7u, cb = rec & 0xFFFu;
            const int p = (int)((rec >> 26) & 1u);
            sh.touched[slot] = 1;
            n_step++;
            if (p != tp) {                          // opponent: average strategy += sigma, sample, tail call
                atomicAdd(&sh.dcnt[slot], 1u);
                int ai = 0;
                if (nl > 1u) {
                    if ((my_call >> 1) != xblk_id) {
                        xblk_id = my_call >> 1;
                        xblk = philox4x32_10(make_uint4((uint32_t)trav, (uint32_t)(trav >> 32), xblk_id, tag), pkey);
                    }
                    const double u = (my_call & 1u) ? u53(xblk.z, xblk.w) : u53(xblk.x, xblk.y);
                    const double* cd = sh.cdf + 3 * slot;
#pragma unroll
                    for (int i = 0; i < 3; i++)
                        if ((uint32_t)i + 1u < nl) ai += (cd[i] <= u) ? 1 : 0;
                }
                node = cb + (uint32_t)ai;
                continue;
            }
            if (nl == 1u) { n_upd++; node = cb; continue; }      // forced traverser move: value = the child's value
            fi++;                                   // traverser: expand every action
            const int o = fi * fstride;
            f.meta[o] = slot | (nl << 11);          // | cursor << 14
            f.cb[o] = (uint16_t)cb;
            node = cb;
            continue;
        }
        if (fi < 0) break;
        const int o = fi * fstride;
        uint32_t meta = f.meta[o];
        const int slot = (int)(meta & 0x7FFu);
        const int nl = (int)((meta >> 11) & 0x7u);
        int cur = (int)((meta >> 14) & 0x7u);
        f.cv[(size_t)(4 * fi + cur) * fstride] = ret;
        cur++;
        if (cur < nl) {
            f.meta[o] = (meta & ~(0x7u << 14)) | ((uint32_t)cur << 14);
            node = (uint32_t)f.cb[o] + (uint32_t)cur;
            n_step++;
            returning = false;
            continue;
        }
        double cv[4], value = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            cv[i] = (i < nl) ? f.cv[(size_t)(4 * fi + i) * fstride] : 0.0;
            if (i < nl) value = __dadd_rn(value, __dmul_rn(sh.sig[4 * slot + i], cv[i]));
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < nl) atomicAdd(&sh.dreg[4 * slot + i], __dadd_rn(cv[i], -value));
        n_upd++;
        ret = value;
        fi--;
        returning = true;
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) mccfr_es_tree_kernel(SolverDev d, int player, long long n_trav, uint2 pkey,
                                                                   unsigned long long first_trav, int nframes, int ncopy) {
    MS_DYN_SMEM(smem_raw);
    const int S = d.n_slots, N = d.n_nodes, T = THREADS, tid = threadIdx.x;
    double* sig = (double*)smem_raw;
    double* cdf = sig + 4 * S;              // [S][3]
    double* dreg = cdf + 3 * S;             // ncopy copies, chosen by lane id
    EsTreeFrames f;
    f.cv = dreg + (size_t)ncopy * 4 * S;
    f.meta = (uint32_t*)(f.cv + (size_t)T * nframes * 4);
    uint32_t* tree = f.meta + (size_t)T * nframes;
    uint32_t* dcnt = tree + N;
    f.cb = (uint16_t*)(dcnt + S);
    uint8_t* touched = (uint8_t*)(f.cb + (size_t)T * nframes);
    for (int s = tid; s < S; s += T) {
        double reg[4], sg[4], cd[4];
        for (int i = 0; i < 4; i++) reg[i] = d.regret[4 * s + i];
        regret_match(reg, d.slot_nlegal[s], sg);
        strategy_cdf(sg, d.slot_nlegal[s], cd);
        for (int i = 0; i < 4; i++) sig[4 * s + i] = sg[i];
        for (int i = 0; i < 3; i++) cdf[3 * s + i] = cd[i];
        dcnt[s] = 0u; touched[s] = 0;
    }
    for (int i = tid; i < ncopy * 4 * S; i += T) dreg[i] = 0.0;
    for (int v = tid; v < N; v += T) {
        const int sl = d.node_slot[v];
        uint32_t rec;
        if (sl < 0) rec = ((uint32_t)((int)d.rx2[v] + 2048) & 0xFFFu) | (TREE_TERMINAL << 12);
        else rec = (uint32_t)d.child_begin[v] | ((uint32_t)sl << 12) | ((uint32_t)d.nchild[v] << 23) | ((uint32_t)d.slot_player[sl] << 26);
        tree[v] = rec;
    }
    __syncthreads();
    MccfrShared sh{nullptr, nullptr, 0, sig, cdf, dreg + (size_t)(tid & (ncopy - 1)) * 4 * S, dcnt, touched, nullptr, nullptr};
    f.cv += tid; f.meta += tid; f.cb += tid;
    unsigned long long nu = 0, nv = 0, ns = 0;
    const int flip = (tid >> 5) & 1;
    for (long long k = blockIdx.x * (long long)T + tid; k < n_trav; k += (long long)gridDim.x * T)
        for (int j = 0; j < 2; j++) {
            const int tp = j ^ flip;
            if (player < 2 && tp != player) continue;
            es_tree_traverse(tree, sh, tp, first_trav + (unsigned long long)k, pkey, f, T, nu, nv, ns);
        }
    __syncthreads();
    for (int i = tid; i < 4 * S; i += T) {
        double v = dreg[i];
        for (int c = 1; c < ncopy; c++) v = __dadd_rn(v, dreg[(size_t)c * 4 * S + i]);
        if (v != 0.0) atomicAdd(&d.delta[i], v);
    }
    for (int s = tid; s < S; s += T) {
        if (dcnt[s]) atomicAdd(&d.delta[4 * S + s], (double)dcnt[s]);
        if (touched[s] && !d.touched[s]) atomicAdd(&d.delta[5 * S + s], 1.0);   // first touch: travels with the delta
    }
    for (int off = 16; off > 0; off >>= 1) {
        nu += __shfl_down_sync(0xffffffffu, nu, off);
        nv += __shfl_down_sync(0xffffffffu, nv, off);
        ns += __shfl_down_sync(0xffffffffu, ns, off);
    }
    if ((tid & 31) == 0) { atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns); }
}

__global__ void __launch_bounds__(256) mccfr_os_kernel(SolverDev d, int player, long long n_trav, uint2 pkey,
                                                       unsigned long long first_trav) {
    MS_DYN_SMEM(smem_raw);
    const int S = d.n_slots, T = blockDim.x, tid = threadIdx.x;
    double* sig = (double*)smem_raw;
    double* dreg = sig + 4 * S;
    double* dwt = dreg + 4 * S;
    unsigned long long* hk = (unsigned long long*)(dwt + S);
    int16_t* hs = (int16_t*)(hk + d.hcap);
    uint8_t* touched = (uint8_t*)(hs + d.hcap);
    for (int s = tid; s < S; s += T) {
        double reg[4], sg[4];
        for (int i = 0; i < 4; i++) reg[i] = d.regret[4 * s + i];
        regret_match(reg, d.slot_nlegal[s], sg);
        for (int i = 0; i < 4; i++) { sig[4 * s + i] = sg[i]; dreg[4 * s + i] = 0.0; }
        dwt[s] = 0.0; touched[s] = 0;
    }
    for (int i = tid; i < d.hcap; i += T) { hk[i] = d.hkeys[i]; hs[i] = d.hslots[i]; }
    __syncthreads();
    MccfrShared sh{hk, hs, d.hcap, sig, nullptr, dreg, nullptr, touched, nullptr, nullptr};
    unsigned long long nu = 0, nv = 0, ns = 0;
    for (long long k = blockIdx.x * (long long)T + tid; k < n_trav; k += (long long)gridDim.x * T)
        for (int tp = 0; tp < 2; tp++) {
            if (player < 2 && tp != player) continue;
            os_traverse(d, sh, dwt, tp, first_trav + (unsigned long long)k, pkey, nu, nv, ns);
        }
    __syncthreads();
    for (int i = tid; i < 4 * S; i += T) { const double v = dreg[i]; if (v != 0.0) atomicAdd(&d.delta[i], v); }
    for (int s = tid; s < S; s += T) {
        if (dwt[s] != 0.0) atomicAdd(&d.delta[4 * S + s], dwt[s]);
        if (touched[s] && !d.touched[s]) atomicAdd(&d.delta[5 * S + s], 1.0);   // first touch: travels with the delta
    }
    for (int off = 16; off > 0; off >>= 1) {
        nu += __shfl_down_sync(0xffffffffu, nu, off);
        nv += __shfl_down_sync(0xffffffffu, nv, off);
        ns += __shfl_down_sync(0xffffffffu, ns, off);
    }
    if ((tid & 31) == 0) { atomicAdd(&d.counters[0], nu); atomicAdd(&d.counters[1], nv); atomicAdd(&d.counters[2], ns); }
}

// table += delta; delta = 0.  strategy_sum += count * sigma with sigma = RM(regret BEFORE the update),
// i.e. the strategy the batch was sampled with.
__global__ void __launch_bounds__(256) mccfr_apply_kernel(SolverDev d) {
    const int S = d.n_slots;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        double reg[4], sg[4];
        for (int i = 0; i < 4; i++) reg[i] = d.regret[4 * s + i];
        const int n = d.slot_nlegal[s];
        const double cnt = d.delta[4 * S + s];
        if (cnt != 0.0) {
            regret_match(reg, n, sg);
            for (int i = 0; i < n; i++) d.strategy[4 * s + i] = __dadd_rn(d.strategy[4 * s + i], __dmul_rn(cnt, sg[i]));
        }
        for (int i = 0; i < 4; i++) {
            const double dv = d.delta[4 * s + i];
            if (dv != 0.0) d.regret[4 * s + i] = __dadd_rn(reg[i], dv);
            d.delta[4 * s + i] = 0.0;
        }
        d.delta[4 * S + s] = 0.0;
        // first touch of the infoset by ANY rank's traversals (the mark is summed with the rest of the delta): the
        // reference creates the InfoNode on first touch (mc_cfr.py:52), and every replica must agree on which exist
        if (d.delta[5 * S + s] != 0.0) { d.touched[s] = 1; d.delta[5 * S + s] = 0.0; }
    }
}

// Multi-GPU exchange without a library collective, PUSH form.  Every rank owns an inbox of [world][6 S] doubles per
// iteration parity, mapped into every peer (CUDA IPC over NVLink / NVSwitch).  One CTA per rank:
//   1. push: copies this rank's delta buffer into slot `rank` of EVERY rank's inbox (plain stores over peer memory:
//      fire-and-forget, bandwidth-bound -- a pull would pay one NVLink round trip per dependent load) and zeroes it;
//   2. barrier: system-scope fence, then the iteration number into its slot of every peer's flag array
//      (st.release.sys), and spins (bounded) until all slots of its own array have reached it (ld.acquire.sys):
//      a rank that sees peer r's flag sees r's pushed deltas;
//   3. sums its own inbox in rank order -- so all replicas compute the same bits -- and applies the sum to its table.
// Inboxes are double buffered by iteration parity: a peer can be at most one iteration ahead (it cannot pass the
// barrier of iteration i + 1 before this rank has signalled it, which this rank does after reading inbox i).
// Measured at 8 GPUs (profiles/README.md): the pull form cost 23 us per iteration (8 dependent rounds of remote loads),
// NCCL all-reduce + apply 36 us.
constexpr int MS_MAX_PEERS = 8;
struct PeerView {
    double* inbox[MS_MAX_PEERS];                // rank r's inbox of the current parity, [world][6 S] (inbox[rank] is local)
    unsigned long long* flags[MS_MAX_PEERS];    // peers' flag arrays ([world] u64 each)
    unsigned long long* my_flags;
    int rank, world;
};

// The barrier's two memory operations, behind inline functions so that tests/emu can run the kernel with two emulated
// "ranks" in one process (host build: C++ atomics with the same ordering).
#ifndef MS_CTA_EMU
__device__ __forceinline__ void peer_signal(unsigned long long* flag, unsigned long long epoch) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(epoch) : "memory");
}
__device__ __forceinline__ unsigned long long peer_poll(const unsigned long long* flag) {
    unsigned long long seen;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
    return seen;
}
__device__ __forceinline__ unsigned long long peer_clock_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double peer_load(const double* p) { return __ldcv(p); }      // written by a peer: not from L1
__device__ __forceinline__ void peer_store(double* p, double v) { __stcg(p, v); }
__device__ __forceinline__ void peer_fence() { __threadfence_system(); }
#endif

constexpr unsigned long long MS_PEER_TIMEOUT_NS = 2000000000ull;     // a peer that has not arrived after 2 s is given up on

// err[0]: 0 = fine; otherwise (1 + the first peer that did not arrive) of the first failed exchange.  Once set, every
// later exchange returns at once without touching the table (ms_solver_peer_error reports MS_ERR_STATE).
// One CTA (all its threads) runs this: as the whole of mccfr_apply_peers_kernel, or as the tail of the fused
// mccfr_static_peers_kernel in the last CTA to finish its traversals.  `s_bad` is one int of shared memory, `s_sum`
// [6 S] doubles of it.
__device__ __forceinline__ void peers_exchange_cta(const SolverDev& d, const PeerView& pv, unsigned long long epoch, unsigned int* err,
                                                   volatile int* s_bad, double* s_sum) {
    const int tid = threadIdx.x, S = d.n_slots;
    if (tid == 0) *s_bad = (*(volatile unsigned int*)err != 0u) ? 1 : 0;
    __syncthreads();
    if (*s_bad) return;
    // 1. push this rank's deltas into every rank's inbox (its own included), and clear them for the next batch
    const int n6 = 6 * S;
    for (int i = tid; i < n6; i += blockDim.x) {
        const double v = __ldcg(d.delta + i);               // accumulated by other CTAs' atomics: read at L2
        d.delta[i] = 0.0;
        for (int r = 0; r < pv.world; r++) peer_store(pv.inbox[r] + (size_t)pv.rank * n6 + i, v);
    }
    __syncthreads();                                        // all pushes of the CTA happen-before the signalling threads ...
    // 2. barrier
    if (tid < pv.world) {
        peer_fence();                                       // ... whose system-scope fence + release store publish them (cumulativity)
        peer_signal(pv.flags[tid] + pv.rank, epoch);
        const unsigned long long t0 = peer_clock_ns();
        while (peer_poll(pv.my_flags + tid) < epoch) {
            if (peer_clock_ns() - t0 > MS_PEER_TIMEOUT_NS) { atomicCAS(err, 0u, 1u + (unsigned)tid); *s_bad = 1; break; }
        }
    }
    __syncthreads();
    if (*s_bad) return;
    // 3. rank-ordered sum of the own inbox into shared memory (per element the loads of all ranks are independent and
    //    issued together: one L2 round trip per element instead of one per rank), then the table update
    const double* in = pv.inbox[pv.rank];
    for (int i = tid; i < n6; i += blockDim.x) {
        double v[MS_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < MS_MAX_PEERS; r++) v[r] = r < pv.world ? peer_load(in + (size_t)r * n6 + i) : 0.0;
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < MS_MAX_PEERS; r++) if (r < pv.world) t = __dadd_rn(t, v[r]);
        s_sum[i] = t;
    }
    __syncthreads();
    for (int s = tid; s < S; s += blockDim.x) {
        const double cnt = s_sum[4 * S + s], tch = s_sum[5 * S + s];
        double reg[4], sg[4];
        for (int i = 0; i < 4; i++) reg[i] = d.regret[4 * s + i];
        const int n = d.slot_nlegal[s];
        if (cnt != 0.0) {
            regret_match(reg, n, sg);
            for (int i = 0; i < n; i++) d.strategy[4 * s + i] = __dadd_rn(d.strategy[4 * s + i], __dmul_rn(cnt, sg[i]));
        }
        for (int i = 0; i < 4; i++) {
            const double dv = s_sum[4 * s + i];
            if (dv != 0.0) d.regret[4 * s + i] = __dadd_rn(reg[i], dv);
        }
        if (tch != 0.0) d.touched[s] = 1;
    }
}

__global__ void __launch_bounds__(1024, 1) mccfr_apply_peers_kernel(SolverDev d, PeerView pv, unsigned long long epoch,
                                                                    unsigned int* err) {
    MS_DYN_SMEM(smem_raw);                  // 16 bytes of flags + [6 S] doubles (dynamic: emulated ranks = blocks must not share it)
    peers_exchange_cta(d, pv, epoch, err, (volatile int*)smem_raw, (double*)(smem_raw + 16));
}

// The fused form (one launch per MCCFR iteration per GPU): the traversals of mccfr_static_kernel, and in the LAST CTA to
// finish them (a ticket counter in global memory) the cross-GPU exchange + table update of peers_exchange_cta.  The
// deltas of a batch are complete only when its last traversal is, so the exchange cannot start earlier; what the fusion
// removes is the launch gap and the second kernel's ramp between the two.  Each CTA's delta flush is a set of device-scope
// atomics followed by __threadfence() and the ticket increment; the last CTA observes every ticket, fences at system
// scope (peers_exchange_cta) and only then signals its peers, so a peer that sees the flag sees the deltas.
// Measured (profiles/fused_probe.py, one GPU, world = 1, 12 waves): {batch, apply} 469 us, {batch, apply_peers} 477 us,
// this kernel 502 us -- with the exchange code inlined behind them the traversal loops themselves run 5 % slower (ptxas
// allocates the same 64 registers differently); an out-of-line tail was worse (the parameter structs then live in local
// memory: 512 us).  The headline therefore uses two launches; this form is kept, tested and reported beside it.
__global__ void __launch_bounds__(STATIC_THREADS, 1) mccfr_static_peers_kernel(SolverDev d, int player, long long n_trav, uint2 pkey,
                                                                               unsigned long long first_trav, StaticDims dm, PeerView pv,
                                                                               unsigned long long epoch, unsigned int* err,
                                                                               unsigned int* ticket) {
    MS_DYN_SMEM(smem_raw);
    mccfr_static_body(d, player, n_trav, pkey, first_trav, dm, smem_raw);
    volatile int* s_flag = (volatile int*)smem_raw;          // the traversal working set is dead from here on
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        s_flag[1] = (t == gridDim.x - 1u) ? 1 : 0;
        if (s_flag[1]) *ticket = 0u;                         // ready for the next launch (stream-ordered)
    }
    __syncthreads();
    if (!s_flag[1]) return;
    __threadfence();
    peers_exchange_cta(d, pv, epoch, err, s_flag, (double*)(smem_raw + 16));
}

// ------------------------------------------------------------------------------------------------
// Best response against the table's average policy (restated open_spiel BestResponsePolicy, see
// oracle/ms_exploit.py for the algorithm; third-party, parity unpinned).  Same level-synchronous
// skeleton as CFR: counterfactual reach top-down, then bottom-up values with one argmax per
// infoset of the best responder.
__device__ __forceinline__ void avg_policy(const SolverDev& d, int s, int kind, double* out) {
    const int n = d.slot_nlegal[s];
    double st[4], tot = 0.0;
    for (int i = 0; i < 4; i++) { st[i] = (i < n) ? d.strategy[4 * s + i] : 0.0; if (i < n) tot = __dadd_rn(tot, st[i]); }
    bool use;
    if (kind == 0) use = tot > 0.0;                                   // LearnedCFRPolicy / InfoNode.policy
    else if (kind == 1) use = d.touched[s] && tot > 1e-12;            // ScopaLearnedPolicy (unseen key -> uniform)
    else use = false;
    for (int i = 0; i < 4; i++) out[i] = (i < n) ? (use ? ms_ddiv_or_zero(st[i], tot) : __ddiv_rn(1.0, (double)n)) : 0.0;
}

__global__ void __launch_bounds__(512, 1) best_response_kernel(SolverDev d, int n_dec, int kind, double* out2) {
    MS_DYN_SMEM(smem_raw);
    __shared__ int s_lvl[MAXL + 1], s_slvl[MAXL + 1];
    const int tid = threadIdx.x, bd = blockDim.x;
    const int N = d.n_nodes, S = d.n_slots, L = d.n_levels;
    CfrSmem m = cfr_carve(smem_raw, N, S, n_dec);   // reg is unused; sig holds the average policy
    for (int i = tid; i <= L; i += bd) { s_lvl[i] = d.level_begin[i]; s_slvl[i] = d.slot_level_begin[i]; }
    for (int i = tid; i < N; i += bd) {
        m.child_begin[i] = d.child_begin[i]; m.nchild[i] = d.nchild[i];
        m.node_slot[i] = d.node_slot[i]; m.rx2[i] = d.rx2[i];
    }
    for (int i = tid; i < n_dec; i += bd) m.chain_nodes[i] = d.chain_nodes[i];
    for (int i = tid; i <= S; i += bd) m.chain_begin[i] = d.chain_begin[i];
    for (int s = tid; s < S; s += bd) {
        m.nlegal[s] = d.slot_nlegal[s];
        double p[4];
        avg_policy(d, s, kind, p);
        for (int i = 0; i < 4; i++) m.sig[4 * s + i] = p[i];
    }
    __syncthreads();
    for (int b = 0; b < 2; b++) {
        if (tid == 0) m.r[0] = 1.0;
        __syncthreads();
        for (int l = 0; l + 1 < L; l++) {
            const int cp = (d.root_cur + l) & 1;
            for (int v = s_lvl[l] + tid; v < s_lvl[l + 1]; v += bd) {
                const int nc = m.nchild[v];
                if (nc == 0) continue;
                const int cb = m.child_begin[v];
                const double cf = m.r[v];
                const double* pol = m.sig + 4 * m.node_slot[v];
                for (int i = 0; i < nc; i++) m.r[cb + i] = (cp == b) ? cf : __dmul_rn(cf, pol[i]);
            }
            __syncthreads();
        }
        for (int l = L - 1; l >= 0; l--) {
            const int cp = (d.root_cur + l) & 1;
            for (int v = s_lvl[l] + tid; v < s_lvl[l + 1]; v += bd) {
                const int nc = m.nchild[v];
                if (nc == 0) {
                    const int rr = m.rx2[v];
                    m.u[v] = 0.5 * (double)(b == 0 ? rr : -rr);
                } else if (cp != b) {
                    const int cb = m.child_begin[v];
                    const double* pol = m.sig + 4 * m.node_slot[v];
                    double acc = 0.0;
                    for (int i = 0; i < nc; i++)
                        if (pol[i] > 0.0) acc = __dadd_rn(acc, __dmul_rn(pol[i], m.u[cb + i]));
                    m.u[v] = acc;
                }
            }
            if (cp == b) {
                for (int s = s_slvl[l] + tid; s < s_slvl[l + 1]; s += bd) {
                    const int n = m.nlegal[s];
                    int best = 0; double bestq = 0.0;
                    for (int a = 0; a < n; a++) {
                        double q = 0.0;
                        for (int k = m.chain_begin[s]; k < m.chain_begin[s + 1]; k++) {
                            const int v = m.chain_nodes[k];
                            q = __dadd_rn(q, __dmul_rn(m.r[v], m.u[m.child_begin[v] + a]));
                        }
                        if (a == 0 || q > bestq) { best = a; bestq = q; }   // first maximum in legal order
                    }
                    for (int k = m.chain_begin[s]; k < m.chain_begin[s + 1]; k++) {
                        const int v = m.chain_nodes[k];
                        m.u[v] = m.u[m.child_begin[v] + best];
                    }
                }
            }
            __syncthreads();
        }
        if (tid == 0) out2[b] = m.u[0];
        __syncthreads();
    }
}

// the table's average policy per slot (probabilities over the legal actions in hand order)
__global__ void __launch_bounds__(256) policy_kernel(SolverDev d, int kind, double* out) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < d.n_slots; s += gridDim.x * blockDim.x) {
        double p[4];
        avg_policy(d, s, kind, p);
        for (int i = 0; i < 4; i++) out[4 * s + i] = p[i];
    }
}

// Batched evaluate_agent (vanilla_cfr.py:157-216 / mc_cfr.py:146-206): one thread plays one episode from
// the root, seat 0 acting with pol0 and seat 1 with pol1 (per-slot probabilities over the legal actions in
// hand order), sampling with numpy's rule from the Philox "EVAL" stream (ctr = episode id, ply/2).
#define MS_TAG_EVAL 0x4C415645u
__global__ void __launch_bounds__(256) eval_kernel(SolverDev d, const double* __restrict__ pol0,
                                                   const double* __restrict__ pol1, long long n, uint2 pkey,
                                                   unsigned long long first, float* __restrict__ out_reward0,
                                                   uchar2* __restrict__ scopas) {
    const uint32_t dealt = dealt_set(d.root);
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        MsState s = d.root;
        const unsigned long long eid = first + (unsigned long long)g;
        uint4 xb = make_uint4(0u, 0u, 0u, 0u);
        for (uint32_t ply = 0; ply < 32u && !st_terminal(s); ply++) {
            const int p = st_cur(s);
            uint32_t list;
            const uint32_t nl = legal_list(s, d.hand_order, p, list);
            uint32_t ai = 0u;
            if (nl > 1u) {
                const int slot = lookup_slot(d.hkeys, d.hslots, d.hcap, infoset_key(s, p));
                const double* pr = (p == 0 ? pol0 : pol1) + 4 * slot;
                if ((ply & 1u) == 0u)
                    xb = philox4x32_10(make_uint4((uint32_t)eid, (uint32_t)(eid >> 32), ply >> 1, MS_TAG_EVAL), pkey);
                const double u = (ply & 1u) ? u53(xb.z, xb.w) : u53(xb.x, xb.y);
                double sg[4];
                for (int i = 0; i < 4; i++) sg[i] = pr[i];
                ai = (uint32_t)sample_action(sg, (int)nl, u);
            }
            step(s, (list >> (4u * ai)) & 0xFu, table_set_from_dealt(s, dealt));
        }
        if (out_reward0) out_reward0[g] = st_terminal(s) ? reward0(s) : 0.f;
        if (scopas) scopas[g] = make_uchar2((unsigned char)st_scopas(s, 0), (unsigned char)st_scopas(s, 1));
    }
}

}  // namespace ms

#ifndef MS_HOST_RULES_ONLY   // tests/emu/ms_solver_host.cpp compiles every kernel above (but the peer exchange) for the
                             // host's CTA emulator; below: the library's host side (CUDA runtime calls, launches, C ABI)
using namespace ms;

// ------------------------------------------------------------------------------------------------
struct ms_solver {
    int device = 0;
    ms_state root{};
    uint32_t hand_order = 0;
    int n_nodes = 0, n_levels = 0, n_slots = 0, n_dec = 0, hcap = 0, nframes = 4, nframes_tree = 4, nframes_es = 4;
    bool static_shape = false;                 // the tree of a fresh 4+4-card deal (ms_static_walk.cuh)
    StaticDims sdm{};
    std::vector<int> level_begin, slot_level_begin;
    // host copies for export
    std::vector<ms_state> h_state; std::vector<int> h_parent, h_child_begin, h_slot; std::vector<uint8_t> h_nchild, h_level;
    std::vector<unsigned long long> h_slot_key; std::vector<uint8_t> h_slot_nlegal, h_slot_legal;
    char* d_block = nullptr;   // one allocation for everything
    SolverDev dev{};
    double* d_value = nullptr; // [2] scratch for returned values
    // peer-memory exchange (ms_solver_ipc_export / _attach / ms_mccfr_apply_peers)
    double* inbox[2] = {nullptr, nullptr};     // [MS_MAX_PEERS][6 S] per iteration parity: where the peers push their deltas
    unsigned long long* flags = nullptr;
    unsigned int* peer_err = nullptr;          // device word set by the peer exchange when a peer did not arrive
    int rank = 0, world = 1, parity = 0;
    unsigned long long epoch = 0;
    bool attached = false;
    void* peer_base[MS_MAX_PEERS] = {};
    uint64_t peer_off[MS_MAX_PEERS][3] = {};
};

namespace {

template <typename T>
T* carve(char*& p, size_t n) {
    T* r = (T*)p;
    p += (n * sizeof(T) + 255) & ~(size_t)255;
    return r;
}

int solver_build(ms_solver* sv) {
    // ---- 1. enumerate the tree on the device
    char* tmp = nullptr;
    size_t tb = (size_t)MAXN * (16 + 4 + 4 + 1 + 8 + 2 + 1) + 4096;
    MS_CUDA(cudaMalloc(&tmp, tb + 4096));
    char* p = tmp;
    TreeOut t;
    t.state = carve<uint4>(p, MAXN); t.parent = carve<int>(p, MAXN); t.child_begin = carve<int>(p, MAXN);
    t.key = carve<unsigned long long>(p, MAXN); t.legal = carve<uint16_t>(p, MAXN);
    t.nchild = carve<uint8_t>(p, MAXN); t.rx2 = carve<int8_t>(p, MAXN);
    t.level_begin = carve<int>(p, MAXL + 2); t.counts = carve<int>(p, 4);
    uint4 root = make_uint4(sv->root.hands, sv->root.table, sv->root.captures, sv->root.meta);
    tree_expand_kernel<<<1, 256>>>(root, sv->hand_order, t);
    MS_LAUNCH_CHECK();
    int counts[4] = {0, 0, 0, 0};
    MS_CUDA(cudaMemcpy(counts, t.counts, sizeof(int) * 3, cudaMemcpyDeviceToHost));
    if (counts[2]) { cudaFree(tmp); return fail(MS_ERR_CAPACITY, "game tree exceeds %d nodes / %d levels", MAXN, MAXL); }
    const int N = counts[0], L = counts[1];
    sv->n_nodes = N; sv->n_levels = L;
    sv->level_begin.resize(L + 1);
    MS_CUDA(cudaMemcpy(sv->level_begin.data(), t.level_begin, sizeof(int) * (L + 1), cudaMemcpyDeviceToHost));
    sv->h_state.resize(N); sv->h_parent.resize(N); sv->h_child_begin.resize(N); sv->h_nchild.resize(N);
    std::vector<unsigned long long> key(N); std::vector<uint16_t> legal(N); std::vector<int8_t> rx2(N);
    MS_CUDA(cudaMemcpy(sv->h_state.data(), t.state, 16 * (size_t)N, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(sv->h_parent.data(), t.parent, 4 * (size_t)N, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(sv->h_child_begin.data(), t.child_begin, 4 * (size_t)N, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(sv->h_nchild.data(), t.nchild, (size_t)N, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(key.data(), t.key, 8 * (size_t)N, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(legal.data(), t.legal, 2 * (size_t)N, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(rx2.data(), t.rx2, (size_t)N, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaFree(tmp));

    // ---- 2. index the infosets on the host (pure bookkeeping: no game rules here).
    // Slots are numbered in breadth-first first-occurrence order: deterministic, hence identical on
    // every GPU -> slot-aligned arrays across ranks.
    sv->h_level.assign(N, 0);
    for (int l = 0; l < L; l++) for (int v = sv->level_begin[l]; v < sv->level_begin[l + 1]; v++) sv->h_level[v] = (uint8_t)l;
    sv->h_slot.assign(N, -1);
    std::unordered_map<unsigned long long, int> slot_of;
    std::vector<int> slot_level;
    std::vector<std::vector<int>> chains;
    sv->h_slot_key.clear(); sv->h_slot_nlegal.clear(); sv->h_slot_legal.clear();
    for (int v = 0; v < N; v++) {
        if (sv->h_nchild[v] == 0) continue;
        auto it = slot_of.find(key[v]);
        int s;
        if (it == slot_of.end()) {
            s = (int)chains.size();
            slot_of.emplace(key[v], s);
            chains.emplace_back();
            slot_level.push_back(sv->h_level[v]);
            sv->h_slot_key.push_back(key[v]);
            sv->h_slot_nlegal.push_back(sv->h_nchild[v]);
            for (int i = 0; i < 4; i++)
                sv->h_slot_legal.push_back(i < sv->h_nchild[v] ? (uint8_t)((legal[v] >> (4 * i)) & 0xF) : (uint8_t)0xFF);
        } else {
            s = it->second;
            if (slot_level[s] != sv->h_level[v])
                return fail(MS_ERR_ARG, "unsupported root: an infoset recurs at two depths (pass moves inside the tree)");
            if (sv->h_slot_nlegal[s] != sv->h_nchild[v]) return fail(MS_ERR_ARG, "inconsistent legal count inside an infoset");
        }
        sv->h_slot[v] = s;
        chains[s].push_back(v);
    }
    const int S = (int)chains.size();
    if (S > MAXS) return fail(MS_ERR_CAPACITY, "%d infosets exceed the table capacity %d", S, MAXS);
    sv->n_slots = S;
    sv->slot_level_begin.assign(L + 1, S);
    {
        int s = 0;
        for (int l = 0; l <= L; l++) {
            while (s < S && slot_level[s] < l) s++;
            sv->slot_level_begin[l] = s;
        }
    }
    std::vector<uint16_t> chain_begin(S + 1), chain_nodes;
    for (int s = 0; s < S; s++) {
        chain_begin[s] = (uint16_t)chain_nodes.size();
        for (int v : chains[s]) chain_nodes.push_back((uint16_t)v);
    }
    chain_begin[S] = (uint16_t)chain_nodes.size();
    sv->n_dec = (int)chain_nodes.size();
    int hcap = 1024;
    while (hcap < 2 * S + 2) hcap *= 2;
    sv->hcap = hcap;
    std::vector<unsigned long long> hk(hcap, 0xFFFFFFFFFFFFFFFFull);
    std::vector<int16_t> hs(hcap, -1);
    for (int s = 0; s < S; s++) {
        uint32_t h = (uint32_t)((sv->h_slot_key[s] * 0x9E3779B97F4A7C15ull) >> 40) & (uint32_t)(hcap - 1);
        while (hk[h] != 0xFFFFFFFFFFFFFFFFull) h = (h + 1) & (uint32_t)(hcap - 1);
        hk[h] = sv->h_slot_key[s]; hs[h] = (int16_t)s;
    }
    const int root_cur = (int)((sv->root.meta >> 17) & 1u);
    // frames needed by the sampled traversals: decision levels of one player
    int dl[2] = {0, 0};
    for (int l = 0; l < L; l++) {
        bool any = false;
        for (int v = sv->level_begin[l]; v < sv->level_begin[l + 1]; v++) any |= sv->h_nchild[v] > 0;
        if (any) dl[(root_cur + l) & 1]++;
    }
    sv->nframes = std::max(1, std::max(dl[0], dl[1]));
    // The tree-walking kernel pushes no frame at a traverser node whose single move leads to the end of the game
    // through forced moves only (its endgame shortcut): on a fresh deal that is every traverser's last card, so
    // three frames suffice instead of four.  Longest chain of frame-pushing nodes of one player, by DP over the tree.
    {
        std::vector<int> need0(N, 0), need1(N, 0);
        for (int v = N - 1; v >= 0; v--) {
            const int nc = sv->h_nchild[v];
            if (nc == 0) continue;
            const int cb = sv->h_child_begin[v];
            int m0 = 0, m1 = 0;
            for (int c = cb; c < cb + nc; c++) { m0 = std::max(m0, need0[c]); m1 = std::max(m1, need1[c]); }
            bool forced = false;
            if (nc == 1) {
                if (sv->h_nchild[cb] == 0) forced = true;
                else if (sv->h_nchild[cb] == 1 && sv->h_nchild[sv->h_child_begin[cb]] == 0) forced = true;
            }
            const int p = (int)((sv->h_state[v].meta >> 17) & 1u);
            const int push = (nc == 1 && forced) ? 0 : 1;
            need0[v] = m0 + (p == 0 ? push : 0);
            need1[v] = m1 + (p == 1 ? push : 0);
        }
        sv->nframes_tree = std::max(1, std::max(need0[0], need1[0]));
        // external sampling pushes a frame at every traverser node with more than one legal action
        std::fill(need0.begin(), need0.end(), 0); std::fill(need1.begin(), need1.end(), 0);
        for (int v = N - 1; v >= 0; v--) {
            const int nc = sv->h_nchild[v];
            if (nc == 0) continue;
            const int cb = sv->h_child_begin[v];
            int m0 = 0, m1 = 0;
            for (int c = cb; c < cb + nc; c++) { m0 = std::max(m0, need0[c]); m1 = std::max(m1, need1[c]); }
            const int p = (int)((sv->h_state[v].meta >> 17) & 1u);
            need0[v] = m0 + ((p == 0 && nc > 1) ? 1 : 0);
            need1[v] = m1 + ((p == 1 && nc > 1) ? 1 : 0);
        }
        sv->nframes_es = std::max(1, std::max(need0[0], need1[0]));
    }

    // Does the tree have the shape of a fresh deal?  9 levels; at ply d < 8 every node has 4 - d/2 children and player d & 1
    // moves; ply 8 is all leaves; the infosets of a ply are a contiguous slot range (they are: breadth-first numbering).
    {
        bool ok = (L == STATIC_PLIES + 1) && root_cur == 0 && S < 2048 && N <= 4096;
        for (int l = 0; ok && l < L; l++)
            for (int v = sv->level_begin[l]; ok && v < sv->level_begin[l + 1]; v++) {
                const int want = l < STATIC_PLIES ? 4 - l / 2 : 0;
                ok = sv->h_nchild[v] == want && (want == 0 || (int)((sv->h_state[v].meta >> 17) & 1u) == (l & 1));
            }
        sv->static_shape = ok;
        if (ok) {
            sv->sdm = static_dims_from(sv->level_begin.data(), sv->slot_level_begin.data());
            if (mccfr_static_smem(S, sv->sdm) + 16 > 227 * 1024) sv->static_shape = false;
        }
    }

    // ---- 3. upload
    size_t total = 0;
    auto sz = [&](size_t n) { size_t b = (n + 255) & ~(size_t)255; total += b; return b; };
    sz(4 * (L + 1)); sz(2 * N); sz(N); sz(2 * N); sz(N); sz(2 * (S + 1)); sz(2 * sv->n_dec); sz(4 * (L + 1)); sz(S); sz(S);
    sz(8 * hcap); sz(2 * hcap); sz(32 * S); sz(32 * S); sz(8 * (6 * S)); sz(8 * (size_t)MS_MAX_PEERS * 6 * S); sz(8 * (size_t)MS_MAX_PEERS * 6 * S); sz(8 * MS_MAX_PEERS); sz(16); sz(S); sz(8 * 4); sz(16);
    MS_CUDA(cudaMalloc(&sv->d_block, total + 4096));
    MS_CUDA(cudaMemset(sv->d_block, 0, total + 4096));
    p = sv->d_block;
    std::vector<uint16_t> cb16(N); std::vector<int16_t> ns16(N); std::vector<uint8_t> splayer(S);
    for (int v = 0; v < N; v++) { cb16[v] = (uint16_t)sv->h_child_begin[v]; ns16[v] = (int16_t)sv->h_slot[v]; }
    for (int s = 0; s < S; s++) splayer[s] = (uint8_t)((sv->h_slot_key[s] >> 52) & 1ull);
    SolverDev& d = sv->dev;
    d.n_nodes = N; d.n_levels = L; d.n_slots = S; d.root_cur = root_cur; d.root = root; d.hand_order = sv->hand_order; d.hcap = hcap;
#define UP(field, T, vec, n)                                                                    \
    do { T* q = carve<T>(p, (n)); MS_CUDA(cudaMemcpy(q, (vec).data(), sizeof(T) * (n), cudaMemcpyHostToDevice)); d.field = q; } while (0)
    UP(level_begin, int, sv->level_begin, (size_t)L + 1);
    UP(child_begin, uint16_t, cb16, (size_t)N);
    UP(nchild, uint8_t, sv->h_nchild, (size_t)N);
    UP(node_slot, int16_t, ns16, (size_t)N);
    UP(rx2, int8_t, rx2, (size_t)N);
    UP(chain_begin, uint16_t, chain_begin, (size_t)S + 1);
    UP(chain_nodes, uint16_t, chain_nodes, (size_t)sv->n_dec);
    UP(slot_level_begin, int, sv->slot_level_begin, (size_t)L + 1);
    UP(slot_nlegal, uint8_t, sv->h_slot_nlegal, (size_t)S);
    UP(slot_player, uint8_t, splayer, (size_t)S);
    UP(hkeys, unsigned long long, hk, (size_t)hcap);
    UP(hslots, int16_t, hs, (size_t)hcap);
#undef UP
    d.regret = carve<double>(p, 4 * (size_t)S);
    d.strategy = carve<double>(p, 4 * (size_t)S);
    d.delta = carve<double>(p, 6 * (size_t)S);
    sv->inbox[0] = carve<double>(p, (size_t)MS_MAX_PEERS * 6 * S);
    sv->inbox[1] = carve<double>(p, (size_t)MS_MAX_PEERS * 6 * S);
    sv->flags = carve<unsigned long long>(p, MS_MAX_PEERS);
    sv->peer_err = carve<unsigned int>(p, 4);
    d.touched = carve<uint8_t>(p, (size_t)S);
    d.counters = carve<unsigned long long>(p, 4);
    sv->d_value = carve<double>(p, 2);
    return MS_OK;
}

int check_dev(const ms_solver* s) {
    if (!s) return fail(MS_ERR_ARG, "null solver");
    int dev = -1;
    MS_CUDA(cudaGetDevice(&dev));
    if (dev != s->device) return fail(MS_ERR_STATE, "solver lives on device %d but the current device is %d", s->device, dev);
    return MS_OK;
}

}  // namespace

extern "C" {

int ms_solver_create(const ms_state* h_root, uint32_t hand_order, ms_solver** out) {
    if (!h_root || !out) return fail(MS_ERR_ARG, "ms_solver_create: bad argument");
    ms_solver* sv = new ms_solver();
    sv->root = *h_root; sv->hand_order = hand_order;
    cudaError_t e = cudaGetDevice(&sv->device);
    if (e != cudaSuccess) { delete sv; return fail(MS_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e)); }
    int rc = solver_build(sv);
    if (rc) { if (sv->d_block) cudaFree(sv->d_block); delete sv; return rc; }
    *out = sv;
    return MS_OK;
}

void ms_solver_destroy(ms_solver* s) {
    if (!s) return;
    if (s->attached)
        for (int r = 0; r < s->world; r++)
            if (r != s->rank && s->peer_base[r]) cudaIpcCloseMemHandle(s->peer_base[r]);
    if (s->d_block) cudaFree(s->d_block);
    delete s;
}

int ms_solver_reset(ms_solver* s, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    // attached peers read this solver's delta buffers at their own pace: clearing them here would race with those reads
    if (s->attached) return fail(MS_ERR_STATE, "ms_solver_reset: the solver is attached to peers (create a new solver instead)");
    const size_t S = s->n_slots;
    cudaStream_t st = (cudaStream_t)stream;
    MS_CUDA(cudaMemsetAsync(s->dev.regret, 0, 32 * S, st));
    MS_CUDA(cudaMemsetAsync(s->dev.strategy, 0, 32 * S, st));
    MS_CUDA(cudaMemsetAsync(s->dev.delta, 0, 48 * S, st));
    MS_CUDA(cudaMemsetAsync(s->dev.touched, 0, S, st));
    MS_CUDA(cudaMemsetAsync(s->dev.counters, 0, 32, st));
    return MS_OK;
}

int ms_solver_counts(const ms_solver* s, int32_t* n_nodes, int32_t* n_slots, int32_t* n_levels) {
    if (!s) return fail(MS_ERR_ARG, "null solver");
    if (n_nodes) *n_nodes = s->n_nodes;
    if (n_slots) *n_slots = s->n_slots;
    if (n_levels) *n_levels = s->n_levels;
    return MS_OK;
}

int ms_solver_export_tree(const ms_solver* s, ms_state* h_states, int32_t* h_parent, uint8_t* h_level, int32_t* h_slot,
                          int32_t* h_child_begin, uint8_t* h_nchild) {
    if (!s) return fail(MS_ERR_ARG, "null solver");
    const size_t N = s->n_nodes;
    if (h_states) memcpy(h_states, s->h_state.data(), 16 * N);
    if (h_parent) memcpy(h_parent, s->h_parent.data(), 4 * N);
    if (h_level) memcpy(h_level, s->h_level.data(), N);
    if (h_slot) memcpy(h_slot, s->h_slot.data(), 4 * N);
    if (h_child_begin) memcpy(h_child_begin, s->h_child_begin.data(), 4 * N);
    if (h_nchild) memcpy(h_nchild, s->h_nchild.data(), N);
    return MS_OK;
}

int ms_solver_export_table(const ms_solver* s, uint64_t* h_keys, uint8_t* h_nlegal, uint8_t* h_legal, double* h_regret,
                           double* h_strategy, uint8_t* h_touched, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    const size_t S = s->n_slots;
    cudaStream_t st = (cudaStream_t)stream;
    if (h_keys) memcpy(h_keys, s->h_slot_key.data(), 8 * S);
    if (h_nlegal) memcpy(h_nlegal, s->h_slot_nlegal.data(), S);
    if (h_legal) memcpy(h_legal, s->h_slot_legal.data(), 4 * S);
    if (h_regret) MS_CUDA(cudaMemcpyAsync(h_regret, s->dev.regret, 32 * S, cudaMemcpyDeviceToHost, st));
    if (h_strategy) MS_CUDA(cudaMemcpyAsync(h_strategy, s->dev.strategy, 32 * S, cudaMemcpyDeviceToHost, st));
    if (h_touched) MS_CUDA(cudaMemcpyAsync(h_touched, s->dev.touched, S, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    return MS_OK;
}

int ms_solver_import_table(ms_solver* s, const double* h_regret, const double* h_strategy, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    const size_t S = s->n_slots;
    cudaStream_t st = (cudaStream_t)stream;
    if (h_regret) MS_CUDA(cudaMemcpyAsync(s->dev.regret, h_regret, 32 * S, cudaMemcpyHostToDevice, st));
    if (h_strategy) MS_CUDA(cudaMemcpyAsync(s->dev.strategy, h_strategy, 32 * S, cudaMemcpyHostToDevice, st));
    MS_CUDA(cudaStreamSynchronize(st));
    return MS_OK;
}

int ms_solver_device_ptrs(ms_solver* s, double** d_regret, double** d_strategy, double** d_delta, size_t* n_table,
                          size_t* n_delta) {
    if (!s) return fail(MS_ERR_ARG, "null solver");
    if (d_regret) *d_regret = s->dev.regret;
    if (d_strategy) *d_strategy = s->dev.strategy;
    if (d_delta) *d_delta = s->dev.delta;
    if (n_table) *n_table = 4 * (size_t)s->n_slots;
    if (n_delta) *n_delta = 6 * (size_t)s->n_slots;
    return MS_OK;
}

static int launch_cfr(ms_solver* s, int iters, int only_player, double r0, double r1, double* d_out, cudaStream_t st) {
    const size_t smem = cfr_smem_bytes(s->n_nodes, s->n_slots, s->n_dec);
    if (smem > 227 * 1024) return fail(MS_ERR_CAPACITY, "CFR working set %zu B exceeds shared memory", smem);
    MS_CUDA(cudaFuncSetAttribute(cfr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfr_kernel<<<1, 512, smem, st>>>(s->dev, s->n_dec, iters, only_player, r0, r1, d_out);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_cfr_iterate(ms_solver* s, int32_t iters, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (iters < 0) return fail(MS_ERR_ARG, "iters < 0");
    if (iters == 0) return MS_OK;
    return launch_cfr(s, iters, -1, 1.0, 1.0, nullptr, (cudaStream_t)stream);
}

int ms_cfr_iterate_many(ms_solver* const* solvers, int32_t n_solvers, int32_t iters, void* stream) {
    if (!solvers || n_solvers < 0 || iters < 0) return fail(MS_ERR_ARG, "ms_cfr_iterate_many: bad argument");
    if (n_solvers == 0 || iters == 0) return MS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<CfrJob> jobs((size_t)n_solvers);
    size_t smem = 0;
    for (int i = 0; i < n_solvers; i++) {
        int rc = check_dev(solvers[i]); if (rc) return rc;
        jobs[i].d = solvers[i]->dev; jobs[i].n_dec = solvers[i]->n_dec;
        smem = std::max(smem, cfr_smem_bytes(solvers[i]->n_nodes, solvers[i]->n_slots, solvers[i]->n_dec));
    }
    if (smem > 227 * 1024) return fail(MS_ERR_CAPACITY, "CFR working set %zu B exceeds shared memory", smem);
    CfrJob* d_jobs = nullptr;
    MS_CUDA(cudaMallocAsync((void**)&d_jobs, sizeof(CfrJob) * (size_t)n_solvers, st));
    MS_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(CfrJob) * (size_t)n_solvers, cudaMemcpyHostToDevice, st));
    MS_CUDA(cudaStreamSynchronize(st));        // `jobs` is pageable host memory: finish the copy before it goes away
    MS_CUDA(cudaFuncSetAttribute(cfr_many_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfr_many_kernel<<<n_solvers, 512, smem, st>>>(d_jobs, iters);
    MS_LAUNCH_CHECK();
    MS_CUDA(cudaFreeAsync(d_jobs, st));
    return MS_OK;
}

int ms_cfr_traverse(ms_solver* s, int32_t player, double reach_p0, double reach_p1, double* h_value, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (player < 0 || player > 1) return fail(MS_ERR_ARG, "player must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_cfr(s, 1, player, reach_p0, reach_p1, s->d_value, st);
    if (rc) return rc;
    if (h_value) {
        MS_CUDA(cudaMemcpyAsync(h_value, s->d_value, 8, cudaMemcpyDeviceToHost, st));
        MS_CUDA(cudaStreamSynchronize(st));
    }
    return MS_OK;
}

int ms_mccfr_inplace(ms_solver* s, int64_t iters, uint64_t philox_seed, uint64_t first_iter, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (iters < 0) return fail(MS_ERR_ARG, "iters < 0");
    if (iters == 0) return MS_OK;
    const uint2 key = make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32));
    if (s->n_slots < (int)TREE_TERMINAL && s->n_nodes <= 4096) {       // walk the enumerated tree (same results)
        const size_t tsmem = 64 * (size_t)s->n_slots + (size_t)s->nframes_tree * 26 + 4 * (size_t)s->n_nodes + s->n_slots + 64;
        if (tsmem <= 227 * 1024) {
            MS_CUDA(cudaFuncSetAttribute(mccfr_inplace_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            mccfr_inplace_tree_kernel<<<1, 32, tsmem, (cudaStream_t)stream>>>(s->dev, (long long)iters, key,
                                                                              (unsigned long long)first_iter, s->nframes_tree);
            MS_LAUNCH_CHECK();
            return MS_OK;
        }
    }
    const int S = s->n_slots, nf = s->nframes;
    size_t smem = 64 * (size_t)S + 8 * (size_t)s->hcap + (size_t)nf * 44 + 2 * (size_t)s->hcap + S + 64;
    if (smem > 227 * 1024) return fail(MS_ERR_CAPACITY, "MCCFR in-place working set %zu B exceeds shared memory", smem);
    MS_CUDA(cudaFuncSetAttribute(mccfr_inplace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mccfr_inplace_kernel<<<1, 32, smem, (cudaStream_t)stream>>>(
        s->dev, (long long)iters, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
        (unsigned long long)first_iter, nf);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

// mode 3 of ms_mccfr_batch_mode: the estimator re-stepping the env at every node (mccfr_batch_kernel)
static int launch_mccfr_restep(ms_solver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav, void* stream) {
    int threads = MCCFR_THREADS;          // deals with more infosets than seed 42 leave room for fewer frames
    while (threads > 128 && mccfr_batch_smem(s->n_slots, s->hcap, s->nframes, threads) > 227 * 1024) threads -= 128;
    const size_t smem = mccfr_batch_smem(s->n_slots, s->hcap, s->nframes, threads);
    if (smem > 227 * 1024) return fail(MS_ERR_CAPACITY, "MCCFR batch working set %zu B exceeds shared memory", smem);
    MS_CUDA(cudaFuncSetAttribute(mccfr_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mccfr_batch_kernel<<<grid_for(n_trav, threads, 1), threads, smem, (cudaStream_t)stream>>>(
        s->dev, player, (long long)n_trav, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
        (unsigned long long)first_trav, s->nframes);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

static int launch_mccfr_tree(ms_solver* s, int threads, int ncopy, int32_t player, int64_t n_trav, uint64_t philox_seed,
                             uint64_t first_trav, void* stream) {
    const size_t smem = mccfr_tree_smem(s->n_slots, s->n_nodes, s->nframes_tree, threads, ncopy);
    const uint2 key = make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32));
    const int grid = grid_for(n_trav, threads, 1);
    if (threads == TREE_THREADS) {
        MS_CUDA(cudaFuncSetAttribute(mccfr_tree_kernel<TREE_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mccfr_tree_kernel<TREE_THREADS><<<grid, TREE_THREADS, smem, (cudaStream_t)stream>>>(
            s->dev, player, (long long)n_trav, key, (unsigned long long)first_trav, s->nframes_tree, ncopy);
    } else {
        MS_CUDA(cudaFuncSetAttribute(mccfr_tree_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mccfr_tree_kernel<512><<<grid, 512, smem, (cudaStream_t)stream>>>(
            s->dev, player, (long long)n_trav, key, (unsigned long long)first_trav, s->nframes_tree, ncopy);
    }
    MS_LAUNCH_CHECK();
    return MS_OK;
}

static int launch_mccfr_static(ms_solver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav, void* stream) {
    const size_t smem = mccfr_static_smem(s->n_slots, s->sdm) + 16;
    MS_CUDA(cudaFuncSetAttribute(mccfr_static_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mccfr_static_kernel<<<grid_for(n_trav, STATIC_THREADS, 1), STATIC_THREADS, smem, (cudaStream_t)stream>>>(
        s->dev, player, (long long)n_trav, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
        (unsigned long long)first_trav, s->sdm);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

// the generic tree-walking kernel (any root whose tree fits): mode 4, and mode 0 on trees without the fresh-deal shape
static int launch_mccfr_generic(ms_solver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav, void* stream) {
    if (s->n_slots >= (int)TREE_TERMINAL || s->n_nodes > 4096) return launch_mccfr_restep(s, player, n_trav, philox_seed, first_trav, stream);
    for (int threads : {TREE_THREADS, 512})
        for (int ncopy : {4, 2, 1})
            if (mccfr_tree_smem(s->n_slots, s->n_nodes, s->nframes_tree, threads, ncopy) <= 227 * 1024)
                return launch_mccfr_tree(s, threads, ncopy, player, n_trav, philox_seed, first_trav, stream);
    return launch_mccfr_restep(s, player, n_trav, philox_seed, first_trav, stream);
}

int ms_mccfr_batch(ms_solver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (player < 0 || player > 2 || n_trav < 0) return fail(MS_ERR_ARG, "ms_mccfr_batch: bad argument");
    if (n_trav == 0) return MS_OK;
    if (s->static_shape) return launch_mccfr_static(s, player, n_trav, philox_seed, first_trav, stream);
    return launch_mccfr_generic(s, player, n_trav, philox_seed, first_trav, stream);
}

int ms_mccfr_batch_mode(ms_solver* s, int32_t mode, int32_t player, int64_t n_trav, uint64_t philox_seed,
                        uint64_t first_trav, void* stream) {
    if (mode == 0) return ms_mccfr_batch(s, player, n_trav, philox_seed, first_trav, stream);
    int rc = check_dev(s); if (rc) return rc;
    if (mode < 0 || mode > 4 || player < 0 || player > 2 || n_trav < 0) return fail(MS_ERR_ARG, "ms_mccfr_batch_mode: bad argument");
    if (n_trav == 0) return MS_OK;
    if (mode == 3) return launch_mccfr_restep(s, player, n_trav, philox_seed, first_trav, stream);
    if (mode == 4) return launch_mccfr_generic(s, player, n_trav, philox_seed, first_trav, stream);
    const uint2 key = make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32));
    if (mode == 1 && s->n_slots < (int)TREE_TERMINAL && s->n_nodes <= 4096) {
        for (int ncopy : {4, 2, 1}) {
            const size_t smem = es_tree_smem(s->n_slots, s->n_nodes, s->nframes_es, TREE_THREADS, ncopy);
            if (smem > 227 * 1024) continue;
            MS_CUDA(cudaFuncSetAttribute(mccfr_es_tree_kernel<TREE_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mccfr_es_tree_kernel<TREE_THREADS><<<grid_for(n_trav, TREE_THREADS, 1), TREE_THREADS, smem, (cudaStream_t)stream>>>(
                s->dev, player, (long long)n_trav, key, (unsigned long long)first_trav, s->nframes_es, ncopy);
            MS_LAUNCH_CHECK();
            return MS_OK;
        }
    }
    if (mode == 1) {
        const size_t smem = es_smem_bytes(s->n_slots, s->hcap, s->nframes, ES_THREADS);
        if (smem > 227 * 1024) return fail(MS_ERR_CAPACITY, "ES working set %zu B exceeds shared memory", smem);
        MS_CUDA(cudaFuncSetAttribute(mccfr_es_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mccfr_es_kernel<<<grid_for(n_trav, ES_THREADS, 1), ES_THREADS, smem, (cudaStream_t)stream>>>(
            s->dev, player, (long long)n_trav, key, (unsigned long long)first_trav, s->nframes);
    } else {
        const size_t smem = 72 * (size_t)s->n_slots + 10 * (size_t)s->hcap + s->n_slots + 64;
        if (smem > 227 * 1024) return fail(MS_ERR_CAPACITY, "OS working set %zu B exceeds shared memory", smem);
        MS_CUDA(cudaFuncSetAttribute(mccfr_os_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 8) per_sm = 8;
        if (per_sm < 1) per_sm = 1;
        mccfr_os_kernel<<<grid_for(n_trav, 256, per_sm), 256, smem, (cudaStream_t)stream>>>(
            s->dev, player, (long long)n_trav, key, (unsigned long long)first_trav);
    }
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_mccfr_apply(ms_solver* s, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    mccfr_apply_kernel<<<(s->n_slots + 255) / 256, 256, 0, (cudaStream_t)stream>>>(s->dev);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_solver_ipc_export(ms_solver* s, void* handle64, uint64_t offsets[3]) {
    int rc = check_dev(s); if (rc) return rc;
    if (!handle64 || !offsets) return fail(MS_ERR_ARG, "ms_solver_ipc_export: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    MS_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, s->d_block));
    offsets[0] = (uint64_t)((char*)s->inbox[0] - s->d_block);
    offsets[1] = (uint64_t)((char*)s->inbox[1] - s->d_block);
    offsets[2] = (uint64_t)((char*)s->flags - s->d_block);
    return MS_OK;
}

int ms_solver_ipc_attach(ms_solver* s, int32_t rank, int32_t world, const void* handles, const uint64_t* offsets) {
    int rc = check_dev(s); if (rc) return rc;
    if (world < 1 || world > MS_MAX_PEERS || rank < 0 || rank >= world || !handles || !offsets)
        return fail(MS_ERR_ARG, "ms_solver_ipc_attach: bad argument (at most %d ranks)", MS_MAX_PEERS);
    if (s->attached) return fail(MS_ERR_STATE, "solver is already attached to its peers");
    for (int r = 0; r < world; r++) {
        for (int k = 0; k < 3; k++) s->peer_off[r][k] = offsets[3 * r + k];
        if (r == rank) { s->peer_base[r] = s->d_block; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * r, 64);
        MS_CUDA(cudaIpcOpenMemHandle(&s->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    s->rank = rank; s->world = world; s->attached = true; s->parity = 0; s->epoch = 0;
    return MS_OK;
}

static PeerView peer_view(ms_solver* s) {
    PeerView pv{};
    for (int r = 0; r < s->world; r++) {
        pv.inbox[r] = (double*)((char*)s->peer_base[r] + s->peer_off[r][s->parity]);
        pv.flags[r] = (unsigned long long*)((char*)s->peer_base[r] + s->peer_off[r][2]);
    }
    pv.my_flags = s->flags;
    pv.rank = s->rank; pv.world = s->world;
    return pv;
}

static void peer_advance(ms_solver* s) { s->parity ^= 1; }     // the next iteration's pushes go to the other inbox

int ms_mccfr_apply_peers(ms_solver* s, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (!s->attached) return fail(MS_ERR_STATE, "ms_mccfr_apply_peers: call ms_solver_ipc_attach first");
    const PeerView pv = peer_view(s);
    s->epoch += 1;
    const size_t smem = 16 + 48 * (size_t)s->n_slots;
    MS_CUDA(cudaFuncSetAttribute(mccfr_apply_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mccfr_apply_peers_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(s->dev, pv, s->epoch, s->peer_err);
    MS_LAUNCH_CHECK();
    peer_advance(s);
    return MS_OK;
}

int ms_mccfr_batch_peers(ms_solver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (!s->attached) return fail(MS_ERR_STATE, "ms_mccfr_batch_peers: call ms_solver_ipc_attach first");
    if (player < 0 || player > 2 || n_trav < 0) return fail(MS_ERR_ARG, "ms_mccfr_batch_peers: bad argument");
    if (!s->static_shape) {               // other roots: the generic traversal kernel, then the exchange kernel
        if (n_trav > 0) { rc = launch_mccfr_generic(s, player, n_trav, philox_seed, first_trav, stream); if (rc) return rc; }
        return ms_mccfr_apply_peers(s, stream);
    }
    const size_t smem = mccfr_static_smem(s->n_slots, s->sdm) + 16;
    const PeerView pv = peer_view(s);
    s->epoch += 1;
    MS_CUDA(cudaFuncSetAttribute(mccfr_static_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mccfr_static_peers_kernel<<<grid_for(n_trav > 0 ? n_trav : 1, STATIC_THREADS, 1), STATIC_THREADS, smem, (cudaStream_t)stream>>>(
        s->dev, player, (long long)n_trav, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
        (unsigned long long)first_trav, s->sdm, pv, s->epoch, s->peer_err, s->peer_err + 1);
    MS_LAUNCH_CHECK();
    peer_advance(s);
    return MS_OK;
}

int ms_solver_peer_error(ms_solver* s, uint32_t* h_err, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (!h_err) return fail(MS_ERR_ARG, "ms_solver_peer_error: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MS_CUDA(cudaMemcpyAsync(h_err, s->peer_err, 4, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    if (*h_err) return fail(MS_ERR_STATE, "peer exchange: rank %u did not arrive within %llu ms; the table was left unchanged from "
                            "that iteration on", *h_err - 1u, (unsigned long long)(MS_PEER_TIMEOUT_NS / 1000000ull));
    return MS_OK;
}

int ms_mccfr_inplace_many(ms_solver* s, int32_t n_runs, int64_t iters, uint64_t philox_seed0, uint64_t first_iter,
                          double* d_regret, double* d_strategy, uint8_t* d_touched, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (n_runs < 0 || iters < 0 || !d_regret || !d_strategy || !d_touched) return fail(MS_ERR_ARG, "ms_mccfr_inplace_many: bad argument");
    if (n_runs == 0 || iters == 0) return MS_OK;
    if (s->n_slots >= (int)TREE_TERMINAL || s->n_nodes > 4096) return fail(MS_ERR_CAPACITY, "ms_mccfr_inplace_many: tree too large");
    const size_t tree_b = (4 * (size_t)s->n_nodes + 15) & ~(size_t)15, per_warp = inplace_many_warp_bytes(s->n_slots, s->nframes_tree);
    int warps = (int)((227 * 1024 - tree_b) / per_warp);
    if (warps > 8) warps = 8;
    if (warps < 1) return fail(MS_ERR_CAPACITY, "ms_mccfr_inplace_many: one run's table exceeds shared memory");
    if (warps > n_runs) warps = n_runs;
    const size_t smem = tree_b + (size_t)warps * per_warp;
    ManyRuns m{d_regret, d_strategy, d_touched, n_runs};
    MS_CUDA(cudaFuncSetAttribute(mccfr_inplace_many_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mccfr_inplace_many_kernel<<<(n_runs + warps - 1) / warps, 256, smem, (cudaStream_t)stream>>>(
        s->dev, m, (long long)iters, (unsigned long long)philox_seed0, (unsigned long long)first_iter, s->nframes_tree, warps);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_solver_counters(ms_solver* s, uint64_t h_out[3], int reset, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (h_out) {
        MS_CUDA(cudaMemcpyAsync(h_out, s->dev.counters, 24, cudaMemcpyDeviceToHost, st));
        MS_CUDA(cudaStreamSynchronize(st));
    }
    if (reset) MS_CUDA(cudaMemsetAsync(s->dev.counters, 0, 32, st));
    return MS_OK;
}

int ms_solver_policy(ms_solver* s, int32_t policy_kind, double* d_policy, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (policy_kind < 0 || policy_kind > 2 || !d_policy) return fail(MS_ERR_ARG, "ms_solver_policy: bad argument");
    policy_kernel<<<(s->n_slots + 255) / 256, 256, 0, (cudaStream_t)stream>>>(s->dev, policy_kind, d_policy);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_eval_policies(ms_solver* s, const double* d_policy_seat0, const double* d_policy_seat1, int64_t n_games,
                     uint64_t philox_seed, uint64_t first_game, float* d_reward0, uint8_t* d_scopas, void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (n_games < 0 || !d_policy_seat0 || !d_policy_seat1) return fail(MS_ERR_ARG, "ms_eval_policies: bad argument");
    if (n_games == 0) return MS_OK;
    eval_kernel<<<grid_for(n_games, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        s->dev, d_policy_seat0, d_policy_seat1, (long long)n_games,
        make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)), (unsigned long long)first_game, d_reward0,
        (uchar2*)d_scopas);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_best_response(ms_solver* s, int32_t policy_kind, double h_br_values[2], void* stream) {
    int rc = check_dev(s); if (rc) return rc;
    if (policy_kind < 0 || policy_kind > 2 || !h_br_values) return fail(MS_ERR_ARG, "ms_best_response: bad argument");
    const size_t smem = cfr_smem_bytes(s->n_nodes, s->n_slots, s->n_dec);
    if (smem > 227 * 1024) return fail(MS_ERR_CAPACITY, "BR working set %zu B exceeds shared memory", smem);
    cudaStream_t st = (cudaStream_t)stream;
    MS_CUDA(cudaFuncSetAttribute(best_response_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    best_response_kernel<<<1, 512, smem, st>>>(s->dev, s->n_dec, policy_kind, s->d_value);
    MS_LAUNCH_CHECK();
    MS_CUDA(cudaMemcpyAsync(h_br_values, s->d_value, 16, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    return MS_OK;
}

}  // extern "C"
#endif  // MS_HOST_RULES_ONLY
