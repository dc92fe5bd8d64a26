// scopa_b200/csrc/ms_static_walk.cuh -- the headline MCCFR kernel: the reference's sampled-CFR estimator (mc_cfr.py:37-86)
// with frozen-sigma batch semantics, specialised for the shape EVERY fresh Miniscopa deal has.
//
// A fresh 4+4-card deal is 8 plies, player 0 first, players alternating, and the mover at ply d always holds
// 4 - d/2 cards: the number of legal actions, the player to move and the depth of the end are functions of the ply
// alone (solver_build checks this on the enumerated tree: `static_shape`).  The recursion of the reference's estimator
// -- at a traverser node: the sampled action, then every action, each with a fresh sampled continuation; at an opponent
// node: one sampled action -- therefore has a data-INDEPENDENT shape: 411 (player 0) / 292 (player 1) node visits per
// traversal, always in the same order.  Only WHICH node sits at each position depends on the random draws.
//
// mccfr_tree_kernel (the generic form, any root) runs that recursion as an explicit DFS with 26-byte frames in shared
// memory and a run-time state machine: 52 k thread-instructions per traversal pair, 77 % of the issue slots
// (profiles/README.md section 1).  Here the recursion is written out as nested loops, one template instantiation per
// ply:
//   * the DFS frames are registers (per traverser level: the path weight, the node record, the packed child values);
//   * reach_opp / sample_own is ONE running product w (opponent sigma and traverser 1/sigma factors from two shared
//     tables), so an update needs no division;
//   * an opponent visit is one 16-byte node record (first child | slot, and the strategy's cdf as three 31-bit integer
//     thresholds), an integer compare per action, and one fp64 multiply;
//   * the last two plies are forced: a ply-6 node's record carries both infoset slots and the leaf's reward;
//   * random numbers: Philox4x32-10 keyed by (traversal id, draw index / 4), four 31-bit uniforms per block, consumed
//     in visiting order ("sequential stream"; oracle: ora_mccfr_batch_seq);
//   * regret deltas go to lane-private accumulator columns in shared memory, nl - 1 per infoset (see StaticShared);
//   * visit counters are arithmetic (the shape is static), not per-visit increments.
// Semantics are those of mccfr_tree_kernel<false> (same estimator, same frozen-sigma batch, same delta layout); the
// random stream differs, so the two agree statistically, and each agrees with its own oracle function to 1e-9.
#pragma once
#include <cstdint>

#include "ms_state.cuh"
#include "ms_tree_walk.cuh"

namespace ms {

constexpr uint32_t MS_TAG_MCCF_SEQ = MS_TAG_MCCF + 64u;
constexpr int STATIC_THREADS = 1024;
constexpr int STATIC_PLIES = 8;          // plies of a fresh deal
constexpr int STATIC_DECIDED = 6;        // plies with more than one legal action

// visits / updates / tree edges of ONE traversal by traverser `tp`, counted as the generic kernels count them (and as the
// reference's recursion visits them): a traverser node with nl actions makes nl + 1 recursive calls; the forced last
// plies are visited by both calls of the forced traverser node but their edges are played once.
__host__ __device__ inline void static_shape_counts(int tp, unsigned long long& visits, unsigned long long& updates,
                                                    unsigned long long& edges) {
    unsigned long long V = 1, U = 0, E = 0;           // a terminal node
    for (int ply = STATIC_PLIES - 1; ply >= 0; ply--) {
        const unsigned long long nl = 4 - ply / 2;
        if ((ply & 1) == tp) { V = 1 + (nl + 1) * V; U = 1 + (nl + 1) * U; E = (nl == 1 ? 1 : nl + 1) * (1 + E); }
        else { V = 1 + V; E = 1 + E; }
    }
    visits = V; updates = U; edges = E;
}

// Sizes and offsets of one deal's static layout (host: static_dims_from; passed to the kernel by value).
//   n5 / n6 / n7: first node id of ply 5 / 6 / 7 (nodes are numbered ply by ply); sb[p]: first infoset slot of ply p (slots
//   are numbered ply by ply too: breadth-first first occurrence), sb[6] = S2 = number of infosets with more than one
//   action; accbase[p]: first accumulator row of ply p; n_acc rows in all.
struct StaticDims { int n5, n6, n7, S2, n_acc; int sb[7]; int accbase[6]; };

__host__ __device__ inline StaticDims static_dims_from(const int* level_begin, const int* slot_level_begin) {
    StaticDims dm{};
    dm.n5 = level_begin[5]; dm.n6 = level_begin[6]; dm.n7 = level_begin[7];
    for (int p = 0; p <= 6; p++) dm.sb[p] = slot_level_begin[p];
    dm.S2 = dm.sb[6];
    int a = 0;
    for (int p = 0; p < 6; p++) { dm.accbase[p] = a; a += (3 - p / 2) * (dm.sb[p + 1] - dm.sb[p]); }   // nl - 1 rows per infoset
    dm.n_acc = a;
    return dm;
}

// Regret deltas without intra-warp collisions and with nl - 1 accumulators per infoset.
//   * The regret delta of a visit is w * (cfv_a - v) with v = sum_j sigma_j cfv_j, and sigma is FROZEN for the batch.
//     With e_i = cfv_i - cfv_last (small exact half-integers) the batch total is  delta_a = D_a - sum_j sigma_j D_j,
//     D_i = sum over visits of w * e_i  (D_last = 0): a visit adds nl - 1 numbers and needs no sigma at all; the
//     sigma-weighted part is applied once per infoset when the CTA flushes (mccfr_static_body).
//   * A shared-memory fp64 atomicAdd is a compare-and-swap loop, and ncu (profiles/README.md, r02b) measured 27 shared
//     wavefronts per executed CAS and 2.7 executions per add when the lanes of a warp -- which walk in lock-step and sit
//     on a handful of infosets -- share accumulator copies: 71 % of the kernel's shared-memory traffic.  Here every
//     accumulator row is [32 lanes] wide and a lane only ever touches its own column: no two lanes of a warp can
//     collide, the access is conflict-free (consecutive doubles), and the CAS only arbitrates between the warps of the
//     CTA (rare).  575 rows x 256 B at most (a deal has at most 1 + 4 + 16 + 48 + 144 + 288 such infosets).
struct StaticShared {
    const uint4* node;        // [n6] plies 0..4: {thr0, thr1, thr2, first child | slot << 12}, thr_i = ceil(cdf_i * 2^31);
                              //      ply 5: {thr0, first child | slot << 12, endgame of child 0, endgame of child 1}
    const double* sig;        // [S2][4] frozen strategies
    const double* rsig;       // [S2][4] 1 / sigma (0 where sigma == 0: the reference's weight is 0 when the sampling prob is 0)
    double* acc;              // [n_acc][32] D accumulators, pre-offset by lane
    const uint16_t* accrow;   // MODE 1 only: first accumulator row of each (local) infoset
    uint32_t* dcnt;           // [S] update counts (strategy delta = count * sigma)
    uint8_t* touched;         // [S]
    // per-thread random stream
    uint4 blk; uint32_t nd; uint32_t t_lo, t_hi, tag; uint2 key;
};

// The ten Philox rounds are inlined at each of the twelve draw sites of the two walks: the key is a kernel parameter, so
// the round-key schedule runs on the uniform datapath and the call / argument marshalling of an out-of-line copy
// disappears (measured: 486 -> 560 G updates/s; the out-of-line form had been chosen for code size, which costs less
// than it saves here: the instruction cache misses of a 60 KB kernel are not what binds it).
#ifndef MS_STATIC_PHILOX_ATTR
#define MS_STATIC_PHILOX_ATTR __forceinline__
#endif
__device__ MS_STATIC_PHILOX_ATTR uint4 static_philox(uint32_t t_lo, uint32_t t_hi, uint32_t blk, uint32_t tag, uint2 key) {
    return philox4x32_10(make_uint4(t_lo, t_hi, blk, tag), key);
}

// The d-th uniform of a traversal is word d & 3 of block d >> 2: the four words of the current block sit in a shift
// register (c.blk.x is the next one), refilled every fourth draw -- no word-select logic per draw.
__device__ __forceinline__ uint32_t static_draw(StaticShared& c) {
    if ((c.nd & 3u) == 0u) c.blk = static_philox(c.t_lo, c.t_hi, c.nd >> 2, c.tag, c.key);
    c.nd++;
    const uint32_t w = c.blk.x;
    c.blk.x = c.blk.y; c.blk.y = c.blk.z; c.blk.z = c.blk.w;
    return w >> 1;                                   // 31-bit uniform: u = (w >> 1) / 2^31
}

// searchsorted(cdf, u, 'right') on the integer thresholds T_i = ceil(cdf_i * 2^31): cdf_i <= u/2^31  <=>  T_i <= u
template <int NL>
__device__ __forceinline__ int static_pick(const uint4& rec, uint32_t u) {
    int ai = (rec.x <= u) ? 1 : 0;
    if (NL > 2) ai += (rec.y <= u) ? 1 : 0;
    if (NL > 3) ai += (rec.z <= u) ? 1 : 0;
    return ai;
}

// TOUCH: record first touches (only while some infoset of the deal has never been visited: the first batches).
// MODE 0: the one-deal solver -- infoset slots are numbered ply by ply, so an infoset's accumulator rows follow from its
//   slot by arithmetic (plane layout: row = accbase[ply] + i * (infosets of the ply) + slot - sb[ply]), and the one-card
//   infosets of the last two plies are counted.  MODE 1: the deal-blocked multi-deal solver (ms_multideal.cu) -- local
//   infoset indices in no particular order, so the first row of an infoset comes from a table (rows i = 0 .. nl - 2 are
//   consecutive); one-card infosets are not stored there, so the forced endgame is just the leaf's reward.
template <int PLY, int TP, bool TOUCH, int MODE = 0>
struct StaticWalk {
    static constexpr int NL = 4 - PLY / 2;
    static constexpr bool MINE = (PLY & 1) == TP;

    // -> 2 * (reward of the traverser) of the sampled line below `node`
    static __device__ __forceinline__ int run(uint32_t node, double w, StaticShared& c, const StaticDims& dm) {
        const uint4 rec = c.node[node];
        const uint32_t slot = rec.w >> 12, cb = rec.w & 0xFFFu;
        if (TOUCH) c.touched[slot] = 1;              // node created on first touch, for both players (mc_cfr.py:52)
        const int ai = static_pick<NL>(rec, static_draw(c));
        if (!MINE) {                                 // opponent: reach *= sigma[a]; tail call (mc_cfr.py:63-65)
            return StaticWalk<PLY + 1, TP, TOUCH, MODE>::run(cb + (uint32_t)ai, __dmul_rn(w, c.sig[4 * slot + ai]), c, dm);
        }
        // traverser: the sampled action first (:58-67), then every action with a fresh sampled continuation (:71-78)
        uint32_t cfvb = 0u;
        int util = 0;
#pragma unroll(PLY >= 4 ? NL + 1 : 1)           // the innermost child loops (3 short iterations) are written out
        for (int j = -1; j < NL; j++) {
            const int a = j < 0 ? ai : j;
            const int r = StaticWalk<PLY + 1, TP, TOUCH, MODE>::run(cb + (uint32_t)a, __dmul_rn(w, c.rsig[4 * slot + a]), c, dm);
            if (j < 0) util = r;
            else cfvb |= ((uint32_t)r & 0xFFu) << (8 * j);
        }
        // regret deltas (:79-84), weight = reach_opp / sample_own = w: D_i += w * (cfv_i - cfv_last), see StaticShared;
        // strategy delta = count * sigma
        const int last = (int)(int8_t)((cfvb >> (8 * (NL - 1))) & 0xFFu);
        double* row = MODE == 0 ? c.acc + 32 * (dm.accbase[PLY] + (int)slot - dm.sb[PLY]) : c.acc + 32 * (int)c.accrow[slot];
        const int plane = MODE == 0 ? 32 * (dm.sb[PLY + 1] - dm.sb[PLY]) : 32;
#pragma unroll
        for (int i = 0; i < NL - 1; i++) {
            const int e2 = (int)(int8_t)((cfvb >> (8 * i)) & 0xFFu) - last;          // 2 * (cfv_i - cfv_last)
            const double val = __dmul_rn(w, 0.5 * (double)e2);
            if (val != 0.0) atomicAdd(row + i * plane, val);
        }
        atomicAdd(&c.dcnt[slot], 1u);
        return util;
    }
};

// plies 6 and 7: one card each, both moves forced, then the end.  The reference's traverser node samples its only
// action and then evaluates it again: both calls walk this same line, so it is played once (and counted twice by the
// arithmetic visit counters).  Regret delta = w * (cfv - v) = 0 exactly; strategy delta = 1 * [1.0].
// e = slot of the ply-6 infoset | slot of the ply-7 infoset << 11 | (2 * reward of player 0 + 16) << 22
template <int TP, bool TOUCH, int MODE>
__device__ __forceinline__ int static_endgame(uint32_t e, StaticShared& c) {
    if (MODE == 0) {
        const uint32_t slot6 = e & 0x7FFu, slot7 = (e >> 11) & 0x7FFu;
        if (TOUCH) { c.touched[slot6] = 1; c.touched[slot7] = 1; }
        atomicAdd(&c.dcnt[TP == 0 ? slot6 : slot7], 1u);
    }
    const int r = (int)((e >> 22) & 0x3Fu) - 16;
    return TP == 0 ? r : -r;
}

// ply 5 (two cards in hand): the node record carries the forced endgames of BOTH children, {thr0, first child | slot << 12,
// endgame of child 0, endgame of child 1}, so the 60 endgames of a traversal cost no further (dependent) load.
template <int TP, bool TOUCH, int MODE>
struct StaticWalk<5, TP, TOUCH, MODE> {
    static __device__ __forceinline__ int run(uint32_t node, double w, StaticShared& c, const StaticDims& dm) {
        const uint4 rec = c.node[node];
        const uint32_t slot = rec.y >> 12;
        if (TOUCH) c.touched[slot] = 1;
        const uint32_t u = static_draw(c);
        const uint32_t e_s = (rec.x <= u) ? rec.w : rec.z;           // the sampled child's endgame
        if (TP == 0) return static_endgame<TP, TOUCH, MODE>(e_s, c);      // opponent node: tail call; the weight is dead below
        const int util = static_endgame<TP, TOUCH, MODE>(e_s, c);         // traverser: sampled action, then both actions
        const int r0 = static_endgame<TP, TOUCH, MODE>(rec.z, c);
        const int r1 = static_endgame<TP, TOUCH, MODE>(rec.w, c);
        const double val = __dmul_rn(w, 0.5 * (double)(r0 - r1));   // D_0 += w * (cfv_0 - cfv_1)
        if (val != 0.0) atomicAdd(MODE == 0 ? c.acc + 32 * (dm.accbase[5] + (int)slot - dm.sb[5]) : c.acc + 32 * (int)c.accrow[slot], val);
        atomicAdd(&c.dcnt[slot], 1u);
        return util;
    }
};

__host__ __device__ inline size_t mccfr_static_smem(int S, const StaticDims& dm) {
    return 16 * (size_t)dm.n6 + sizeof(double) * 8 * (size_t)dm.S2 + sizeof(double) * 32 * (size_t)dm.n_acc +
           4 * (size_t)S + (size_t)S + 64;
}

}  // namespace ms
