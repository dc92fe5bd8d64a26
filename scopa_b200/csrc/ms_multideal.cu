// scopa_b200/csrc/ms_multideal.cu -- the sampled-CFR estimator of the reference over MANY deals with one
// HBM-resident open-addressing infoset table ("multi-deal / chance-sampled
// root ... the regime where the HBM / atomic roofline is actually the bound").  sm_100a only.
//
// The reference solves one fixed deal (seed 42): MCCFRTrainer.iteration (src/algorithms/mc_cfr.py:88-92) calls
// _sample (:37-86) on game.new_initial_state() for each player, and its info_sets dict (:28-35) holds 738 nodes
// -- 53 KB, which ms_solver.cu keeps in shared memory.  Here the root is a chance node over D deals
// (MiniScopaEnv.reset(seed_d), src/envs/mini_scopa_game.py:131-138): traversal t draws its deal from a Philox
// stream, then runs the same _sample recursion for both players.  Infosets are merged by information content
// (player, hand SET, ordered table = the 64-bit key of ms_state.cuh); the reference's string lists the hand in
// deal order, which is presentation.  Regret / strategy arrays are indexed like the reference's (by card id),
// compacted to the cards in hand: column = rank of the card id inside the hand mask.  Legal actions keep the
// reference's order (deal order) for sampling, so with D = 1 this is the batch estimator of ms_solver.cu on
// the same Philox streams (tests/test_gpu_multideal.py checks that, and the oracle follows it for D > 1).
//
// Table: capacity 2^k slots of 128 B, one cache line per infoset, four 32-byte sectors:
//   [key u64 | visit count u32 | pad] [regret f64 x4] [regret delta f64 x4] [strategy sum f64 x4]
// A node visit reads sectors 0-1 of one line (key probe + frozen regrets), an update issues fp64 RED.ADDs into
// sector 2 and a u32 RED.ADD into sector 0; md_apply_kernel folds deltas in after the batch (same frozen-sigma
// batch semantics as ms_mccfr_batch + ms_mccfr_apply).  Keys are claimed with atomicCAS (0 = empty: a key with
// an empty hand is never a decision node), so the table is initialised by one memset.
#include <cstring>

#include "ms_common.cuh"
#include "ms_state.cuh"
#include "ms_tree_walk.cuh"
#include "ms_static_walk.cuh"

#ifndef MS_DYN_SMEM   // the host emulation (tests/emu) supplies its own: one buffer per emulated block
#define MS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace ms {

struct __align__(128) MdSlot {
    unsigned long long key;      // 0 = empty
    unsigned int cnt;            // traverser visits in the running batch: strategy_sum += cnt * sigma
    unsigned int pad0;
    unsigned long long pad1, pad2;
    double regret[4];
    double delta[4];
    double strategy[4];
};
static_assert(sizeof(MdSlot) == 128, "one cache line per infoset");

constexpr int MD_MAX_PEERS = 8;

struct MdDev {
    MdSlot* slots;               // this rank's shard of the table (the whole table on one GPU)
    unsigned long long mask;     // capacity of a shard - 1
    int shift;                   // 64 - log2(capacity of a shard)
    unsigned int max_probe;
    const uint4* roots;          // [n_deals] dealt states
    const uint32_t* hand_order;  // [n_deals]
    unsigned int n_deals;
    unsigned int* dirty;         // one bit per slot: touched by an update of the running batch (L2-resident: 16 MB at 2^27)
    unsigned long long* counters;   // [0] updates [1] visits [2] env steps [3] infosets [4] overflow / invariant flag
    // Table sharded over the GPUs of one box ("shard by hash(key) % G"; ms_md_ipc_attach): infoset `key`
    // lives on rank md_owner(key), in that rank's shard, which every rank has mapped (CUDA IPC over NVLink / NVSwitch).
    // A global slot id is owner << lshift | slot inside the shard.  One GPU: world = 1, peer_slots[0] = slots.
    int world, rank, lshift;
    MdSlot* peer_slots[MD_MAX_PEERS];
    unsigned int* peer_dirty[MD_MAX_PEERS];
};

// Which rank's shard holds `key`: a second multiplicative hash, independent of the probe position inside the shard.
__device__ __forceinline__ int md_owner(const MdDev& t, unsigned long long key) {
    if (t.world <= 1) return 0;
    return (int)__umulhi((uint32_t)((key * 0xD6E8FEB86659FD93ull) >> 32), (uint32_t)t.world);
}
__device__ __forceinline__ MdSlot* md_slot(const MdDev& t, long long g) {
    return t.peer_slots[(int)(g >> t.lshift)] + (g & (long long)t.mask);
}
// Table words another GPU may have written (peers' REDs, the owner's apply step) are read past L1 and, on a sharded
// table, as system-scope volatile loads: a peer's line is never served from a cache on this side of the link.
template <class T>
__device__ __forceinline__ T md_ld(const MdDev& t, const T* p) { return t.world > 1 ? __ldcv(p) : __ldcg(p); }
// Updates: RED.ADDs performed at the owner's L2; system scope when the owner is (or may be) another GPU.
__device__ __forceinline__ void md_add(const MdDev& t, double* p, double v) {
    if (t.world > 1) atomicAdd_system(p, v); else atomicAdd(p, v);
}
__device__ __forceinline__ void md_add(const MdDev& t, unsigned int* p, unsigned int v) {
    if (t.world > 1) atomicAdd_system(p, v); else atomicAdd(p, v);
}
__device__ __forceinline__ void md_mark_dirty(const MdDev& t, long long g) {
    const long long l = g & (long long)t.mask;
    unsigned int* w = t.peer_dirty[(int)(g >> t.lshift)] + (l >> 5);
    if (t.world > 1) atomicOr_system(w, 1u << (l & 31)); else atomicOr(w, 1u << (l & 31));
}

constexpr uint32_t MS_TAG_DEAL = 0x4C414544u;   // "DEAL"
constexpr int MD_THREADS = 768;
constexpr int MD_FRAMES = 4;                    // traverser nodes on a path = cards in a hand
constexpr int MD_SIG_FRAMES = 3;                // the 4th traverser node holds one card: sigma = [1]

__device__ __forceinline__ unsigned long long md_hash(unsigned long long key, int shift) {
    return (key * 0x9E3779B97F4A7C15ull) >> shift;
}

// Find the slot of `key`, claiming an empty one on first touch (mc_cfr.py:32-35 _get_node), and fetch its frozen
// regrets (sectors 0-1 of the line, requested together with the key).  Linear probing, one line per round trip: a
// key always sits in the first slot of its probe sequence that was free, the invariant lock-free insertion relies
// on.  The DRAM-resident regime is bound by random line transactions (ms_debug_random_access_peaks: ~18 G lines/s
// on a B200 whatever is asked of each line), so nothing is requested speculatively -- probing h and h+1 together was
// measured and doubled the traversal time there -- and the table is sized for a low load factor instead, because a
// warp waits for the slowest of its 32 lookups.
// Returns -1 (and raises the overflow flag) when the probe limit is reached.
// Returns the GLOBAL slot id (= the slot on one GPU).
__device__ __forceinline__ long long md_find_regrets(const MdDev& t, unsigned long long key, double* reg, uint32_t& n_ins) {
    unsigned long long h = md_hash(key, t.shift);
    const int owner = md_owner(t, key);
    MdSlot* base = t.peer_slots[owner];
    for (unsigned int probe = 0; probe < t.max_probe; probe++) {
        MdSlot* a = base + h;
        unsigned long long ka = md_ld(t, &a->key);
        const double2 a01 = md_ld(t, (const double2*)&a->regret[0]), a23 = md_ld(t, (const double2*)&a->regret[2]);
        bool fresh = false;
        if (ka == 0ull) { ka = t.world > 1 ? atomicCAS_system(&a->key, 0ull, key) : atomicCAS(&a->key, 0ull, key); fresh = ka == 0ull; }
        if (fresh || ka == key) {
            n_ins += fresh ? 1u : 0u;
            reg[0] = fresh ? 0.0 : a01.x; reg[1] = fresh ? 0.0 : a01.y;
            reg[2] = fresh ? 0.0 : a23.x; reg[3] = fresh ? 0.0 : a23.y;
            return ((long long)owner << t.lshift) | (long long)h;
        }
        h = (h + 1ull) & t.mask;
    }
    t.counters[4] = 1ull;
    reg[0] = reg[1] = reg[2] = reg[3] = 0.0;
    return -1;
}

// regret matching over the columns of the table (ascending card id): mc_cfr.py:20-24
__device__ __forceinline__ void md_regret_match(const double* reg, int n, double* out) {
    double pos[4];
    double norm = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        pos[i] = (i < n && reg[i] > 0.0) ? reg[i] : 0.0;
        if (i < n) norm = __dadd_rn(norm, pos[i]);
    }
    const double uni = __ddiv_rn(1.0, (double)n);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = (i < n) ? (norm > 0.0 ? ms_ddiv_or_zero(pos[i], norm) : uni) : 0.0;
}

// column of legal action k: how many cards of the hand have a smaller id
__device__ __forceinline__ int md_col(uint32_t hand, uint32_t list, int k) {
    const uint32_t c = (list >> (4 * k)) & 0xFu;
    return __popc(hand & ((1u << c) - 1u));
}
__device__ __forceinline__ double md_pick(const double* v, int col) {
    return col == 0 ? v[0] : (col == 1 ? v[1] : (col == 2 ? v[2] : v[3]));
}

// The traversal's form of regret matching: sigma_k = pos[column of action k] * (1 / sum of pos), one division per
// node instead of one per action (md_regret_match above, used by the apply step, divides; the two agree to an
// ulp, far inside the 1e-9 the parity tests ask for).  Uniform when no regret is positive.
__device__ __forceinline__ void md_sigma(const double* reg, int n, uint32_t hand, uint32_t list, double* sg) {
    double pos[4];
    double norm = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        pos[i] = (i < n && reg[i] > 0.0) ? reg[i] : 0.0;
        norm = __dadd_rn(norm, pos[i]);
    }
    const bool any = norm > 0.0;
    const double inv = __ddiv_rn(1.0, any ? norm : 1.0);
    const double uni = n == 2 ? 0.5 : (n == 3 ? (1.0 / 3.0) : (n == 4 ? 0.25 : 1.0));
#pragma unroll
    for (int k = 0; k < 4; k++) sg[k] = (k < n) ? (any ? __dmul_rn(md_pick(pos, md_col(hand, list, k)), inv) : uni) : 0.0;
}

// np.random.choice(legal, p=sigma): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right') -- with the
// normalisation moved to the other side of the comparison (cdf_i <= u * cdf_last)
__device__ __forceinline__ int md_sample(const double* sg, int n, double u) {
    double cdf[4];
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) { if (i < n) acc = __dadd_rn(acc, sg[i]); cdf[i] = acc; }
    const double thr = __dmul_rn(u, acc);
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < n && cdf[i] <= thr) idx++;
    return idx < n ? idx : n - 1;
}

// Traversal frames: SoA in dynamic shared memory, [frame][thread], addressed from the compile-time CTA size so that
// a frame access is one 32-bit offset computation (six 64-bit base pointers cost the kernel its register budget).
#ifndef md_smem   // tests/emu/ms_multideal_host.cpp maps the name onto the emulated block's shared-memory buffer
extern __shared__ __align__(16) unsigned char md_smem[];
#endif
template <int T>
struct MdFrames {
    int tid;
    static constexpr int kRo = 16 * T * MD_FRAMES, kSp = kRo + 8 * T * MD_FRAMES, kSig = kSp + 8 * T * MD_FRAMES,
                         kMeta = kSig + 32 * T * MD_SIG_FRAMES, kCfv = kMeta + 8 * T * MD_FRAMES,
                         kHo = kCfv + 4 * T * MD_FRAMES, kDealt = kHo + 4 * T, kBytes = kDealt + 4 * T;
    __device__ __forceinline__ uint4& st(int fi) const { return ((uint4*)md_smem)[fi * T + tid]; }
    __device__ __forceinline__ double& ro(int fi) const { return ((double*)(md_smem + kRo))[fi * T + tid]; }
    __device__ __forceinline__ double& sp(int fi) const { return ((double*)(md_smem + kSp))[fi * T + tid]; }
    __device__ __forceinline__ double& sig(int fi, int k) const { return ((double*)(md_smem + kSig))[(fi * 4 + k) * T + tid]; }
    __device__ __forceinline__ uint2& meta(int fi) const { return ((uint2*)(md_smem + kMeta))[fi * T + tid]; }
    __device__ __forceinline__ uint32_t& cfv(int fi) const { return ((uint32_t*)(md_smem + kCfv))[fi * T + tid]; }
    // per-traversal constants read at every node: kept here rather than in registers the kernel does not have
    // (a spilled copy costs an L2 round trip: the 28 KB of L1 left beside the frames does not hold the spill slots)
    __device__ __forceinline__ uint32_t& hand_order() const { return ((uint32_t*)(md_smem + kHo))[tid]; }
    __device__ __forceinline__ uint32_t& dealt() const { return ((uint32_t*)(md_smem + kDealt))[tid]; }
};

struct MdCounts { uint32_t upd, vis, step, ins; };   // per thread and launch: a thread runs far fewer than 2^32 visits

// MCCFRTrainer._sample (mc_cfr.py:37-86) from the root of one deal, as an explicit depth-first search (the
// recursion shape, the Philox addressing by call index and the forced-endgame shortcut are those of
// mccfr_traverse in ms_solver.cu; what differs is where the table lives).
template <int T>
__device__ void md_traverse(const MdDev& t, const MsState root, int tp,
                            unsigned long long trav, uint2 pkey, const MdFrames<T>& f, MdCounts& c) {
    MsState s = root;
    double ro = 1.0, sp = 1.0;
    int fi = -1;
    uint4 xblk = make_uint4(0u, 0u, 0u, 0u);
    uint32_t xblk_id = 0xFFFFFFFFu;
    bool pend = false, returning = false;
    uint32_t pend_a = 0u, call = 0u;
    int ret_x2 = 0;
    const uint32_t tag = MS_TAG_MCCF + (uint32_t)tp;
    while (true) {
        if (!returning) {
            if (pend) { step(s, pend_a, table_set_from_dealt(s, f.dealt())); c.step++; pend = false; }
            const uint32_t my_call = call++;
            c.vis++;
            if (st_terminal(s)) {
                const int r = reward0_x2(s);
                ret_x2 = (tp == 0) ? r : -r;
                returning = true;
                continue;
            }
            const int p = st_cur(s);
            uint32_t list;
            const uint32_t nl = legal_list(s, f.hand_order(), p, list);
            if (nl == 1u) {
                // Forced move.  Infosets with one card in hand are NOT stored: their strategy is the constant
                // [1.0], their regret stays 0 and the reference's strategy_sum there is its visit count, which
                // no consumer reads (the policy is [1.0] either way).  They are about half of all node visits.
                const uint32_t a1 = list & 0xFu;
                if (p != tp) { pend_a = a1; pend = true; continue; }
                // traverser's last card: if the rest of the game is forced too, both recursive calls of the
                // reference walk the same line -- played once, accounted twice (see ms_solver.cu)
                MsState t2 = s;
                int below = 0;
                bool forced = false;
                uint32_t act = a1;
#pragma unroll 1
                for (int k = 0; k < 2; k++) {
                    step(t2, act, table_set_from_dealt(t2, f.dealt()));
                    below++;
                    if (st_terminal(t2)) { forced = true; break; }
                    if (k == 1 || __popc(st_hand(t2, p ^ 1)) != 1) break;
                    uint32_t l2;
                    legal_list(t2, f.hand_order(), p ^ 1, l2);
                    act = l2 & 0xFu;
                }
                if (forced) {
                    const int r = reward0_x2(t2);
                    ret_x2 = (tp == 0) ? r : -r;
                    c.upd++;                  // regret delta = w * 0 exactly; strategy_sum += 1.0 (not stored)
                    c.vis += 2 * below; call += 2u * (uint32_t)below; c.step += below;
                    returning = true;
                    continue;
                }
                // not forced (cannot happen from an 8-ply root, kept for generality): general path below
                fi++;
                if (fi >= MD_FRAMES) { t.counters[4] = 2ull; return; }
                f.st(fi) = s; f.ro(fi) = ro; f.sp(fi) = sp;
                f.meta(fi) = make_uint2(0xFFFFFFFFu, list | (1u << 27));
                f.cfv(fi) = 0u;
                pend_a = a1; pend = true;
                continue;
            }
            double reg[4], sg[4];
            const long long slot = md_find_regrets(t, infoset_key(s, p), reg, c.ins);
            md_sigma(reg, (int)nl, st_hand(s, p), list, sg);
            if ((my_call >> 1) != xblk_id) {
                xblk_id = my_call >> 1;
                xblk = philox4x32_10(make_uint4((uint32_t)trav, (uint32_t)(trav >> 32), xblk_id, tag), pkey);
            }
            const double u = (my_call & 1u) ? u53(xblk.z, xblk.w) : u53(xblk.x, xblk.y);
            const int ai = md_sample(sg, (int)nl, u);
            const uint32_t a = (list >> (4 * ai)) & 0xFu;
            const double sga = ai == 0 ? sg[0] : (ai == 1 ? sg[1] : (ai == 2 ? sg[2] : sg[3]));
            if (p != tp) {                // opponent: reach *= sigma[a]; tail call (mc_cfr.py:63-65)
                ro = __dmul_rn(ro, sga);
                pend_a = a; pend = true;
                continue;
            }
            fi++;
            if (fi >= MD_SIG_FRAMES) { t.counters[4] = 2ull; return; }
            f.st(fi) = s; f.ro(fi) = ro; f.sp(fi) = sp;
            f.meta(fi) = make_uint2((uint32_t)slot, list | (nl << 27));
            f.cfv(fi) = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) f.sig(fi, k) = sg[k];
            sp = __dmul_rn(sp, sga);
            pend_a = a; pend = true;
            continue;
        }
        // ---- a child returned ret_x2 to the top frame
        if (fi < 0) break;
        uint2 meta = f.meta(fi);
        const int nl = (int)((meta.y >> 27) & 0x7u);
        int cur = (int)((meta.y >> 24) & 0x7u);
        uint32_t cfvb = f.cfv(fi);
        if (cur == 0) meta.y = (meta.y & 0xFF00FFFFu) | (((uint32_t)ret_x2 & 0xFFu) << 16);   // util of the sampled action
        else cfvb |= ((uint32_t)ret_x2 & 0xFFu) << (8 * (cur - 1));
        cur++;
        double sg[4];
        if (nl > 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) sg[k] = f.sig(fi, k);
        } else { sg[0] = 1.0; sg[1] = sg[2] = sg[3] = 0.0; }
        if (cur <= nl) {                  // evaluate action i = cur-1 with a fresh sampled continuation (:71-78)
            const int i = cur - 1;
            meta.y = (meta.y & 0xF8FFFFFFu) | ((uint32_t)cur << 24);
            f.meta(fi) = meta; f.cfv(fi) = cfvb;
            s = f.st(fi);
            ro = f.ro(fi);
            sp = __dmul_rn(f.sp(fi), i == 0 ? sg[0] : (i == 1 ? sg[1] : (i == 2 ? sg[2] : sg[3])));
            pend_a = (meta.y >> (4 * i)) & 0xFu; pend = true;
            returning = false;
            continue;
        }
        // ---- all actions evaluated: regret / strategy deltas (:79-84)
        const long long slot = (long long)(int)meta.x;
        if (nl > 1 && slot >= 0) {        // |A| = 1: cfv - v == 0 exactly
            double cfv[4];
            double v = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                cfv[i] = 0.5 * (double)(int)(int8_t)((cfvb >> (8 * i)) & 0xFFu);
                if (i < nl) v = __dadd_rn(v, __dmul_rn(sg[i], cfv[i]));
            }
            const double fro = f.ro(fi), fsp = f.sp(fi);
            const double w = fsp > 0.0 ? __ddiv_rn(fro, fsp) : 0.0;
            const uint32_t hand = st_hand(f.st(fi), tp);
            const uint32_t list = meta.y & 0xFFFFu;
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < nl) md_add(t, &md_slot(t, slot)->delta[md_col(hand, list, i)], __dmul_rn(w, __dadd_rn(cfv[i], -v)));
        }
        if (slot >= 0) {
            md_add(t, &md_slot(t, slot)->cnt, 1u);
            md_mark_dirty(t, slot);
        }
        c.upd++;
        ret_x2 = (int)(int8_t)((meta.y >> 16) & 0xFFu);
        fi--;
        returning = true;
    }
}


template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) md_mccfr_kernel(MdDev t, int player, long long n_trav, uint2 pkey,
                                                                 unsigned long long first_trav) {
    constexpr int T = THREADS;
    const int tid = threadIdx.x;
    const MdFrames<T> f{tid};
    MdCounts c{0u, 0u, 0u, 0u};
    const long long gstride = (long long)gridDim.x * T;
    for (long long k = blockIdx.x * (long long)T + tid; k < n_trav; k += gstride) {
        const unsigned long long trav = first_trav + (unsigned long long)k;
        const uint4 x = philox4x32_10(make_uint4((uint32_t)trav, (uint32_t)(trav >> 32), 0u, MS_TAG_DEAL), pkey);
        const uint32_t deal = __umulhi(x.x, t.n_deals);
        const MsState root = t.roots[deal];
        f.hand_order() = t.hand_order[deal];
        f.dealt() = dealt_set(root);
        for (int tp = 0; tp < 2; tp++) {
            if (player < 2 && tp != player) continue;
            md_traverse<T>(t, root, tp, trav, pkey, f, c);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        c.upd += __shfl_down_sync(0xffffffffu, c.upd, off);
        c.vis += __shfl_down_sync(0xffffffffu, c.vis, off);
        c.step += __shfl_down_sync(0xffffffffu, c.step, off);
        c.ins += __shfl_down_sync(0xffffffffu, c.ins, off);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&t.counters[0], c.upd); atomicAdd(&t.counters[1], c.vis);
        atomicAdd(&t.counters[2], c.step);
        if (c.ins) atomicAdd(&t.counters[3], c.ins);
    }
}

// after a batch: strategy_sum += cnt * sigma(frozen regrets); regret += delta; delta = 0; cnt = 0.
// Walks the dirty bitmap (one thread per 32 slots), so the cost follows the slots a batch updated, not the capacity.
__global__ void __launch_bounds__(256) md_apply_kernel(MdDev t) {
    const unsigned long long words = (t.mask + 1ull) >> 5;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long wi = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; wi < words; wi += stride) {
        unsigned int w = t.dirty[wi];
        if (w == 0u) continue;
        t.dirty[wi] = 0u;
        while (w) {
            const int bit = __ffs((int)w) - 1;
            w &= w - 1u;
            MdSlot* sl = t.slots + (wi << 5) + bit;
            const uint4 head = *(const uint4*)sl;                 // key (x, y) | cnt (z)
            const unsigned int cnt = head.z;
            const int n = __popc((head.y >> 4) & 0xFFFFu);        // hand mask = key bits 36-51
            double reg[4], sg[4];
#pragma unroll
            for (int i = 0; i < 4; i++) reg[i] = sl->regret[i];
            md_regret_match(reg, n, sg);
            const double cn = (double)cnt;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (i < n) {
                    sl->strategy[i] = __dadd_rn(sl->strategy[i], __dmul_rn(cn, sg[i]));
                    const double dl = sl->delta[i];
                    if (dl != 0.0) { sl->regret[i] = __dadd_rn(reg[i], dl); sl->delta[i] = 0.0; }
                }
            }
            sl->cnt = 0u;
        }
    }
}

__global__ void __launch_bounds__(256) md_export_kernel(MdDev t, unsigned long long* keys, double* regret,
                                                        double* strategy, long long max_n, unsigned long long* n_out) {
    const unsigned long long cap = t.mask + 1ull;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long h = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; h < cap; h += stride) {
        const MdSlot* sl = t.slots + h;
        const unsigned long long key = sl->key;
        if (key == 0ull) continue;
        const unsigned long long i = atomicAdd(n_out, 1ull);
        if ((long long)i >= max_n) continue;
        keys[i] = key;
#pragma unroll
        for (int a = 0; a < 4; a++) { regret[4 * i + a] = sl->regret[a]; strategy[4 * i + a] = sl->strategy[a]; }
    }
}

__global__ void __launch_bounds__(256) md_lookup_kernel(MdDev t, const unsigned long long* __restrict__ keys, long long n,
                                                        double* regret, double* strategy, uint8_t* found) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = keys[i];
        unsigned long long h = md_hash(key, t.shift);
        const MdSlot* base = t.peer_slots[md_owner(t, key)];      // any rank can look any infoset up
        long long slot = -1;
        for (unsigned int probe = 0; probe < t.max_probe; probe++) {
            const unsigned long long k = md_ld(t, &base[h].key);
            if (k == key) { slot = (long long)h; break; }
            if (k == 0ull) break;
            h = (h + 1ull) & t.mask;
        }
        if (key == 0ull) slot = -1;
        if (found) found[i] = slot >= 0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            if (regret) regret[4 * i + a] = slot >= 0 ? md_ld(t, &base[slot].regret[a]) : 0.0;
            if (strategy) strategy[4 * i + a] = slot >= 0 ? md_ld(t, &base[slot].strategy[a]) : 0.0;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Random-access ceilings of the table's access pattern, measured on the box (there is no published figure):
// 148 x 768 threads touch pseudo-random 128-byte lines of a 2^k-line buffer.
//   [0] dependent reads: the next address depends on the loaded data (one access in flight per thread -- the
//       shape of a depth-first traversal), 64 bytes (key sector + regret sector) per access;
//   [1] independent reads, 8 in flight per thread, 64 bytes per access (the DRAM / L2 random-access ceiling);
//   [2] fp64 RED.ADD x4 into one sector of a random line (the regret-delta update).
__global__ void __launch_bounds__(768, 1) rnd_dependent_kernel(const MdSlot* tab, unsigned long long mask, int iters,
                                                               unsigned long long* sink) {
    unsigned long long x = (blockIdx.x * 768ull + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345ull;
    unsigned long long acc = 0;
    for (int i = 0; i < iters; i++) {
        const MdSlot* sl = tab + ((x >> 20) & mask);
        const unsigned long long k = __ldcg(&sl->key);
        const double2 r = __ldcg((const double2*)&sl->regret[0]);
        acc += k + (unsigned long long)__double_as_longlong(r.x);
        x = (x + k + 1ull) * 0xD1342543DE82EF95ull + 1442695040888963407ull;   // k is 0 in a zeroed buffer, but unknown to the compiler
    }
    if (acc == 0x123456789ull) *sink = acc;
}
__global__ void __launch_bounds__(768, 1) rnd_independent_kernel(const MdSlot* tab, unsigned long long mask, int iters,
                                                                 unsigned long long* sink) {
    unsigned long long x = (blockIdx.x * 768ull + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345ull;
    unsigned long long acc = 0;
    for (int i = 0; i < iters; i += 8) {
        unsigned long long k[8]; double2 r[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            x = x * 0xD1342543DE82EF95ull + 1442695040888963407ull;
            const MdSlot* sl = tab + ((x >> 20) & mask);
            k[j] = __ldcg(&sl->key);
            r[j] = __ldcg((const double2*)&sl->regret[0]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += k[j] + (unsigned long long)__double_as_longlong(r[j].x);
    }
    if (acc == 0x123456789ull) *sink = acc;
}
__global__ void __launch_bounds__(768, 1) rnd_red_kernel(MdSlot* tab, unsigned long long mask, int iters) {
    unsigned long long x = (blockIdx.x * 768ull + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345ull;
    for (int i = 0; i < iters; i++) {
        x = x * 0xD1342543DE82EF95ull + 1442695040888963407ull;
        MdSlot* sl = tab + ((x >> 20) & mask);
#pragma unroll
        for (int a = 0; a < 4; a++) atomicAdd(&sl->delta[a], 1.0);
    }
}

// ------------------------------------------------------------------------------------------------
// Deal-blocked form: the HBM table as the backing store of an on-chip solver.
//
// md_mccfr_kernel above samples a deal per traversal and touches the table in place: every node visit is a random
// line transaction, and the dependent chain of a depth-first traversal keeps it latency-bound (section 11 of
// DESIGN.md).  Here a CTA takes one deal at a time and runs a whole block of traversals on it (`pairs_per_visit`,
// thousands), exactly like the one-deal solver: the deal's enumerated tree, the frozen strategies of its infosets,
// their cdfs and the private delta tables live in shared memory (the walk is StaticWalk<.., MODE 1> of
// ms_static_walk.cuh), and the HBM table is read once when the visit starts (one gather of the deal's <= 501 infoset lines) and
// written once when it ends (REDs of the non-zero deltas and visit counts).  Random-line traffic per traversal drops
// by three orders of magnitude.  Chance sampling is per visit: visit v plays deal
// mulhi32(x0, D), x = Philox4x32-10(key = seed, ctr = (v lo, v hi, 1, "DEAL")); its traversals have the global
// ids v * pairs_per_visit + i and use the sequential "MCCF"+64 streams of the one-deal solver's headline kernel.
//
// Fresh 4+4-card deals all have the same tree SHAPE (levels of 1, 4, 16, 48, 144, 288, 576, 576, 576 nodes with 4, 4,
// 3, 3, 2, 2, 1, 1 legal actions; node i of a level has children begin[L+1] + i * n_legal + a, a in legal_actions()
// order), so a deal is described by: the local infoset index of each of its 501 multi-action nodes, the table slot,
// action count and column permutation (legal order -> ascending card id) of each local infoset, and the 576
// terminal rewards.  md_build_kernel derives that once per deal with the env's step() and claims the table slots.
constexpr int MDB_NODES = 2229, MDB_MULTI = 501, MDB_LEAVES = 576, MDB_FIRST_LEAF = 1653, MDB_LOCAL = 512;
constexpr int MDB_SLOTS = MDB_LOCAL + 2 * 576;      // + one private dummy slot per one-card node (levels 6, 7)
__device__ __constant__ int c_mdb_begin[10] = {0, 1, 5, 21, 69, 213, 501, 1077, 1653, 2229};
__device__ __constant__ int c_mdb_nl[9] = {4, 4, 3, 3, 2, 2, 1, 1, 0};

struct MdDealInfo {
    uint32_t local_slot[MDB_LOCAL];     // bits 0-29 table slot | 30-31 number of actions - 2
    uint16_t node_local[MDB_LOCAL];     // nodes 0..500 -> local infoset index
    uint8_t local_perm[MDB_LOCAL];      // 2 bits per legal action: its column (rank of the card id in the hand)
    int8_t rx2[MDB_LEAVES];             // 2 * reward of player 0 at the 576 leaves
    uint16_t acc_row[MDB_LOCAL];        // first regret-accumulator row of each local infoset (it owns nl - 1 consecutive rows)
    uint32_t n_local;
    uint32_t n_acc;                     // accumulator rows of the deal (<= 575)
    uint32_t pad[2];
};
static_assert(sizeof(MdDealInfo) % 16 == 0, "deal records are copied with 128-bit accesses");

__global__ void __launch_bounds__(256) md_build_kernel(MdDev t, MdDealInfo* __restrict__ info) {
    __shared__ uint4 lvl[2][576];
    __shared__ unsigned long long hkey[1024];
    __shared__ uint16_t hloc[1024];
    __shared__ uint8_t hmeta[1024];          // perm of the infoset claimed at this hash position
    __shared__ uint8_t hnl[1024];
    __shared__ int part[257], part2[257];
    const int tid = threadIdx.x;
    for (unsigned int deal = blockIdx.x; deal < t.n_deals; deal += gridDim.x) {
        MdDealInfo* out = info + deal;
        const uint32_t ho = t.hand_order[deal];
        for (int i = tid; i < 1024; i += 256) hkey[i] = 0ull;
        if (tid == 0) lvl[0][0] = t.roots[deal];
        __syncthreads();
        for (int L = 0; L < 8; L++) {
            const uint4* cur = lvl[L & 1];
            uint4* nxt = lvl[(L + 1) & 1];
            const int W = c_mdb_begin[L + 1] - c_mdb_begin[L], nl = c_mdb_nl[L], pl = L & 1;
            for (int i = tid; i < W; i += 256) {
                const MsState s = cur[i];
                uint32_t list;
                legal_list(s, ho, pl, list);
                if (L < 6) {
                    const unsigned long long key = infoset_key(s, pl);
                    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 54);
                    while (true) {
                        const unsigned long long k = atomicCAS(&hkey[h], 0ull, key);
                        if (k == 0ull || k == key) break;
                        h = (h + 1u) & 1023u;
                    }
                    uint32_t perm = 0u;
                    for (int a = 0; a < nl; a++) perm |= (uint32_t)md_col(st_hand(s, pl), list, a) << (2 * a);
                    hmeta[h] = (uint8_t)perm; hnl[h] = (uint8_t)nl;       // every node of an infoset writes the same
                    out->node_local[c_mdb_begin[L] + i] = (uint16_t)h;    // hash position for now, local index below
                }
                for (int a = 0; a < nl; a++) {
                    MsState c = s;
                    step(c, (list >> (4 * a)) & 0xFu);
                    nxt[i * nl + a] = c;
                    if (L == 7) out->rx2[i * nl + a] = (int8_t)reward0_x2(c);
                }
            }
            __syncthreads();
        }
        // hash positions -> dense local indices, in position order
        int mine = 0, mine2 = 0;
        for (int j = 0; j < 4; j++)
            if (hkey[4 * tid + j] != 0ull) { mine++; mine2 += (int)hnl[4 * tid + j] - 1; }
        part[tid + 1] = mine; part2[tid + 1] = mine2;
        if (tid == 0) { part[0] = 0; part2[0] = 0; }
        __syncthreads();
        if (tid == 0) for (int j = 1; j <= 256; j++) { part[j] += part[j - 1]; part2[j] += part2[j - 1]; }
        __syncthreads();
        int idx = part[tid], row = part2[tid];
        uint32_t ins = 0u;
        for (int j = 0; j < 4; j++) {
            const int h = 4 * tid + j;
            if (hkey[h] != 0ull) {
                hloc[h] = (uint16_t)idx;
                double unused[4];
                const long long slot = md_find_regrets(t, hkey[h], unused, ins);     // claims the slot on first sight
                out->local_slot[idx] = ((uint32_t)(slot < 0 ? 0 : slot) & 0x3FFFFFFFu) | ((uint32_t)(hnl[h] - 2) << 30);
                out->local_perm[idx] = hmeta[h];
                out->acc_row[idx] = (uint16_t)row;
                row += (int)hnl[h] - 1;
                idx++;
            }
        }
        if (ins) atomicAdd(&t.counters[3], (unsigned long long)ins);
        if (tid == 0) { out->n_local = (uint32_t)part[256]; out->n_acc = (uint32_t)part2[256]; }
        __syncthreads();
        for (int n = tid; n < MDB_MULTI; n += 256) out->node_local[n] = hloc[out->node_local[n]];
        __syncthreads();
    }
}

constexpr int MDB_THREADS = 1024;
constexpr int MDB_ACC_MAX = 3 * (1 + 4) + 2 * (16 + 48) + (144 + 288);      // nl - 1 rows per infoset, every node its own infoset

__host__ __device__ inline size_t mdb_smem_bytes() {
    return 16 * (size_t)MDB_LOCAL + sizeof(double) * 8 * MDB_LOCAL + sizeof(double) * 32 * MDB_ACC_MAX + 4 * MDB_LOCAL +
           sizeof(MdDealInfo) + 64;
}

// Round 2: the walk is ms_static_walk.cuh's (MODE 1) -- the estimator's recursion written out as nested loops with its
// state in registers, integer cdf thresholds, the sequential Philox stream, lane-private nl - 1 regret accumulators --
// instead of the generic DFS with frames in shared memory: the same change that took the one-deal solver from 95 to 650 G
// updates/s.  One-card infosets are not stored in this table, so the forced endgames count nothing and a ply-5 record
// carries just the two leaf rewards.
__global__ void __launch_bounds__(MDB_THREADS, 1) md_blocked_kernel(MdDev t, const MdDealInfo* __restrict__ info, int player,
                                                                  unsigned long long first_visit, long long n_visits,
                                                                  int pairs_per_visit, uint2 pkey) {
    MS_DYN_SMEM(smem_raw);
    constexpr int T = MDB_THREADS, SL = MDB_LOCAL;
    const int tid = threadIdx.x, lane = tid & 31;
    uint4* node = (uint4*)smem_raw;                        // [SL] records of the 501 multi-action nodes
    double* sig = (double*)(node + SL);                    // [SL][4] frozen strategies, legal (= deal) order
    double* rsig = sig + 4 * SL;                           // [SL][4] 1 / sigma (0 where sigma == 0)
    double* acc = rsig + 4 * SL;                           // [MDB_ACC_MAX][32] lane-private D accumulators
    uint32_t* dcnt = (uint32_t*)(acc + 32 * MDB_ACC_MAX);  // [SL] update counts per local infoset
    MdDealInfo* di = (MdDealInfo*)(dcnt + SL);
    uint32_t* thr = (uint32_t*)acc;                        // [SL][3] staging of the thresholds (acc is zeroed afterwards)

    StaticShared c;
    c.node = node; c.sig = sig; c.rsig = rsig; c.acc = acc + lane; c.accrow = di->acc_row; c.dcnt = dcnt; c.touched = nullptr;
    c.key = pkey; c.blk = make_uint4(0u, 0u, 0u, 0u);
    const StaticDims dm{};                                 // (MODE 1 takes its rows from c.accrow)
    unsigned long long v0, u0, e0, v1, u1, e1;
    static_shape_counts(0, v0, u0, e0);
    static_shape_counts(1, v1, u1, e1);
    unsigned long long nu = 0, nv = 0, ns = 0;
    const int flip = (tid >> 5) & 1;

    for (long long vi = blockIdx.x; vi < n_visits; vi += gridDim.x) {
        const unsigned long long visit = first_visit + (unsigned long long)vi;
        const uint4 x = philox4x32_10(make_uint4((uint32_t)visit, (uint32_t)(visit >> 32), 1u, MS_TAG_DEAL), pkey);
        const uint32_t deal = __umulhi(x.x, t.n_deals);
        __syncthreads();                                   // the previous visit is flushed
        {   // stage the deal record
            const uint4* src = (const uint4*)(info + deal);
            uint4* dst = (uint4*)di;
            for (int i = tid; i < (int)(sizeof(MdDealInfo) / 16); i += T) dst[i] = src[i];
        }
        for (int i = tid; i < SL; i += T) dcnt[i] = 0u;
        __syncthreads();
        const int n_local = (int)di->n_local, n_acc = (int)di->n_acc;
        for (int j = tid; j < n_local; j += T) {           // gather: one table line per infoset of the deal
            const uint32_t ls = di->local_slot[j];
            const MdSlot* sl = md_slot(t, (long long)(ls & 0x3FFFFFFFu));     // on a sharded table: 7 of 8 lines over NVLink
            const int nl = (int)(ls >> 30) + 2;
            const double2 r01 = md_ld(t, (const double2*)&sl->regret[0]), r23 = md_ld(t, (const double2*)&sl->regret[2]);
            const double reg[4] = {r01.x, r01.y, r23.x, r23.y};
            double sc[4], sg[4], cd[4];
            md_regret_match(reg, nl, sc);                  // over the table's columns, like the apply step
            const uint32_t perm = di->local_perm[j];
#pragma unroll
            for (int k = 0; k < 4; k++) sg[k] = (k < nl) ? md_pick(sc, (int)((perm >> (2 * k)) & 3u)) : 0.0;
            strategy_cdf(sg, nl, cd);
#pragma unroll
            for (int k = 0; k < 4; k++) { sig[4 * j + k] = sg[k]; rsig[4 * j + k] = sg[k] > 0.0 ? __ddiv_rn(1.0, sg[k]) : 0.0; }
#pragma unroll
            for (int k = 0; k < 3; k++) thr[3 * j + k] = (k + 1 < nl) ? (uint32_t)ceil(cd[k] * 2147483648.0) : 0x80000000u;
        }
        __syncthreads();
        for (int n = tid; n < MDB_MULTI; n += T) {         // node records (ms_static_walk.cuh)
            int L = 0;
            while (n >= c_mdb_begin[L + 1]) L++;
            const int i = n - c_mdb_begin[L], nl = c_mdb_nl[L], j = (int)di->node_local[n];
            const uint32_t link = (uint32_t)(c_mdb_begin[L + 1] + i * nl) | ((uint32_t)j << 12);
            if (L < 5) node[n] = make_uint4(thr[3 * j], thr[3 * j + 1], thr[3 * j + 2], link);
            else {          // ply 5: child k is ply-6 node 2 i + k of its level, whose forced line ends at leaf 2 i + k
                const uint32_t e0r = (uint32_t)((int)di->rx2[2 * i] + 16) << 22, e1r = (uint32_t)((int)di->rx2[2 * i + 1] + 16) << 22;
                node[n] = make_uint4(thr[3 * j], link, e0r, e1r);
            }
        }
        __syncthreads();
        for (int i = tid; i < 32 * n_acc; i += T) acc[i] = 0.0;
        __syncthreads();
        for (int k = tid; k < pairs_per_visit; k += T) {
            const unsigned long long trav = visit * (unsigned long long)pairs_per_visit + (unsigned long long)k;
            c.t_lo = (uint32_t)trav; c.t_hi = (uint32_t)(trav >> 32);
            for (int j = 0; j < 2; j++) {
                const int tp = j ^ flip;
                if (player < 2 && tp != player) continue;
                c.nd = 0u; c.tag = MS_TAG_MCCF_SEQ + (uint32_t)tp;
                if (tp == 0) { StaticWalk<0, 0, false, 1>::run(0u, 1.0, c, dm); nu += u0; nv += v0; ns += e0; }
                else { StaticWalk<0, 1, false, 1>::run(0u, 1.0, c, dm); nu += u1; nv += v1; ns += e1; }
            }
        }
        __syncthreads();
        for (int j = tid; j < n_local; j += T) {           // scatter: deltas and visit counts back to the table
            const unsigned int cnt = dcnt[j];
            if (cnt == 0u) continue;
            const uint32_t ls = di->local_slot[j];
            const long long slot = (long long)(ls & 0x3FFFFFFFu);
            const int nl = (int)(ls >> 30) + 2;
            const uint32_t perm = di->local_perm[j];
            // delta_a = D_a - sum_i sigma_i D_i with D_last = 0 (ms_static_walk.cuh)
            double D0 = 0.0, D1 = 0.0, D2 = 0.0, sum = 0.0;
            for (int i = 0; i < nl - 1; i++) {
                const double* col = acc + 32 * (size_t)((int)di->acc_row[j] + i);
                double tt = 0.0;
                for (int l = 0; l < 32; l++) tt = __dadd_rn(tt, col[l]);
                if (i == 0) D0 = tt; else if (i == 1) D1 = tt; else D2 = tt;
                sum = __dadd_rn(sum, __dmul_rn(sig[4 * j + i], tt));
            }
            for (int k = 0; k < nl; k++) {
                const double Dk = k == 0 ? D0 : (k == 1 ? D1 : D2);
                const double v = k < nl - 1 ? __dadd_rn(Dk, -sum) : -sum;
                if (v != 0.0) md_add(t, &md_slot(t, slot)->delta[(perm >> (2 * k)) & 3u], v);
            }
            md_add(t, &md_slot(t, slot)->cnt, cnt);
            md_mark_dirty(t, slot);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        nu += __shfl_down_sync(0xffffffffu, nu, off);
        nv += __shfl_down_sync(0xffffffffu, nv, off);
        ns += __shfl_down_sync(0xffffffffu, ns, off);
    }
    if ((tid & 31) == 0) { atomicAdd(&t.counters[0], nu); atomicAdd(&t.counters[1], nv); atomicAdd(&t.counters[2], ns); }
}

// ------------------------------------------------------------------------------------------------
// Barrier between the GPUs that share a sharded table.  One iteration is
//     md_blocked_kernel (gather of frozen regrets and REDs of deltas, both through peer memory)
//     -> barrier -> md_apply_kernel (every rank folds the deltas of ITS shard in) -> barrier -> next iteration,
// and the barrier is this kernel, stream-ordered between them: one thread per peer fences at system scope, writes the
// epoch into its slot of that peer's flag array (st.release.sys) and waits (bounded) until its own array shows the
// epoch from every peer (ld.acquire.sys).  The kernel before it on the stream has completed, so its peer stores and
// REDs are performed before the flags are written; a rank that passes the barrier therefore sees every rank's updates.
// err[0]: 0 = fine, else 1 + the first peer that did not arrive in time; sticky (ms_md_peer_error).
struct MdPeerSync {
    unsigned long long* flags[MD_MAX_PEERS];    // rank r's flag array ([MD_MAX_PEERS] u64), mapped here
    unsigned long long* my_flags;
    int rank, world;
};
#ifndef MS_CTA_EMU
__device__ __forceinline__ void md_peer_signal(unsigned long long* flag, unsigned long long epoch) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(epoch) : "memory");
}
__device__ __forceinline__ unsigned long long md_peer_poll(const unsigned long long* flag) {
    unsigned long long seen;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
    return seen;
}
__device__ __forceinline__ unsigned long long md_peer_clock_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void md_peer_fence() { __threadfence_system(); }
#endif
constexpr unsigned long long MD_PEER_TIMEOUT_NS = 2000000000ull;

__global__ void __launch_bounds__(32) md_peer_barrier_kernel(MdPeerSync ps, unsigned long long epoch, unsigned int* err) {
    const int tid = threadIdx.x;
    if (tid >= ps.world) return;
    if (*(volatile unsigned int*)err != 0u) return;
    md_peer_fence();
    md_peer_signal(ps.flags[tid] + ps.rank, epoch);
    const unsigned long long t0 = md_peer_clock_ns();
    while (md_peer_poll(ps.my_flags + tid) < epoch) {
        if (md_peer_clock_ns() - t0 > MD_PEER_TIMEOUT_NS) { atomicCAS(err, 0u, 1u + (unsigned)tid); break; }
    }
}

}  // namespace ms

#ifndef MS_HOST_RULES_ONLY   // tests/emu/ms_multideal_host.cpp compiles every kernel above for the host's CTA emulator;
                             // below: the library's host side (CUDA runtime calls, launches, C ABI)
using namespace ms;

struct ms_mdsolver {
    MdDev dev;
    int log2cap;
    int64_t n_deals;
    uint4* d_roots; uint32_t* d_hand_order;
    unsigned long long* d_counters;    // 5 counters + 1 export cursor
    MdDealInfo* d_info;                // per-deal tree descriptions of the deal-blocked form (built on first use)
    // sharded table (ms_md_ipc_export / ms_md_ipc_attach / ms_md_peer_barrier)
    unsigned long long* d_flags;       // [MD_MAX_PEERS] barrier flags, inside the dirty bitmap's allocation (one IPC handle)
    unsigned int* d_peer_err;          // device word set by the barrier when a peer did not arrive
    int attached, rank, world;
    unsigned long long epoch;
    void* peer_slots_base[MD_MAX_PEERS];
    void* peer_aux_base[MD_MAX_PEERS];
};

// the dirty bitmap's allocation: [capacity / 32 words] then, 256-byte aligned, [MD_MAX_PEERS] u64 flags and the error word
static size_t md_aux_flags_off(int log2cap) { return ((((size_t)1 << log2cap) >> 5) * sizeof(unsigned int) + 255) & ~(size_t)255; }
static size_t md_aux_bytes(int log2cap) { return md_aux_flags_off(log2cap) + 256; }

extern "C" {

int ms_md_create(const int64_t* d_seeds, int64_t n_deals, int32_t log2_capacity, void* stream, ms_mdsolver** out) {
    if (!out) return fail(MS_ERR_ARG, "ms_md_create: out is NULL");
    *out = nullptr;
    if (!d_seeds || n_deals < 1 || n_deals > 0x7FFFFFFFll) return fail(MS_ERR_ARG, "ms_md_create: bad deal list");
    if (log2_capacity < 10 || log2_capacity > 30) return fail(MS_ERR_ARG, "ms_md_create: log2_capacity must be in [10, 30]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(MS_ERR_CUDA, "ms_md_create: no CUDA device");
    ms_mdsolver* s = new ms_mdsolver();
    std::memset(s, 0, sizeof(*s));
    s->log2cap = log2_capacity; s->n_deals = n_deals;
    const size_t cap = (size_t)1 << log2_capacity;
    cudaError_t e = cudaMalloc(&s->dev.slots, cap * sizeof(MdSlot));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_roots, (size_t)n_deals * sizeof(uint4));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_hand_order, (size_t)n_deals * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_counters, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&s->dev.dirty, md_aux_bytes(log2_capacity));
    if (e != cudaSuccess) {
        cudaFree(s->dev.slots); cudaFree(s->dev.dirty); cudaFree(s->d_roots); cudaFree(s->d_hand_order); cudaFree(s->d_counters);
        delete s;
        return fail(e == cudaErrorMemoryAllocation ? MS_ERR_CAPACITY : MS_ERR_CUDA, "ms_md_create: cudaMalloc failed: %s",
                    cudaGetErrorString(e));
    }
    s->dev.mask = (unsigned long long)cap - 1ull;
    s->dev.shift = 64 - log2_capacity;
    s->dev.max_probe = cap < 8192 ? (unsigned int)cap : 8192u;
    s->dev.roots = s->d_roots; s->dev.hand_order = s->d_hand_order;
    s->dev.n_deals = (unsigned int)n_deals;
    s->dev.counters = s->d_counters;
    s->dev.world = 1; s->dev.rank = 0; s->dev.lshift = log2_capacity;
    s->dev.peer_slots[0] = s->dev.slots; s->dev.peer_dirty[0] = s->dev.dirty;
    s->world = 1;
    s->d_flags = (unsigned long long*)((char*)s->dev.dirty + md_aux_flags_off(log2_capacity));
    s->d_peer_err = (unsigned int*)(s->d_flags + MD_MAX_PEERS);
    int rc = ms_deal_from_seeds(d_seeds, n_deals, (ms_state*)s->d_roots, s->d_hand_order, stream);
    if (rc == MS_OK) rc = ms_md_reset(s, stream);
    if (rc != MS_OK) { ms_md_destroy(s); return rc; }
    *out = s;
    return MS_OK;
}

void ms_md_destroy(ms_mdsolver* s) {
    if (!s) return;
    if (s->attached)
        for (int r = 0; r < s->world; r++) {
            if (r == s->rank) continue;
            if (s->peer_slots_base[r]) cudaIpcCloseMemHandle(s->peer_slots_base[r]);
            if (s->peer_aux_base[r]) cudaIpcCloseMemHandle(s->peer_aux_base[r]);
        }
    cudaFree(s->dev.slots); cudaFree(s->dev.dirty); cudaFree(s->d_roots); cudaFree(s->d_hand_order); cudaFree(s->d_counters);
    cudaFree(s->d_info);
    delete s;
}

int ms_md_reset(ms_mdsolver* s, void* stream) {
    if (!s) return fail(MS_ERR_ARG, "ms_md_reset: NULL handle");
    // peers read and update this shard at their own pace: clearing it under them would race
    if (s->attached) return fail(MS_ERR_STATE, "ms_md_reset: the table is shared with peers (create a new solver instead)");
    MS_CUDA(cudaMemsetAsync(s->dev.slots, 0, ((size_t)1 << s->log2cap) * sizeof(MdSlot), (cudaStream_t)stream));
    MS_CUDA(cudaMemsetAsync(s->dev.dirty, 0, md_aux_bytes(s->log2cap), (cudaStream_t)stream));
    MS_CUDA(cudaMemsetAsync(s->d_counters, 0, 8 * sizeof(unsigned long long), (cudaStream_t)stream));
    if (s->d_info) {            // the deal descriptions name table slots: they die with the table's keys
        MS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        MS_CUDA(cudaFree(s->d_info));
        s->d_info = nullptr;
    }
    return MS_OK;
}

int ms_md_mccfr_batch(ms_mdsolver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav,
                      void* stream) {
    if (!s) return fail(MS_ERR_ARG, "ms_md_mccfr_batch: NULL handle");
    if (player < 0 || player > 2 || n_trav < 0 || n_trav > 10000000000ll)   // per-warp counters are 32-bit
        return fail(MS_ERR_ARG, "ms_md_mccfr_batch: bad argument");
    if (n_trav == 0) return MS_OK;
    // 768 threads per CTA, one CTA per SM: measured 6.35 ms per 341 k traversal pairs at 65 536 deals against 6.9 ms
    // with 640 or 512 threads (more registers, no spills, fewer chains in flight)
    const int smem = MdFrames<MD_THREADS>::kBytes;
    MS_CUDA(cudaFuncSetAttribute(md_mccfr_kernel<MD_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    md_mccfr_kernel<MD_THREADS><<<grid_for(n_trav, MD_THREADS, 1), MD_THREADS, smem, (cudaStream_t)stream>>>(
        s->dev, player, (long long)n_trav, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
        (unsigned long long)first_trav);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_md_apply(ms_mdsolver* s, void* stream) {
    if (!s) return fail(MS_ERR_ARG, "ms_md_apply: NULL handle");
    const int64_t cap = (int64_t)1 << s->log2cap;
    md_apply_kernel<<<grid_for(cap >> 5, 256, 8), 256, 0, (cudaStream_t)stream>>>(s->dev);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_md_counters(ms_mdsolver* s, uint64_t h_out[5], int reset, void* stream) {
    if (!s || !h_out) return fail(MS_ERR_ARG, "ms_md_counters: bad argument");
    MS_CUDA(cudaMemcpyAsync(h_out, s->d_counters, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (reset) {   // the traffic counters; the infoset count and the overflow flag describe the table and stay
        MS_CUDA(cudaMemsetAsync(s->d_counters, 0, 3 * sizeof(unsigned long long), (cudaStream_t)stream));
    }
    if (h_out[4] == 1) return fail(MS_ERR_CAPACITY, "multi-deal infoset table is full (capacity 2^%d, %llu infosets)",
                                   s->log2cap, (unsigned long long)h_out[3]);
    if (h_out[4] == 2) return fail(MS_ERR_STATE, "multi-deal traversal left the 8-ply game shape");
    return MS_OK;
}

int ms_md_export(ms_mdsolver* s, uint64_t* d_keys, double* d_regret, double* d_strategy, int64_t max_n, int64_t* h_n,
                 void* stream) {
    if (!s || !h_n || max_n < 0 || (max_n > 0 && (!d_keys || !d_regret || !d_strategy)))      // max_n = 0: count only
        return fail(MS_ERR_ARG, "ms_md_export: bad argument");
    MS_CUDA(cudaMemsetAsync(s->d_counters + 5, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    const int64_t cap = (int64_t)1 << s->log2cap;
    md_export_kernel<<<grid_for(cap, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        s->dev, (unsigned long long*)d_keys, d_regret, d_strategy, (long long)max_n, s->d_counters + 5);
    MS_LAUNCH_CHECK();
    unsigned long long n = 0;
    MS_CUDA(cudaMemcpyAsync(&n, s->d_counters + 5, sizeof(n), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *h_n = (int64_t)n;
    if ((int64_t)n > max_n && max_n > 0) return fail(MS_ERR_CAPACITY, "ms_md_export: %llu infosets, buffers hold %lld", n, (long long)max_n);
    return MS_OK;
}

int ms_md_lookup(ms_mdsolver* s, const uint64_t* d_keys, int64_t n, double* d_regret, double* d_strategy,
                 uint8_t* d_found, void* stream) {
    if (!s || (!d_keys && n > 0) || n < 0) return fail(MS_ERR_ARG, "ms_md_lookup: bad argument");
    if (n == 0) return MS_OK;
    md_lookup_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        s->dev, (const unsigned long long*)d_keys, (long long)n, d_regret, d_strategy, d_found);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_md_mccfr_blocked(ms_mdsolver* s, int32_t player, int64_t first_visit, int64_t n_visits, int32_t pairs_per_visit,
                        uint64_t philox_seed, void* stream) {
    if (!s) return fail(MS_ERR_ARG, "ms_md_mccfr_blocked: NULL handle");
    if (player < 0 || player > 2 || first_visit < 0 || n_visits < 0 || pairs_per_visit < 1 || pairs_per_visit > (1 << 24))
        return fail(MS_ERR_ARG, "ms_md_mccfr_blocked: bad argument");
    if (n_visits == 0) return MS_OK;
    int owner_bits = 0;
    while ((1 << owner_bits) < s->world) owner_bits++;
    if (s->log2cap + owner_bits > 30) return fail(MS_ERR_ARG, "ms_md_mccfr_blocked: global table slots must fit 30 bits");
    if (!s->d_info) {                                      // describe every deal's tree once; claims the table slots
        MS_CUDA(cudaMalloc(&s->d_info, (size_t)s->n_deals * sizeof(MdDealInfo)));
        md_build_kernel<<<grid_for(s->n_deals * 256, 256, 4), 256, 0, (cudaStream_t)stream>>>(s->dev, s->d_info);
        MS_LAUNCH_CHECK();
    }
    const int smem = (int)mdb_smem_bytes();
    MS_CUDA(cudaFuncSetAttribute(md_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = (int)(n_visits < kNumSMs ? n_visits : kNumSMs);
    md_blocked_kernel<<<grid, MDB_THREADS, smem, (cudaStream_t)stream>>>(
        s->dev, s->d_info, player, (unsigned long long)first_visit, (long long)n_visits, pairs_per_visit,
        make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)));
    MS_LAUNCH_CHECK();
    return MS_OK;
}

/* ---- table sharded over the GPUs of one box ---- */

int ms_md_ipc_export(ms_mdsolver* s, void* handles128) {
    if (!s || !handles128) return fail(MS_ERR_ARG, "ms_md_ipc_export: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    MS_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handles128, s->dev.slots));
    MS_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)((char*)handles128 + 64), s->dev.dirty));
    return MS_OK;
}

int ms_md_ipc_attach(ms_mdsolver* s, int32_t rank, int32_t world, const void* handles) {
    if (!s || !handles || world < 1 || world > MD_MAX_PEERS || rank < 0 || rank >= world)
        return fail(MS_ERR_ARG, "ms_md_ipc_attach: bad argument (at most %d ranks)", MD_MAX_PEERS);
    if (s->attached) return fail(MS_ERR_STATE, "the table is already attached to its peers");
    if (s->d_info) return fail(MS_ERR_STATE, "ms_md_ipc_attach: attach before the first traversal (the deal descriptions name table slots)");
    int owner_bits = 0;
    while ((1 << owner_bits) < world) owner_bits++;
    if (s->log2cap + owner_bits > 30) return fail(MS_ERR_ARG, "ms_md_ipc_attach: 2^%d slots x %d ranks exceed the 30-bit global slot id", s->log2cap, world);
    for (int r = 0; r < world; r++) {
        if (r == rank) { s->peer_slots_base[r] = s->dev.slots; s->peer_aux_base[r] = s->dev.dirty; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + 128 * (size_t)r, 64);
        MS_CUDA(cudaIpcOpenMemHandle(&s->peer_slots_base[r], h, cudaIpcMemLazyEnablePeerAccess));
        std::memcpy(&h, (const char*)handles + 128 * (size_t)r + 64, 64);
        MS_CUDA(cudaIpcOpenMemHandle(&s->peer_aux_base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    for (int r = 0; r < world; r++) {
        s->dev.peer_slots[r] = (MdSlot*)s->peer_slots_base[r];
        s->dev.peer_dirty[r] = (unsigned int*)s->peer_aux_base[r];
    }
    s->dev.world = world; s->dev.rank = rank;
    s->rank = rank; s->world = world; s->attached = 1; s->epoch = 0;
    return MS_OK;
}

int ms_md_peer_barrier(ms_mdsolver* s, void* stream) {
    if (!s) return fail(MS_ERR_ARG, "ms_md_peer_barrier: NULL handle");
    if (!s->attached) return MS_OK;                        // one GPU: stream order is the barrier
    MdPeerSync ps{};
    const size_t flags_off = md_aux_flags_off(s->log2cap);
    for (int r = 0; r < s->world; r++) ps.flags[r] = (unsigned long long*)((char*)s->peer_aux_base[r] + flags_off);
    ps.my_flags = s->d_flags; ps.rank = s->rank; ps.world = s->world;
    s->epoch++;
    md_peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ps, s->epoch, s->d_peer_err);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_md_peer_error(ms_mdsolver* s, uint32_t* h_err, void* stream) {
    if (!s || !h_err) return fail(MS_ERR_ARG, "ms_md_peer_error: bad argument");
    MS_CUDA(cudaMemcpyAsync(h_err, s->d_peer_err, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (*h_err) return fail(MS_ERR_STATE, "multi-deal peer barrier: rank %u did not arrive within %llu ms", *h_err - 1u,
                            (unsigned long long)(MD_PEER_TIMEOUT_NS / 1000000ull));
    return MS_OK;
}

int ms_md_info(const ms_mdsolver* s, int64_t* n_deals, int64_t* capacity, int64_t* table_bytes) {
    if (!s) return fail(MS_ERR_ARG, "ms_md_info: NULL handle");
    if (n_deals) *n_deals = s->n_deals;
    if (capacity) *capacity = (int64_t)1 << s->log2cap;
    if (table_bytes) *table_bytes = ((int64_t)1 << s->log2cap) * (int64_t)sizeof(MdSlot);
    return MS_OK;
}

int ms_debug_random_access_peaks(int32_t log2_lines, double h_out[3], void* stream) {
    if (!h_out || log2_lines < 10 || log2_lines > 30) return fail(MS_ERR_ARG, "ms_debug_random_access_peaks: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t lines = (size_t)1 << log2_lines;
    MdSlot* d = nullptr;
    unsigned long long* sink = nullptr;
    MS_CUDA(cudaMalloc(&d, lines * sizeof(MdSlot)));
    MS_CUDA(cudaMalloc(&sink, 8));
    MS_CUDA(cudaMemsetAsync(d, 0, lines * sizeof(MdSlot), st));
    cudaEvent_t e0, e1;
    MS_CUDA(cudaEventCreate(&e0)); MS_CUDA(cudaEventCreate(&e1));
    const int iters = 512, grid = kNumSMs, block = 768;
    const double ops = (double)iters * grid * block;
    for (int which = 0; which < 3; which++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {          // first repetition is the warm-up
            MS_CUDA(cudaEventRecord(e0, st));
            if (which == 0) rnd_dependent_kernel<<<grid, block, 0, st>>>(d, lines - 1, iters, sink);
            else if (which == 1) rnd_independent_kernel<<<grid, block, 0, st>>>(d, lines - 1, iters, sink);
            else rnd_red_kernel<<<grid, block, 0, st>>>(d, lines - 1, iters);
            MS_LAUNCH_CHECK();
            MS_CUDA(cudaEventRecord(e1, st));
            MS_CUDA(cudaEventSynchronize(e1));
            float t = 0.f;
            MS_CUDA(cudaEventElapsedTime(&t, e0, e1));
            if (rep > 0 && t < best) best = t;
        }
        h_out[which] = ops / (best * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    MS_CUDA(cudaFree(d)); MS_CUDA(cudaFree(sink));
    return MS_OK;
}

}  // extern "C"
#endif  // MS_HOST_RULES_ONLY
