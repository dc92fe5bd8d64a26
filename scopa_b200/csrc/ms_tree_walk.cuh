// scopa_b200/csrc/ms_tree_walk.cuh -- the reference's sampled-CFR estimator (mc_cfr.py:37-86) as a depth-first walk
// over an enumerated game tree held in shared memory, plus the small numeric helpers it shares with the other
// solver kernels.  Included by ms_solver.cu (one deal) and ms_multideal.cu (deal-blocked multi-deal solver).
#pragma once
#include <cstdint>

#include "ms_state.cuh"
#include "ms_div.cuh"

namespace ms {

// regret matching: InfoNode.get_strategy (vanilla_cfr.py:23-30) == current_strategy (mc_cfr.py:20-24)
__device__ __forceinline__ void regret_match(const double* reg, int n, double* out) {
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

// np.random.choice(legal, p=sigma): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')
__device__ __forceinline__ int sample_action(const double* sg, int n, double u) {
    double cdf[4];
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) { if (i < n) acc = __dadd_rn(acc, sg[i]); cdf[i] = acc; }
    const double last = acc;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < n && ms_ddiv_or_zero(cdf[i], last) <= u) idx++;
    return idx < n ? idx : n - 1;
}

// normalised cdf of one strategy, numpy's rule: cdf = cumsum(p); cdf /= cdf[-1]
__device__ __forceinline__ void strategy_cdf(const double* sg, int n, double* cdf) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) { if (i < n) acc = __dadd_rn(acc, sg[i]); cdf[i] = acc; }
    const double last = acc;
#pragma unroll
    for (int i = 0; i < 4; i++) cdf[i] = (i < n) ? ms_ddiv_or_zero(cdf[i], last) : 2.0;
}

struct MccfrShared {
    const unsigned long long* hk; const int16_t* hs; int hcap;
    const double* sig;        // frozen strategies (batch mode), [S][4]
    const double* cdf;        // their normalised cdfs (batch mode), [S][4]
    double* dreg;             // per-CTA private regret deltas (batch mode), [S][4]
    uint32_t* dcnt;           // per-CTA update counts (batch mode), [S]
    uint8_t* touched;         // per-CTA touched flags, [S]
    double* reg; double* str; // in-place mode: the table itself
};

constexpr int TREE_THREADS = 1024;
constexpr uint32_t TREE_TERMINAL = 0x7FFu;

struct TreeFrames {       // SoA in shared memory: [frame][thread]
    double* ro; double* sp; uint32_t* meta; uint32_t* cfv; uint16_t* cb;
};

__host__ __device__ inline size_t mccfr_tree_smem(int S, int n_nodes, int nframes, int threads, int ncopy) {
    return sizeof(double) * (7 + 4 * (size_t)ncopy) * (size_t)S + 4 * (size_t)n_nodes + 4 * (size_t)S + (size_t)S +
           (size_t)threads * nframes * (8 + 8 + 4 + 4 + 2) + 64;
}

// INPLACE = the reference's own schedule (one traversal at a time, every update visible to the next node visit:
// sh.reg / sh.str are the table itself); otherwise frozen-sigma batch semantics (sh.sig / sh.cdf / sh.dreg / sh.dcnt).
template <bool INPLACE>
__device__ void mccfr_tree_traverse(const uint32_t* __restrict__ tree, const MccfrShared& sh, int tp, unsigned long long trav,
                                    uint2 pkey, const TreeFrames& f, int fstride, unsigned long long& n_upd,
                                    unsigned long long& n_vis, unsigned long long& n_step) {
    uint32_t node = 0u;
    double ro = 1.0, sp = 1.0;
    int fi = -1;
    uint4 xblk = make_uint4(0u, 0u, 0u, 0u);
    uint32_t xblk_id = 0xFFFFFFFFu, call = 0u;
    int ret_x2 = 0;
    bool returning = false;
    const uint32_t tag = MS_TAG_MCCF + (uint32_t)tp;
    while (true) {
        if (!returning) {
            const uint32_t rec = tree[node];
            const uint32_t my_call = call++;
            n_vis++;
            const uint32_t slot = (rec >> 12) & 0x7FFu;
            if (slot == TREE_TERMINAL) {
                const int r = (int)(rec & 0xFFFu) - 2048;
                ret_x2 = (tp == 0) ? r : -r;
                returning = true;
                continue;
            }
            const uint32_t nl = (rec >> 23) & 0x7u, cb = rec & 0xFFFu;
            const int p = (int)((rec >> 26) & 1u);
            sh.touched[slot] = 1;     // node created on first touch, for both players (mc_cfr.py:52)
            if (nl == 1u) {           // forced move: sigma = [1.0], no random word (see mccfr_traverse)
                if (p != tp) { node = cb; n_step++; continue; }
                const uint32_t r1 = tree[cb];
                uint32_t leaf = r1, slot2 = TREE_TERMINAL;
                int below = 1;
                bool forced = ((r1 >> 12) & 0x7FFu) == TREE_TERMINAL;
                if (!forced && ((r1 >> 23) & 0x7u) == 1u) {          // the opponent's reply is forced as well
                    slot2 = (r1 >> 12) & 0x7FFu;
                    leaf = tree[r1 & 0xFFFu];
                    below = 2;
                    forced = ((leaf >> 12) & 0x7FFu) == TREE_TERMINAL;
                }
                if (forced) {         // both recursive calls of the reference walk this line: played once, accounted twice
                    if (slot2 != TREE_TERMINAL) sh.touched[slot2] = 1;
                    const int r = (int)(leaf & 0xFFFu) - 2048;
                    ret_x2 = (tp == 0) ? r : -r;
                    // regret delta = w * 0 exactly; strategy_sum += 1.0 * sigma = 1.0
                    if (INPLACE) sh.str[4 * slot] = __dadd_rn(sh.str[4 * slot], 1.0);
                    else atomicAdd(&sh.dcnt[slot], 1u);
                    n_upd++;
                    n_vis += 2 * below; call += 2u * (uint32_t)below; n_step += below;
                    returning = true;
                    continue;
                }
            }
            int ai = 0;
            double sga = 1.0;
            if (nl > 1u) {
                if ((my_call >> 1) != xblk_id) {
                    xblk_id = my_call >> 1;
                    xblk = philox4x32_10(make_uint4((uint32_t)trav, (uint32_t)(trav >> 32), xblk_id, tag), pkey);
                }
                const double u = (my_call & 1u) ? u53(xblk.z, xblk.w) : u53(xblk.x, xblk.y);
                if (INPLACE) {
                    double sg[4];
                    regret_match(sh.reg + 4 * slot, (int)nl, sg);
                    ai = sample_action(sg, (int)nl, u);
                    sga = ai == 0 ? sg[0] : (ai == 1 ? sg[1] : (ai == 2 ? sg[2] : sg[3]));
                } else {
                    // searchsorted(cdf, u, 'right') reading only the first nl-1 entries: cdf[nl-1] is exactly 1.0 > u
                    // (a lane's cdf row is the widest shared-memory read of a visit)
                    const double* cd = sh.cdf + 3 * slot;
#pragma unroll
                    for (int i = 0; i < 3; i++)
                        if ((uint32_t)i + 1u < nl) ai += (cd[i] <= u) ? 1 : 0;
                    sga = sh.sig[4 * slot + ai];
                }
            }
            node = cb + (uint32_t)ai;
            n_step++;
            if (p != tp) {            // opponent: reach *= sigma[a]; tail call (mc_cfr.py:63-65)
                ro = __dmul_rn(ro, sga);
                continue;
            }
            fi++;                     // traverser: push a frame, descend into the sampled action first (:58-67)
            const int o = fi * fstride;
            f.ro[o] = ro; f.sp[o] = sp;
            f.meta[o] = slot | (nl << 11);          // | cursor << 14 | util byte << 17
            f.cfv[o] = 0u;
            f.cb[o] = (uint16_t)cb;
            sp = __dmul_rn(sp, sga);
            continue;
        }
        // ---- a child returned ret_x2 to the top frame
        if (fi < 0) break;
        const int o = fi * fstride;
        uint32_t meta = f.meta[o];
        const int slot = (int)(meta & 0x7FFu);
        const int nl = (int)((meta >> 11) & 0x7u);
        int cur = (int)((meta >> 14) & 0x7u);
        uint32_t cfvb = f.cfv[o];
        if (cur == 0) meta = (meta & 0x1FFFFu) | (((uint32_t)ret_x2 & 0xFFu) << 17);   // util of the sampled action
        else cfvb |= ((uint32_t)ret_x2 & 0xFFu) << (8 * (cur - 1));
        cur++;
        if (cur <= nl) {              // evaluate action i = cur-1 with a fresh sampled continuation (:71-78)
            const int i = cur - 1;
            f.meta[o] = (meta & ~(0x7u << 14)) | ((uint32_t)cur << 14);
            f.cfv[o] = cfvb;
            ro = f.ro[o];
            double sgi;
            if (INPLACE) {            // unchanged since entry: an infoset cannot recur below itself
                double sg[4];
                regret_match(sh.reg + 4 * slot, nl, sg);
                sgi = i == 0 ? sg[0] : (i == 1 ? sg[1] : (i == 2 ? sg[2] : sg[3]));
            } else sgi = sh.sig[4 * slot + i];
            sp = __dmul_rn(f.sp[o], sgi);
            node = (uint32_t)f.cb[o] + (uint32_t)i;
            n_step++;
            returning = false;
            continue;
        }
        // ---- all actions evaluated: regret / strategy deltas (:79-84)
        if (INPLACE || nl > 1) {      // batch mode skips |A| = 1: cfv - v == 0 exactly
            double cfv[4], sg[4];
            double v = 0.0;
            if (INPLACE) regret_match(sh.reg + 4 * slot, nl, sg);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (!INPLACE) sg[i] = (i < nl) ? sh.sig[4 * slot + i] : 0.0;
                cfv[i] = 0.5 * (double)(int)(int8_t)((cfvb >> (8 * i)) & 0xFFu);
                if (i < nl) v = __dadd_rn(v, __dmul_rn(sg[i], cfv[i]));
            }
            const double fro = f.ro[o], fsp = f.sp[o];
            const double w = fsp > 0.0 ? __ddiv_rn(fro, fsp) : 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < nl) {
                    if (INPLACE) {
                        sh.reg[4 * slot + i] = __dadd_rn(sh.reg[4 * slot + i], __dmul_rn(w, __dadd_rn(cfv[i], -v)));
                        sh.str[4 * slot + i] = __dadd_rn(sh.str[4 * slot + i], __dmul_rn(1.0, sg[i]));   // reach_probs[tp] is always 1.0
                    } else atomicAdd(&sh.dreg[4 * slot + i], __dmul_rn(w, __dadd_rn(cfv[i], -v)));
                }
        }
        if (!INPLACE) atomicAdd(&sh.dcnt[slot], 1u);   // strategy delta = count * sigma (sigma is frozen for the batch)
        n_upd++;
        ret_x2 = (int)(int8_t)((meta >> 17) & 0xFFu);
        fi--;
        returning = true;
    }
}

}  // namespace ms
