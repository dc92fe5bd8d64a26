// scopa_b200/csrc/ms_state.cuh -- packed Miniscopa game state and the game rules as device code.
//
// Replaces the list-of-Card-objects state of the reference
//   /root/reference/src/envs/mini_scopa_game.py:36-114 (Player, MiniScopaGame)
//   /root/reference/src/envs/mini_scopa_game.py:140-167 (MiniScopaEnv.step)
// with one 16-byte word group that a thread holds in four registers and a warp loads/stores as
// coalesced 128-bit accesses.
//
// Layout (MsState = uint4, little end first):
//   x : hand[0] (bits 0-15, bit c = card id c) | hand[1] (bits 16-31)
//   y : table, ORDERED: nibble i = card id of the i-th oldest table card (unused nibbles are 0).
//       The reference's capture tie-breaks and its infoset strings depend on table ORDER
//       (SURVEY.md H1/H2), so a set bitmask is not enough.
//   z : captures[0] | captures[1] << 16        (bit = card id; capture ORDER is presentation only
//                                               and is re-derived on the host)
//   w : bits 0-3 table_len | 4-7 scopas[0] | 8-11 scopas[1] | 12-16 step_count | 17 current player
//       | 18 terminal | 19-23 max_steps (8 from reset, 16 on cloned states: openspiel_mini_scopa.py:108)
//
// Card id = suit_idx*4 + card_idx == the reference's action id (mini_scopa_game.py:149-153).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ms {

typedef uint4 MsState;

// rank of card id c in nibble c: ids 0..15 -> 2,5,8,10, 2,5,7,9, 3,6,8,9, 3,6,7,10
// (mini_scopa_game.py:18-23)
#define MS_RANK_LUT 0xA76398639752A852ull

__device__ __forceinline__ uint32_t card_rank(uint32_t c) {
    return (uint32_t)(MS_RANK_LUT >> (4u * c)) & 0xFu;
}

__device__ __forceinline__ uint32_t st_hand(const MsState& s, int p) { return p ? (s.x >> 16) : (s.x & 0xFFFFu); }
__device__ __forceinline__ uint32_t st_caps(const MsState& s, int p) { return p ? (s.z >> 16) : (s.z & 0xFFFFu); }
__device__ __forceinline__ uint32_t st_table_len(const MsState& s) { return s.w & 0xFu; }
__device__ __forceinline__ uint32_t st_scopas(const MsState& s, int p) { return (s.w >> (4 + 4 * p)) & 0xFu; }
__device__ __forceinline__ uint32_t st_step_count(const MsState& s) { return (s.w >> 12) & 0x1Fu; }
__device__ __forceinline__ int st_cur(const MsState& s) { return (int)((s.w >> 17) & 1u); }
__device__ __forceinline__ bool st_terminal(const MsState& s) { return (s.w >> 18) & 1u; }
__device__ __forceinline__ uint32_t st_max_steps(const MsState& s) { return (s.w >> 19) & 0x1Fu; }

__device__ __forceinline__ MsState st_make(uint32_t hand0, uint32_t hand1, uint32_t max_steps) {
    MsState s;
    s.x = (hand0 & 0xFFFFu) | (hand1 << 16);
    s.y = 0u; s.z = 0u;
    s.w = (max_steps & 0x1Fu) << 19;
    return s;
}

// Each rank value appears on exactly two cards (mini_scopa_game.py:18-23), so the only table card
// that can have the rank of a played card c is its twin: nibble c of this LUT.
#define MS_TWIN_LUT 0x369872DCBE10FA54ull
__device__ __forceinline__ uint32_t card_twin(uint32_t c) {
    return (uint32_t)(MS_TWIN_LUT >> (4u * c)) & 0xFu;
}

// set of card ids on the table (from the ordered nibble list)
__device__ __forceinline__ uint32_t table_set(uint32_t order, uint32_t len) {
    uint32_t m = 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if ((uint32_t)i >= len) break;
        m |= 1u << ((order >> (4 * i)) & 0xFu);
    }
    return m;
}

// position (0..7) of the lowest nibble of `order` equal to `c` (caller guarantees there is one)
__device__ __forceinline__ uint32_t nibble_pos(uint32_t order, uint32_t c) {
    const uint32_t x = order ^ (c * 0x11111111u);
    const uint32_t z = (x - 0x11111111u) & ~x & 0x88888888u;   // lowest flagged nibble is an exact zero nibble
    return (uint32_t)(__ffs((int)z) - 1) >> 2;
}

// drop nibble p from an ordered nibble list (higher nibbles move down)
__device__ __forceinline__ uint32_t nibble_remove(uint32_t order, uint32_t p) {
    const uint32_t low = (1u << (4u * p)) - 1u;
    return (order & low) | ((order >> 4) & ~low);
}

// Capture resolution -- MiniScopaGame.card_in_table (mini_scopa_game.py:66-91).
// Returns the bitmask over TABLE POSITIONS (bit i = i-th oldest card) of the captured cards, 0 if
// the card is placed.  `tset` = set of card ids on the table.
//   * equal rank present  -> exactly the first such card in table order (:72-74); only the played
//     card's twin can match, so this is one bit test plus a nibble search;
//   * otherwise the reference's 1-D DP returns the first-found subset, which is the subset with the
//     minimum numeric position mask (proof in DESIGN.md "Capture rule"): computed here with prefix
//     reachable-sum bitsets R[i] (bit s set <=> some subset of the i oldest cards sums to s) and a
//     top-down walk that takes card i only when the remaining target is NOT reachable without it.
//     Loops are bounded by the table length (early exit), which is warp-coherent in tree traversals.
__device__ __forceinline__ uint32_t capture_mask(uint32_t order, uint32_t len, uint32_t card, uint32_t tset) {
    const uint32_t twin = card_twin(card);
    const bool twin_hit = (tset >> twin) & 1u;
    const bool need_dp = !twin_hit && len >= 2u;   // a sum needs two cards (a single equal card is the twin)
    // Loop bound shared by the converged lanes (the longest table among those that need the subset
    // search): the loops below then run without divergence, lanes with shorter tables are predicated.
    const uint32_t bound = __reduce_max_sync(__activemask(), need_dp ? len : 0u);
    uint32_t m = 0u;
    if (bound) {
        const uint32_t rank = card_rank(card);
        uint32_t R[9];
        R[0] = 1u;
        uint32_t Rall = 1u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if ((uint32_t)i >= bound) break;
            const uint32_t ri = card_rank((order >> (4 * i)) & 0xFu);
            if ((uint32_t)i < len) Rall = (Rall | (Rall << ri)) & 0x7FFu;
            R[i + 1] = Rall;
        }
        uint32_t t = (need_dp && ((Rall >> rank) & 1u)) ? rank : 0u;
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            if ((uint32_t)i >= bound) continue;
            if ((uint32_t)i < len && t > 0u && !((R[i] >> t) & 1u)) {
                m |= 1u << i;
                t -= card_rank((order >> (4 * i)) & 0xFu);
            }
        }
    }
    return twin_hit ? (1u << nibble_pos(order, twin)) : m;
}

// MiniScopaEnv.step (mini_scopa_game.py:140-167) + MiniScopaGame.play_card (:93-104).
// Illegal action (card not in the mover's hand, or id outside 0..15) = silent pass that still
// advances step_count and the turn (:155-167).  A step on a terminal state is a no-op (:141-143).
// Returns the table-position capture mask (0 when the card was placed or the move was a pass).
// `tset` must be the set of card ids currently on the table (table_set(), or dealt & ~hands & ~caps).
__device__ __forceinline__ uint32_t step(MsState& s, uint32_t action, uint32_t tset) {
    if (st_terminal(s)) return 0u;
    const int p = st_cur(s);
    const uint32_t hand = st_hand(s, p);
    uint32_t capm = 0u;
    if (action < 16u && ((hand >> action) & 1u)) {
        const uint32_t len = st_table_len(s);
        uint32_t order = s.y;
        capm = capture_mask(order, len, action, tset);
        if (capm) {
            uint32_t taken = 1u << action, k = len, m = capm;
            while (m) {                                    // captured positions, highest first
                const uint32_t i = 31u - (uint32_t)__clz((int)m);
                m ^= 1u << i;
                taken |= 1u << ((order >> (4u * i)) & 0xFu);
                order = nibble_remove(order, i);
                k--;
            }
            s.y = order;
            s.z |= taken << (16 * p);
            s.w = (s.w & ~0xFu) | k;
            if (k == 0u) s.w += 1u << (4 + 4 * p);        // scopa (:101-102)
        } else {
            s.y = order | (action << (4u * len));         // placed at the end of the table (:104)
            s.w += 1u;                                    // table_len++ (<= 8 by construction)
        }
        s.x &= ~((1u << action) << (16 * p));             // hand.remove(card)
    }
    s.w += 1u << 12;                                      // step_count++
    const bool term = (s.x == 0u) || (st_step_count(s) >= st_max_steps(s));
    s.w ^= 1u << 17;                                      // next player
    if (term) s.w |= 1u << 18;
    return capm;
}

__device__ __forceinline__ uint32_t step(MsState& s, uint32_t action) {
    return step(s, action, table_set(s.y, st_table_len(s)));
}

// every card of the deal that is in a hand, on the table or captured: constant along a game, so
// table set = dealt & ~hands & ~captures without touching the ordered list
__device__ __forceinline__ uint32_t dealt_set(const MsState& s) {
    return ((s.x | (s.x >> 16) | s.z | (s.z >> 16)) & 0xFFFFu) | table_set(s.y, st_table_len(s));
}
__device__ __forceinline__ uint32_t table_set_from_dealt(const MsState& s, uint32_t dealt) {
    return dealt & ~((s.x | (s.x >> 16) | s.z | (s.z >> 16)) & 0xFFFFu);
}

// evaluate_game (mini_scopa_game.py:106-114): s_i = |captures_i| + 2*scopas_i, r_i = s_i - mean.
// For two players r_0 = (s_0 - s_1)/2 = -r_1 (and 0 when both are 0).
__device__ __forceinline__ float reward0(const MsState& s) {
    int s0 = __popc(st_caps(s, 0)) + 2 * (int)st_scopas(s, 0);
    int s1 = __popc(st_caps(s, 1)) + 2 * (int)st_scopas(s, 1);
    return 0.5f * (float)(s0 - s1);
}
// 2*r_0 as an exact small integer (range [-9, 9] observed; used where a byte is stored)
__device__ __forceinline__ int reward0_x2(const MsState& s) {
    int s0 = __popc(st_caps(s, 0)) + 2 * (int)st_scopas(s, 0);
    int s1 = __popc(st_caps(s, 1)) + 2 * (int)st_scopas(s, 1);
    return s0 - s1;
}

// MiniScopaState.legal_actions (openspiel_mini_scopa.py:22-47): action ids of the player's cards in
// HAND (= deal) order; [0] if the hand is empty and the state is not terminal; [] if terminal.
// hand_order: nibbles 0-3 = player 0's dealt cards in order, nibbles 4-7 = player 1's.
// Returns the count and packs the ids into nibbles of `list` (nibble k = k-th legal action).
__device__ __forceinline__ uint32_t legal_list(const MsState& s, uint32_t hand_order, int player, uint32_t& list) {
    list = 0u;
    if (st_terminal(s)) return 0u;
    uint32_t hand = st_hand(s, player);
    uint32_t ord = (hand_order >> (16 * player)) & 0xFFFFu;
    uint32_t n = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t c = (ord >> (4 * i)) & 0xFu;
        if ((hand >> c) & 1u) { list |= c << (4u * n); n++; hand &= ~(1u << c); }
    }
    if (n == 0u) return 1u;     // fallback [0] (:47); list == 0 already encodes action 0
    return n;
}

// 64-bit infoset key: what information_state_string (openspiel_mini_scopa.py:86-95) shows, packed:
// player | hand mask | table_len | ordered table.  The hand's ORDER in the string is the deal order
// restricted to the mask, so within one deal the key and the string are in bijection.
__device__ __forceinline__ uint64_t infoset_key(const MsState& s, int player) {
    return ((uint64_t)(player & 1) << 52) | ((uint64_t)st_hand(s, player) << 36) |
           ((uint64_t)st_table_len(s) << 32) | (uint64_t)s.y;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: the same stream specification as the oracle
// (oracle/ms_oracle.c implements it independently).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
#define MS_TAG_ROLL 0x4C4C4F52u
#define MS_TAG_MCCF 0x4643434Du
#define MS_TAG_SDCF 0x46434453u

// 53-bit uniform in [0,1) from two 32-bit words, numpy's random_sample formula
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

}  // namespace ms
