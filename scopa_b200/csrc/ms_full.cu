// scopa_b200/csrc/ms_full.cu -- the 40-card Scopa environment (two players) as device code.  sm_100a only.
//
// Replaces (paths relative to /root/reference/):
//   FullDeck / FullScopaGame / FullScopaEnv     src/envs/full_scopa_game.py:21-342
//   FullScopaState.legal_actions                src/envs/openspiel_full_scopa.py:22-41
// SURVEY.md section 8(f) row 4 ("... then full 40-card Scopa").  Same rule family as Miniscopa (ms_state.cuh): a
// played card takes the first table card of equal rank, else the first subset of table cards -- in the
// reference's ascending enumeration of position masks, i.e. the subset with the smallest mask -- whose ranks sum
// to its rank (find_capture_combinations :101-128, play_card takes combinations[0] :137-141).  What is new:
// ranks 1..10 in four suits, four cards on the table at the start, six hands of three cards dealt from the
// shuffled deck as the game goes (:88-99), the last capturer sweeping the table at the end, and the traditional
// scoring (carte, denari, sette bello, primiera, scope; evaluate_game :174-226).
//
// Card id = suit_idx * 10 + (rank - 1) = the reference's action id (:262-266); suits denari, coppe, spade, bastoni.
//
// Packed state, 32 bytes (two 128-bit accesses per lane), words w0..w7:
//   w0 w1 : ordered table, entries 0..9, six bits each (bits 0-59 of the 64-bit pair), oldest first
//   w2 w7 : ordered table, entries 10..15 (bits 0-35 of the pair); 200 k random games never hold more than 11
//   w3 w4 : captures[0] / captures[1], cards 0..31 (bit = card id)
//   w5    : bits 0-7 captures[0] cards 32..39 | 8-15 captures[1] cards 32..39 | 16-21 cards still in hand (bit
//           16 + 3 p + i = i-th card dealt to player p in this round) | 22-26 table_len | 27-28 last capturer + 1
//           (0 = nobody yet) | 29 current player | 30 terminal | 31 evaluate_game ran twice (see fs_step)
//   w6    : bits 0-5 scopas[0] | 6-11 scopas[1] | 12-14 round_number | 15-22 step_count | 23-30 final score
//           difference s0 - s1 as a signed byte (rewards are +-(s0 - s1) / 2)
// Beside each game: the shuffled deck (ms_full_deck, 4 x 64 bits, ten 6-bit ids per word).  The table is
// deck[0..3]; round r deals deck[4 + 6 r + 3 p + i] to player p, so hands are presence bits, not card lists.
#include <mutex>

#include "ms_common.cuh"
#include "ms_state.cuh"

namespace ms {

struct FsState { uint32_t w[8]; };
struct FsDeck { unsigned long long w[4]; };

constexpr uint32_t MS_TAG_FULL = 0x4C4C5546u;   // "FULL"
constexpr int FS_MAX_TABLE = 16;
constexpr uint32_t FS_MAX_STEPS = 200u;         // FullScopaEnv.max_steps (:239)

__device__ __forceinline__ FsState fs_load(const uint4* p, long long g) {
    const uint4 a = p[2 * g], b = p[2 * g + 1];
    FsState s;
    s.w[0] = a.x; s.w[1] = a.y; s.w[2] = a.z; s.w[3] = a.w; s.w[4] = b.x; s.w[5] = b.y; s.w[6] = b.z; s.w[7] = b.w;
    return s;
}
__device__ __forceinline__ void fs_store(uint4* p, long long g, const FsState& s) {
    p[2 * g] = make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]);
    p[2 * g + 1] = make_uint4(s.w[4], s.w[5], s.w[6], s.w[7]);
}
__device__ __forceinline__ FsDeck fs_load_deck(const ulonglong4* p, long long g) {
    const ulonglong4 d = p[g];
    FsDeck k; k.w[0] = d.x; k.w[1] = d.y; k.w[2] = d.z; k.w[3] = d.w;
    return k;
}
__device__ __forceinline__ uint32_t fs_deck_card(const FsDeck& d, int pos) {
    const int q = pos / 10;
    const unsigned long long x = q == 0 ? d.w[0] : (q == 1 ? d.w[1] : (q == 2 ? d.w[2] : d.w[3]));
    return (uint32_t)(x >> (6 * (pos % 10))) & 0x3Fu;
}

__device__ __forceinline__ uint32_t fs_rank(uint32_t c) { return c - 10u * ((c * 205u) >> 11) + 1u; }   // c % 10 + 1
__device__ __forceinline__ uint32_t fs_table_len(const FsState& s) { return (s.w[5] >> 22) & 0x1Fu; }
__device__ __forceinline__ int fs_cur(const FsState& s) { return (int)((s.w[5] >> 29) & 1u); }
__device__ __forceinline__ bool fs_terminal(const FsState& s) { return (s.w[5] >> 30) & 1u; }
__device__ __forceinline__ uint32_t fs_round(const FsState& s) { return (s.w[6] >> 12) & 0x7u; }
__device__ __forceinline__ uint32_t fs_step_count(const FsState& s) { return (s.w[6] >> 15) & 0xFFu; }
__device__ __forceinline__ uint32_t fs_hand_bits(const FsState& s, int p) { return (s.w[5] >> (16 + 3 * p)) & 0x7u; }
__device__ __forceinline__ unsigned long long fs_caps(const FsState& s, int p) {
    return (unsigned long long)s.w[3 + p] | ((unsigned long long)((s.w[5] >> (8 * p)) & 0xFFu) << 32);
}
__device__ __forceinline__ void fs_add_caps(FsState& s, int p, unsigned long long m) {
    s.w[3 + p] |= (uint32_t)m;
    s.w[5] |= ((uint32_t)(m >> 32) & 0xFFu) << (8 * p);
}
__device__ __forceinline__ unsigned long long fs_tlo(const FsState& s) { return (unsigned long long)s.w[0] | ((unsigned long long)s.w[1] << 32); }
__device__ __forceinline__ unsigned long long fs_thi(const FsState& s) { return (unsigned long long)s.w[2] | ((unsigned long long)s.w[7] << 32); }
__device__ __forceinline__ void fs_set_table(FsState& s, unsigned long long lo, unsigned long long hi) {
    s.w[0] = (uint32_t)lo; s.w[1] = (uint32_t)(lo >> 32); s.w[2] = (uint32_t)hi; s.w[7] = (uint32_t)(hi >> 32);
}
__device__ __forceinline__ uint32_t fs_table_card(unsigned long long lo, unsigned long long hi, int i) {
    return (uint32_t)((i < 10 ? lo >> (6 * i) : hi >> (6 * (i - 10))) & 0x3Full);
}

__device__ __forceinline__ FsState fs_initial(const FsDeck& d) {
    FsState s;
#pragma unroll
    for (int i = 0; i < 8; i++) s.w[i] = 0u;
    unsigned long long lo = 0ull;
#pragma unroll
    for (int i = 0; i < 4; i++) lo |= (unsigned long long)fs_deck_card(d, i) << (6 * i);   // table = deck.deal(4) (:79)
    fs_set_table(s, lo, 0ull);
    s.w[5] = (0x3Fu << 16) | (4u << 22);                                                  // three cards each (:82-84)
    return s;
}

// the six cards dealt in the state's round, 6 bits each: card i of player p at bits 6 * (3 p + i).  Fused kernels
// extract them once per round instead of going through the deck at every ply.
__device__ __forceinline__ unsigned long long fs_round_cards(const FsState& s, const FsDeck& d) {
    const int base = 4 + 6 * (int)fs_round(s);
    unsigned long long hc = 0ull;
#pragma unroll
    for (int k = 0; k < 6; k++) hc |= (unsigned long long)fs_deck_card(d, base + k) << (6 * k);
    return hc;
}
__device__ __forceinline__ uint32_t fs_hand_card(unsigned long long hc, int p, int i) {
    return (uint32_t)(hc >> (6 * (3 * p + i))) & 0x3Fu;
}

// FullScopaState.legal_actions (openspiel_full_scopa.py:22-41): the hand in list (= deal) order; [0] when the hand
// is empty and the game is not over; [] when it is.  Packs the ids into bytes of `list`, returns the count.
__device__ __forceinline__ uint32_t fs_legal_list(const FsState& s, unsigned long long hc, int p, uint32_t& list) {
    list = 0u;
    if (fs_terminal(s)) return 0u;
    const uint32_t bits = fs_hand_bits(s, p);
    uint32_t n = 0u;
#pragma unroll
    for (int i = 0; i < 3; i++)
        if ((bits >> i) & 1u) { list |= fs_hand_card(hc, p, i) << (8u * n); n++; }
    return n ? n : 1u;
}

// find_capture_combinations(card)[0] as a mask over table positions (0 = the card is placed).
// `scr` = this thread's column of a shared-memory scratch array, FS_THREADS apart, FS_MAX_TABLE entries:
// entry i = (reachable-sum bitset of the i oldest cards) | (rank of card i) << 11.  The first version kept the
// prefix bitsets in a local-memory array and fetched every table card with a variable 64-bit shift; its capture
// (profiles/README.md section 7) put a quarter of the kernel's instructions on those two lines, at 12 of 32 lanes
// (the loops ran to each lane's own table length).  Here the table slides through a register six bits at a time,
// the loops run to a bound shared by the converged lanes, and the backtrack is one shared-memory load per card.
constexpr int FS_THREADS = 256;
__device__ __forceinline__ uint32_t fs_capture_mask(unsigned long long lo, unsigned long long hi, uint32_t len, uint32_t card,
                                                    uint16_t* scr) {
    const uint32_t target = fs_rank(card);
    const uint32_t bound = __reduce_max_sync(__activemask(), len);
    uint32_t r = 1u, first_equal = 0xFFu;       // r: bit v set <=> some subset of the cards seen so far sums to v
    unsigned long long cur = lo;
    for (uint32_t i = 0; i < bound; i++) {
        if (i == 10u) cur = hi;
        const uint32_t ri = fs_rank((uint32_t)cur & 0x3Fu);
        cur >>= 6;
        scr[i * FS_THREADS] = (uint16_t)(r | (ri << 11));
        if (i < len) {
            if (ri == target && first_equal == 0xFFu) first_equal = i;
            r = (r | (r << ri)) & 0x7FFu;
        }
    }
    if (first_equal != 0xFFu) return 1u << first_equal;              // exact rank match has priority (:107-110)
    uint32_t t = ((r >> target) & 1u) ? target : 0u;
    uint32_t m = 0u;
    if (__any_sync(__activemask(), t > 0u)) {
        for (int i = (int)bound - 1; i >= 0; i--) {                  // smallest mask: take card i only when it is needed
            const uint32_t e = scr[i * FS_THREADS];
            if ((uint32_t)i < len && t > 0u && !((e >> t) & 1u)) {
                m |= 1u << i;
                t -= e >> 11;
            }
        }
    }
    return m;
}

__device__ __forceinline__ int fs_primiera(unsigned long long caps) {    // calculate_primiera_score (:160-172)
    int total = 0;
#pragma unroll
    for (int su = 0; su < 4; su++) {
        const uint32_t f = (uint32_t)(caps >> (10 * su)) & 0x3FFu;         // bit r-1 = rank r of this suit
        int v;
        if (f & (1u << 6)) v = 21;            // 7
        else if (f & (1u << 5)) v = 18;       // 6
        else if (f & 1u) v = 16;              // ace
        else if (f & (1u << 4)) v = 15;
        else if (f & (1u << 3)) v = 14;
        else if (f & (1u << 2)) v = 13;
        else if (f & (1u << 1)) v = 12;
        else if (f) v = 10;                   // 8, 9, 10
        else return 0;                        // a suit is missing: no primiera
        total += v;
    }
    return total;
}

// evaluate_game (:174-226).  `twice`: the reference calls it a second time when the last card falls on step 200
// (:278-290); the sweep then appends the table to the last capturer's list again, which counts those cards twice.
// `detail` (may be NULL): cards, denari, primiera sum and score of each player, [c0 c1 d0 d1 p0 p1 s0 s1].
__device__ __forceinline__ void fs_evaluate(FsState& s, bool twice, int* detail = nullptr) {
    const uint32_t len = fs_table_len(s), last = (s.w[5] >> 27) & 3u;
    const unsigned long long lo = fs_tlo(s), hi = fs_thi(s);
    int extra_cards = 0, extra_denari = 0;
    if (len > 0u && last != 0u) {             // the table is NOT cleared (:187-188)
        unsigned long long m = 0ull;
        int den = 0;
        for (uint32_t i = 0; i < len; i++) {
            const uint32_t c = fs_table_card(lo, hi, (int)i);
            m |= 1ull << c;
            den += c < 10u;
        }
        fs_add_caps(s, (int)last - 1, m);
        if (twice) { extra_cards = (int)len; extra_denari = den; }
    }
    const unsigned long long c0 = fs_caps(s, 0), c1 = fs_caps(s, 1);
    int n0 = __popcll(c0), n1 = __popcll(c1), d0 = __popcll(c0 & 0x3FFull), d1 = __popcll(c1 & 0x3FFull);
    if (last == 1u) { n0 += extra_cards; d0 += extra_denari; }
    if (last == 2u) { n1 += extra_cards; d1 += extra_denari; }
    int s0 = (int)(s.w[6] & 0x3Fu), s1 = (int)((s.w[6] >> 6) & 0x3Fu);     // scope
    if (n0 != n1) { if (n0 > n1) s0++; else s1++; }                        // carte
    if (d0 != d1) { if (d0 > d1) s0++; else s1++; }                        // denari
    if ((c0 >> 6) & 1ull) s0++; else if ((c1 >> 6) & 1ull) s1++;           // sette bello
    const int p0 = fs_primiera(c0), p1 = fs_primiera(c1);
    if (p0 != p1) { if (p0 > p1) s0++; else s1++; }                        // primiera (both 0: nobody)
    s.w[6] = (s.w[6] & ~(0xFFu << 23)) | (((uint32_t)(s0 - s1) & 0xFFu) << 23);
    s.w[5] |= 1u << 30;
    if (twice) s.w[5] |= 1u << 31;
    if (detail) {
        detail[0] = n0; detail[1] = n1; detail[2] = d0; detail[3] = d1;
        detail[4] = p0; detail[5] = p1; detail[6] = s0; detail[7] = s1;
    }
}
__device__ __forceinline__ int fs_score_diff(const FsState& s) { return (int)(int8_t)((s.w[6] >> 23) & 0xFFu); }

// FullScopaEnv.step (:252-296) + FullScopaGame.play_card (:130-158).  A card the mover does not hold (or an id
// outside 0..39, which raises IndexError in the reference) is a silent pass that still advances step_count and the
// turn; a step on a finished game is a no-op (:253-255).  Returns false if the table outgrew the packed state.
// `hc` = fs_round_cards() of the state's round (stale once the step has dealt a new round).
__device__ __forceinline__ bool fs_step(FsState& s, unsigned long long hc, uint32_t action, uint16_t* scr) {
    if (fs_terminal(s)) return true;
    const int p = fs_cur(s);
    const uint32_t bits = fs_hand_bits(s, p);
    int hp = -1;
#pragma unroll
    for (int i = 2; i >= 0; i--)
        if (((bits >> i) & 1u) && fs_hand_card(hc, p, i) == action) hp = i;
    bool ok = true;
    if (hp >= 0) {
        unsigned long long lo = fs_tlo(s), hi = fs_thi(s);
        uint32_t len = fs_table_len(s);
        uint32_t m = fs_capture_mask(lo, hi, len, action, scr);
        if (m) {
            unsigned long long taken = 1ull << action;
            while (m) {                                   // captured positions, highest first
                const int i = 31 - __clz((int)m);
                m ^= 1u << i;
                taken |= 1ull << fs_table_card(lo, hi, i);
                if (i < 10) {
                    const unsigned long long low = (1ull << (6 * i)) - 1ull;
                    lo = (lo & low) | ((lo >> 6) & ~low) | ((hi & 0x3Full) << 54);
                    hi >>= 6;
                } else {
                    const unsigned long long low = (1ull << (6 * (i - 10))) - 1ull;
                    hi = (hi & low) | ((hi >> 6) & ~low);
                }
                len--;
            }
            fs_add_caps(s, p, taken);
            s.w[5] = (s.w[5] & ~(3u << 27)) | ((uint32_t)(p + 1) << 27);        // last_capture = player (:148)
            if (len == 0u) s.w[6] += 1u << (6 * p);                              // scopa (:151-152)
        } else {
            if (len >= (uint32_t)FS_MAX_TABLE) ok = false;
            else if (len < 10u) lo |= (unsigned long long)action << (6 * len);
            else hi |= (unsigned long long)action << (6 * (len - 10u));
            if (ok) len++;
        }
        fs_set_table(s, lo, hi);
        s.w[5] = (s.w[5] & ~(0x1Fu << 22)) | (len << 22);
        s.w[5] &= ~(1u << (16 + 3 * p + hp));                                    // hand.remove(card)
    }
    s.w[6] += 1u << 15;                                                          // step_count++
    bool end_cards = false;
    if (((s.w[5] >> 16) & 0x3Fu) == 0u) {                                        // every hand is empty (:275)
        if (fs_round(s) < 5u) { s.w[6] += 1u << 12; s.w[5] |= 0x3Fu << 16; }     // deal_new_round (:92-99)
        else end_cards = true;
    }
    const bool end_steps = fs_step_count(s) >= FS_MAX_STEPS;                     // safety limit (:286-290)
    if (end_cards || end_steps) {
        const int reps = (end_cards && end_steps) ? 2 : 1;                       // one inlined copy of the scoring code
#pragma unroll 1
        for (int rep = 0; rep < reps; rep++) fs_evaluate(s, rep == 1);
    }
    s.w[5] ^= 1u << 29;                                                          // next agent (:293)
    return ok;
}

__global__ void __launch_bounds__(256) full_init_kernel(const ulonglong4* __restrict__ decks, long long n, uint4* __restrict__ states) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x)
        fs_store(states, g, fs_initial(fs_load_deck(decks, g)));
}

__global__ void __launch_bounds__(FS_THREADS) full_step_kernel(uint4* __restrict__ states, const ulonglong4* __restrict__ decks,
                                                        const uint8_t* __restrict__ actions, float2* __restrict__ rewards,
                                                        uint8_t* __restrict__ done, long long n, unsigned int* overflow) {
    __shared__ uint16_t scratch[FS_MAX_TABLE * FS_THREADS];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        FsState s = fs_load(states, g);
        const FsDeck d = fs_load_deck(decks, g);
        if (!fs_step(s, fs_round_cards(s, d), actions[g], scratch + threadIdx.x)) *overflow = 1u;
        fs_store(states, g, s);
        const bool t = fs_terminal(s);
        if (rewards) {
            const float r0 = t ? 0.5f * (float)fs_score_diff(s) : 0.f;
            rewards[g] = make_float2(r0, 0.f - r0);
        }
        if (done) done[g] = t;
    }
}

__global__ void __launch_bounds__(256) full_legal_kernel(const uint4* __restrict__ states, const ulonglong4* __restrict__ decks,
                                                         int player, uint8_t* __restrict__ ordered, uint8_t* __restrict__ count,
                                                         long long n) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const FsState s = fs_load(states, g);
        const FsDeck d = fs_load_deck(decks, g);
        uint32_t list;
        const uint32_t nl = fs_legal_list(s, fs_round_cards(s, d), player < 0 ? fs_cur(s) : player, list);
        if (ordered) {
            for (int i = 0; i < 3; i++) ordered[3 * g + i] = (uint32_t)i < nl ? (uint8_t)((list >> (8 * i)) & 0xFFu) : (uint8_t)0xFF;
        }
        if (count) count[g] = (uint8_t)nl;
    }
}

// FullScopaGame.evaluate_game() on its own (:174-226): sweep + scoring of whatever the state holds
__global__ void __launch_bounds__(256) full_evaluate_kernel(uint4* __restrict__ states, float2* __restrict__ rewards,
                                                            int* __restrict__ detail, long long n) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        FsState s = fs_load(states, g);
        int det[8];
        fs_evaluate(s, false, det);
        fs_store(states, g, s);
        const float r0 = 0.5f * (float)fs_score_diff(s);
        if (rewards) rewards[g] = make_float2(r0, 0.f - r0);
        if (detail) {
#pragma unroll
            for (int i = 0; i < 8; i++) detail[8 * g + i] = det[i];
        }
    }
}

// n random-policy games played to the end (36 plies) in one launch; "FULL" Philox stream:
// ctr = (game id lo, hi, ply / 4, tag), word ply % 4, action = legal[mulhi32(x, n_legal)].
__global__ void __launch_bounds__(FS_THREADS) full_rollout_kernel(const uint4* __restrict__ states, const ulonglong4* __restrict__ decks,
                                                           long long n, uint2 key, unsigned long long game_offset,
                                                           uint8_t* __restrict__ actions, float2* __restrict__ rewards,
                                                           uint4* __restrict__ final_states, unsigned int* overflow) {
    __shared__ uint16_t scratch[FS_MAX_TABLE * FS_THREADS];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        FsState s = fs_load(states, g);
        const FsDeck d = fs_load_deck(decks, g);
        const unsigned long long gid = game_offset + (unsigned long long)g;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        bool ok = true;
        unsigned long long hc = fs_round_cards(s, d);
        uint32_t* out = actions ? (uint32_t*)(actions + 36 * g) : nullptr;       // 36 bytes per game: 4-byte aligned
        uint32_t packed = 0u;
        // ONE step() site in a rolled loop: unrolled by four this kernel was 100 KB of SASS and its top stall was
        // instruction fetch (profiles/README.md section 7)
#pragma unroll 1
        for (int ply = 0; ply < 36; ply++) {
            const int q = ply & 3;
            if (q == 0) x = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(ply >> 2), MS_TAG_FULL), key);
            const uint32_t xw = q == 0 ? x.x : (q == 1 ? x.y : (q == 2 ? x.z : x.w));
            uint32_t list;
            const uint32_t nl = fs_legal_list(s, hc, fs_cur(s), list);
            const uint32_t a = nl ? (list >> (8u * __umulhi(xw, nl))) & 0xFFu : 0u;
            const uint32_t round_before = fs_round(s);
            ok &= fs_step(s, hc, a, scratch + threadIdx.x);
            if (fs_round(s) != round_before) hc = fs_round_cards(s, d);          // a new hand was dealt
            packed |= a << (8 * q);
            if (q == 3) {
                if (out) out[ply >> 2] = packed;
                packed = 0u;
            }
        }
        if (!ok) *overflow = 1u;
        if (rewards) {
            const float r0 = fs_terminal(s) ? 0.5f * (float)fs_score_diff(s) : 0.f;
            rewards[g] = make_float2(r0, 0.f - r0);
        }
        if (final_states) fs_store(final_states, g, s);
    }
}

}  // namespace ms

#ifndef MS_HOST_RULES_ONLY   // tests/emu/ms_full_host.cpp compiles everything above for the host (CPU checks of the rules)
namespace ms {

// the deck kernel lives in ms_env.cu (it shares the MT19937 seeding with the 16-card deal)
int full_deck_from_seeds(const int64_t* d_seeds, int64_t n, void* d_decks, int zero_means_42, int force_slow, void* stream);

// device flag "a table outgrew FS_MAX_TABLE", one per device (allocated on first use on that device)
static unsigned int* g_overflow[64] = {};
static std::mutex g_overflow_mu;
static int full_overflow_flag(unsigned int** out) {
    int dev = 0;
    MS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(MS_ERR_ARG, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_overflow_mu);
    if (!g_overflow[dev]) {
        MS_CUDA(cudaMalloc(&g_overflow[dev], sizeof(unsigned int)));
        MS_CUDA(cudaMemset(g_overflow[dev], 0, sizeof(unsigned int)));
    }
    *out = g_overflow[dev];
    return MS_OK;
}

}  // namespace ms

using namespace ms;

extern "C" {

int ms_full_deal_from_seeds(const int64_t* d_seeds, int64_t n, ms_full_state* d_states, ms_full_deck* d_decks, void* stream) {
    if (n < 0 || (n > 0 && (!d_seeds || !d_states || !d_decks))) return fail(MS_ERR_ARG, "ms_full_deal_from_seeds: bad argument");
    if (n == 0) return MS_OK;
    int rc = full_deck_from_seeds(d_seeds, n, d_decks, 1, 0, stream);
    if (rc) return rc;
    full_init_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const ulonglong4*)d_decks, (long long)n, (uint4*)d_states);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_full_deck_from_seeds(const int64_t* d_seeds, int64_t n, ms_full_deck* d_decks, int slow_path, void* stream) {
    if (n < 0 || (n > 0 && (!d_seeds || !d_decks))) return fail(MS_ERR_ARG, "ms_full_deck_from_seeds: bad argument");
    if (n == 0) return MS_OK;
    return full_deck_from_seeds(d_seeds, n, d_decks, 0, slow_path ? 1 : 0, stream);
}

int ms_full_step(ms_full_state* d_states, const ms_full_deck* d_decks, const uint8_t* d_actions, float* d_rewards,
                 uint8_t* d_done, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!d_states || !d_decks || !d_actions))) return fail(MS_ERR_ARG, "ms_full_step: bad argument");
    if (n == 0) return MS_OK;
    unsigned int* ov;
    int rc = full_overflow_flag(&ov);
    if (rc) return rc;
    full_step_kernel<<<grid_for(n, FS_THREADS, 8), FS_THREADS, 0, (cudaStream_t)stream>>>((uint4*)d_states, (const ulonglong4*)d_decks, d_actions,
                                                                           (float2*)d_rewards, d_done, (long long)n, ov);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_full_legal_actions(const ms_full_state* d_states, const ms_full_deck* d_decks, int player, uint8_t* d_ordered,
                          uint8_t* d_count, int64_t n, void* stream) {
    if (n < 0 || player < -1 || player > 1 || (n > 0 && (!d_states || !d_decks))) return fail(MS_ERR_ARG, "ms_full_legal_actions: bad argument");
    if (n == 0) return MS_OK;
    full_legal_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)d_states, (const ulonglong4*)d_decks, player,
                                                                            d_ordered, d_count, (long long)n);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_full_rollout_random(const ms_full_state* d_states, const ms_full_deck* d_decks, int64_t n, uint64_t philox_seed,
                           uint64_t game_offset, uint8_t* d_actions, float* d_rewards, ms_full_state* d_final, void* stream) {
    if (n < 0 || (n > 0 && (!d_states || !d_decks))) return fail(MS_ERR_ARG, "ms_full_rollout_random: bad argument");
    if (n == 0) return MS_OK;
    unsigned int* ov;
    int rc = full_overflow_flag(&ov);
    if (rc) return rc;
    full_rollout_kernel<<<grid_for(n, FS_THREADS, 8), FS_THREADS, 0, (cudaStream_t)stream>>>(
        (const uint4*)d_states, (const ulonglong4*)d_decks, (long long)n, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
        (unsigned long long)game_offset, d_actions, (float2*)d_rewards, (uint4*)d_final, ov);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_full_table_overflow(int* h_flag, void* stream) {
    if (!h_flag) return fail(MS_ERR_ARG, "ms_full_table_overflow: bad argument");
    unsigned int* ov;
    int rc = full_overflow_flag(&ov);
    if (rc) return rc;
    unsigned int v = 0;
    MS_CUDA(cudaMemcpyAsync(&v, ov, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *h_flag = (int)v;
    return MS_OK;
}

int ms_full_deal_from_seeds_host(const int64_t* h_seeds, int64_t n, ms_full_state* h_states, ms_full_deck* h_decks) {
    if (n < 0 || (n > 0 && (!h_seeds || !h_states || !h_decks))) return fail(MS_ERR_ARG, "ms_full_deal_from_seeds_host: bad argument");
    if (n == 0) return MS_OK;
    char* d = nullptr;
    const size_t o_d = 32 * (size_t)n, o_s = o_d + 32 * (size_t)n;      // decks | states | seeds
    MS_CUDA(cudaMalloc(&d, o_s + 8 * (size_t)n));
    MS_CUDA(cudaMemcpy(d + o_s, h_seeds, 8 * (size_t)n, cudaMemcpyHostToDevice));
    int rc = ms_full_deal_from_seeds((const int64_t*)(d + o_s), n, (ms_full_state*)(d + o_d), (ms_full_deck*)d, nullptr);
    if (rc) { cudaFree(d); return rc; }
    MS_CUDA(cudaMemcpy(h_decks, d, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(h_states, d + o_d, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaFree(d));
    return MS_OK;
}

int ms_full_step_host(ms_full_state* h_states, const ms_full_deck* h_decks, const uint8_t* h_actions, float* h_rewards,
                      uint8_t* h_done, int64_t n) {
    if (n < 0 || (n > 0 && (!h_states || !h_decks || !h_actions))) return fail(MS_ERR_ARG, "ms_full_step_host: bad argument");
    if (n == 0) return MS_OK;
    char* d = nullptr;
    const size_t o_s = 32 * (size_t)n, o_r = o_s + 32 * (size_t)n, o_a = o_r + 8 * (size_t)n, o_f = o_a + (size_t)n;
    MS_CUDA(cudaMalloc(&d, o_f + (size_t)n + 256));
    MS_CUDA(cudaMemcpy(d, h_decks, 32 * (size_t)n, cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(d + o_s, h_states, 32 * (size_t)n, cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(d + o_a, h_actions, (size_t)n, cudaMemcpyHostToDevice));
    int rc = ms_full_step((ms_full_state*)(d + o_s), (const ms_full_deck*)d, (const uint8_t*)(d + o_a), (float*)(d + o_r),
                          (uint8_t*)(d + o_f), n, nullptr);
    if (rc) { cudaFree(d); return rc; }
    MS_CUDA(cudaMemcpy(h_states, d + o_s, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    if (h_rewards) MS_CUDA(cudaMemcpy(h_rewards, d + o_r, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    if (h_done) MS_CUDA(cudaMemcpy(h_done, d + o_f, (size_t)n, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaFree(d));
    return MS_OK;
}

int ms_full_evaluate_host(ms_full_state* h_states, float* h_rewards, int32_t* h_detail, int64_t n) {
    if (n < 0 || (n > 0 && !h_states)) return fail(MS_ERR_ARG, "ms_full_evaluate_host: bad argument");
    if (n == 0) return MS_OK;
    char* d = nullptr;
    const size_t o_r = 32 * (size_t)n, o_d = o_r + 8 * (size_t)n;
    MS_CUDA(cudaMalloc(&d, o_d + 32 * (size_t)n));
    MS_CUDA(cudaMemcpy(d, h_states, 32 * (size_t)n, cudaMemcpyHostToDevice));
    full_evaluate_kernel<<<grid_for(n, 256, 8), 256>>>((uint4*)d, (float2*)(d + o_r), (int*)(d + o_d), (long long)n);
    ms::g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(d); return fail(MS_ERR_CUDA, "full_evaluate_kernel launch failed: %s", cudaGetErrorString(e)); }
    MS_CUDA(cudaMemcpy(h_states, d, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    if (h_rewards) MS_CUDA(cudaMemcpy(h_rewards, d + o_r, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    if (h_detail) MS_CUDA(cudaMemcpy(h_detail, d + o_d, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaFree(d));
    return MS_OK;
}

int ms_full_rollout_random_host(const int64_t* h_seeds, int64_t n, uint64_t philox_seed, uint64_t game_offset,
                                uint8_t* h_actions, float* h_rewards) {
    if (n < 0 || (n > 0 && !h_seeds)) return fail(MS_ERR_ARG, "ms_full_rollout_random_host: bad argument");
    if (n == 0) return MS_OK;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    const size_t o_k = align256(32 * (size_t)n), o_s = align256(o_k + 8 * (size_t)n), o_r = align256(o_s + 32 * (size_t)n),
                 o_a = align256(o_r + 8 * (size_t)n), tot = align256(o_a + 36 * (size_t)n);   // decks | seeds | states | rewards | actions
    char* d; cudaStream_t st0;
    int rc = scratch_get(tot, &d, &st0);
    if (rc) return rc;
    // same three-stream pipeline as ms_rollout_random_host: H2D of stage c+1, kernels of stage c, D2H of stage c-1
    cudaStream_t pipe[3];
    rc = host_pipe_streams(pipe);
    if (rc) return rc;
    int c = 0;
    for (int64_t lo = 0, m = 0; lo < n; lo += m, c++) {
        m = host_stage_size(lo, n);
        cudaStream_t st = pipe[c % 3];
        MS_CUDA(cudaMemcpyAsync(d + o_k + 8 * lo, h_seeds + lo, 8 * m, cudaMemcpyHostToDevice, st));
        rc = ms_full_deal_from_seeds((const int64_t*)(d + o_k) + lo, m, (ms_full_state*)(d + o_s) + lo, (ms_full_deck*)d + lo, st);
        if (rc) return rc;
        rc = ms_full_rollout_random((const ms_full_state*)(d + o_s) + lo, (const ms_full_deck*)d + lo, m, philox_seed,
                                    game_offset + (uint64_t)lo, (uint8_t*)(d + o_a) + 36 * lo, (float*)(d + o_r) + 2 * lo, nullptr, st);
        if (rc) return rc;
        if (h_actions) MS_CUDA(cudaMemcpyAsync(h_actions + 36 * lo, d + o_a + 36 * lo, 36 * m, cudaMemcpyDeviceToHost, st));
        if (h_rewards) MS_CUDA(cudaMemcpyAsync(h_rewards + 2 * lo, d + o_r + 8 * lo, 8 * m, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < 3; i++) MS_CUDA(cudaStreamSynchronize(pipe[i]));
    return MS_OK;
}

// test hook: the stage schedule both host rollouts (this file's and ms_env.cu's) walk -- size of the stage that starts at game lo of n
int64_t ms_debug_host_stage_size(int64_t lo, int64_t n) { return (lo < 0 || n <= lo) ? 0 : host_stage_size(lo, n); }

}  // extern "C"
#endif  // MS_HOST_RULES_ONLY
