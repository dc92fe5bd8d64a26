// scopa_b200/csrc/ms_env.cu -- env kernels (deal, step, legal moves, capture, infoset keys, fused
// random rollout) and their C-ABI entry points.  sm_100a only.
//
// All kernels are one-thread-per-game grid-stride loops over 16-byte packed states: a warp reads
// and writes 512 contiguous bytes per state access (128-bit per lane), the rules run entirely in
// registers (ms_state.cuh), and grids are sized in multiples of the 148 SMs.
#include <cstring>
#include <mutex>

#include "ms_common.cuh"
#include "ms_state.cuh"

namespace ms {

std::atomic<uint64_t> g_launches{0};
char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

// ------------------------------------------------------------------------------------------------
// K0  deal: CPython random.seed(int) + random.shuffle of the 16-card deck, one thread per seed.
//   replaces MiniDeck.__init__ / MiniScopaGame.reset (src/envs/mini_scopa_game.py:25-28, :56-64).
// random.seed(n) = MT19937 init_by_array(32-bit words of |n|): two dependent passes over the 624-word
// state, the second reading every word the first wrote.  Storing that state per thread (2.5 KB) made
// the first version of this kernel DRAM-bound on local memory (profiles/README.md, r01: 7.5 GB of
// traffic per 1 M seeds).  Pass 1 is therefore RECOMPUTED instead of stored: it runs once to obtain the
// two words pass 2 starts from, then again in lock-step with pass 2, whose results are kept only for
// the ~2x40 state words the first outputs depend on (output j of the first block needs words j, j+1 and
// j+397).  About 1.5x the integer work, no memory traffic.  Pass 1 starts from init_genrand(19650218),
// which does not depend on the seed: that table is computed once on the host, staged in shared memory by
// every CTA and read with one 128-bit broadcast load per four chain steps (profiles/deal_variants.cu
// measures the variants: the chain runs at the integer-ALU rate, 3 ALU + 2 multiplier-pipe instructions
// per step).
__constant__ uint32_t g_mt_init[624];
// MT outputs available on the fast path.  A 16-card shuffle takes 19 draws on average (15 + rejections); the exact
// tail (convolution of the 15 geometric laws) is P(> 40) = 3.2e-4, P(> 48) = 4e-6, P(> 64) = 2.9e-10.  A window of
// 40 sent one warp in a hundred through the slow path below and doubled the kernel's time.
constexpr int DEAL_WIN = 64;

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// random.shuffle of the deck given a word source: for i in reversed(range(1, 16)): j = randbelow(i + 1);
// swap.  randbelow(n): k = n.bit_length(); r = getrandbits(k) = word >> (32 - k); retry while r >= n.
// Returns false if the source ran dry (fast path only).
template <typename Gen>
__device__ __forceinline__ bool shuffle_deck(Gen& gen, unsigned long long& perm) {
    perm = 0xFEDCBA9876543210ull;   // nibble i = card id i (deck order, mini_scopa_game.py:26)
    for (int i = 15; i >= 1; i--) {
        const uint32_t nn = (uint32_t)i + 1u;
        const int kbits = 32 - __clz(nn);
        uint32_t r;
        do {
            uint32_t w;
            if (!gen.next(w)) return false;
            r = w >> (32 - kbits);
        } while (r >= nn);
        unsigned long long ci = (perm >> (4 * i)) & 0xFull, cr = (perm >> (4 * r)) & 0xFull;
        perm &= ~((0xFull << (4 * i)) | (0xFull << (4 * r)));
        perm |= (cr << (4 * i)) | (ci << (4 * r));
    }
    return true;
}

struct MtFull {            // textbook generator over a full 624-word state (slow path)
    uint32_t* mt;
    int kk;
    __device__ __forceinline__ bool next(uint32_t& out) {
        if (kk >= 624) {
            int q;
            for (q = 0; q < 227; q++) {
                uint32_t t = (mt[q] & 0x80000000u) | (mt[q + 1] & 0x7fffffffu);
                mt[q] = mt[q + 397] ^ (t >> 1) ^ ((t & 1u) ? 0x9908b0dfu : 0u);
            }
            for (; q < 623; q++) {
                uint32_t t = (mt[q] & 0x80000000u) | (mt[q + 1] & 0x7fffffffu);
                mt[q] = mt[q - 227] ^ (t >> 1) ^ ((t & 1u) ? 0x9908b0dfu : 0u);
            }
            uint32_t t = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (t >> 1) ^ ((t & 1u) ? 0x9908b0dfu : 0u);
            kk = 0;
        }
        out = mt_temper(mt[kk++]);
        return true;
    }
};

// Slow path (a shuffle that needs more than DEAL_WIN draws: 3 in 10^10 seeds): the whole state in
// local memory, textbook init_by_array + generator.
// textbook init_by_array over a full 624-word state (slow paths only)
__device__ __noinline__ void mt_seed_full(uint32_t key0, uint32_t key1, uint32_t* mt) {
    const bool two = key1 != 0u;
    for (int i = 0; i < 624; i++) mt[i] = g_mt_init[i];
    int i = 1, j = 0;
    for (int k = 0; k < 624; k++) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + ((j == 1) ? key1 : key0) + (uint32_t)j;
        i++; j++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
        if (j >= (two ? 2 : 1)) j = 0;
    }
    for (int k = 0; k < 623; k++) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
}

__device__ __noinline__ unsigned long long deal_slow(uint32_t key0, uint32_t key1) {
    uint32_t mt[624];
    mt_seed_full(key0, key1, mt);
    MtFull gen{mt, 624};
    unsigned long long perm;
    shuffle_deck(gen, perm);
    return perm;
}

// random.seed(n) for the key words of |n|: fills lo[0 .. WIN + 1] = mt[0 .. WIN + 1] and hi[0 .. WIN - 1] =
// mt[397 .. 397 + WIN - 1] of the seeded state (see the header comment: pass 1 is recomputed, not stored).
// T4 = the init_genrand(19650218) table in shared memory.  Word i is written by step k = i - 1 with key word
// j = k % len, so even words take key[1] + 1 (or key[0] for one-word keys) and odd words key[0].
#define MS_P1STEP(x, t, kw) (x) = ((t) ^ (((x) ^ ((x) >> 30)) * 1664525u)) + (kw)
#define MS_P2STEP(y, p, i) (y) = ((p) ^ (((y) ^ ((y) >> 30)) * 1566083941u)) - (uint32_t)(i)
#define MS_LOCK4(body)                                                   \
    {                                                                    \
        const uint4 t = T4[q];                                           \
        const int i = 4 * q;                                             \
        MS_P1STEP(p1, t.x, kodd); MS_P2STEP(p2, p1, i);     body(i)      \
        MS_P1STEP(p1, t.y, key0); MS_P2STEP(p2, p1, i + 1); body(i + 1)  \
        MS_P1STEP(p1, t.z, kodd); MS_P2STEP(p2, p1, i + 2); body(i + 2)  \
        MS_P1STEP(p1, t.w, key0); MS_P2STEP(p2, p1, i + 3); body(i + 3)  \
    }
#define MS_KEEP_NONE(w)
#define MS_KEEP_LO(w) if ((w) <= WIN + 1) lo[(w)] = p2;
#define MS_KEEP_HI(w) if ((w) >= 397 && (w) < 397 + WIN) hi[(w) - 397] = p2;

template <int WIN>
__device__ __forceinline__ void mt_seed_window(const uint4* __restrict__ T4, uint32_t key0, uint32_t key1,
                                               uint32_t* lo, uint32_t* hi) {
    static_assert(WIN % 4 == 0 && WIN + 4 < 396 && 397 + WIN < 620, "window must be inside the first generator block");
    const bool two = key1 != 0u;   // key length: 32-bit words of |seed|, at least one
    const uint32_t kodd = two ? key1 + 1u : key0;   // key[j] + j for j = 1 / one-word keys

    // ---- pass 1, first run (nothing stored)
    const uint4 t0 = T4[0];
    uint32_t prev = (t0.y ^ ((t0.x ^ (t0.x >> 30)) * 1664525u)) + key0;   // word 1
    const uint32_t first1 = prev;
    MS_P1STEP(prev, t0.z, kodd);
    MS_P1STEP(prev, t0.w, key0);
#pragma unroll 4
    for (int q = 1; q < 156; q++) {
        const uint4 t = T4[q];
        MS_P1STEP(prev, t.x, kodd); MS_P1STEP(prev, t.y, key0); MS_P1STEP(prev, t.z, kodd); MS_P1STEP(prev, t.w, key0);
    }
    // step 624 wraps: mt[0] = mt[623]; word 1 is rewritten with j = 623 % len
    const uint32_t m1w = (first1 ^ ((prev ^ (prev >> 30)) * 1664525u)) + kodd;

    // ---- pass 2 (words 2..623, then the wrap to word 1) in lock-step with a second run of pass 1
    uint32_t p1 = first1, p2 = m1w;
    MS_P1STEP(p1, t0.z, kodd); MS_P2STEP(p2, p1, 2); lo[2] = p2;
    MS_P1STEP(p1, t0.w, key0); MS_P2STEP(p2, p1, 3); lo[3] = p2;
    constexpr int Q_LO = (WIN + 2 + 3) / 4;       // groups 1 .. Q_LO - 1 hold the kept words 4 .. WIN + 1
    constexpr int Q_HI = (397 + WIN + 3) / 4;     // groups 99 .. Q_HI - 1 hold the kept words 397 .. 397 + WIN - 1
#pragma unroll 1
    for (int q = 1; q < Q_LO; q++) MS_LOCK4(MS_KEEP_LO)
#pragma unroll 4
    for (int q = Q_LO; q < 99; q++) MS_LOCK4(MS_KEEP_NONE)
#pragma unroll 1
    for (int q = 99; q < Q_HI; q++) MS_LOCK4(MS_KEEP_HI)
#pragma unroll 4
    for (int q = Q_HI; q < 156; q++) MS_LOCK4(MS_KEEP_NONE)
    lo[1] = (m1w ^ ((p2 ^ (p2 >> 30)) * 1566083941u)) - 1u;
    lo[0] = 0x80000000u;
}

__device__ __forceinline__ void stage_mt_table(uint4* T4) {
    for (int i = threadIdx.x; i < 624; i += blockDim.x) ((uint32_t*)T4)[i] = g_mt_init[i];
    __syncthreads();
}

// random.shuffle on the fast path.  The reference's loop is "for each position: draw until accepted"
// (random.py: shuffle -> _randbelow_with_getrandbits).  Run in that order, a warp pays its worst lane's
// rejections at EVERY position and the lanes read different outputs; here every lane looks at output kk
// in the same iteration and either accepts it for its current position or not -- the same draws in the
// same order per lane, uniform addressing, and a warp's trip count is its worst lane's TOTAL number of draws.
// Returns false if the window ran dry.
template <int WIN>
__device__ __forceinline__ bool shuffle_deck_window(const uint32_t* lo, const uint32_t* hi, unsigned long long& perm) {
    perm = 0xFEDCBA9876543210ull;   // nibble i = card id i (deck order, mini_scopa_game.py:26)
    int i = 15;
    uint32_t a = lo[0];
    for (int kk = 0; kk < WIN && i >= 1; kk++) {
        const uint32_t b = lo[kk + 1];
        uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
        y = hi[kk] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        a = b;
        const uint32_t nn = (uint32_t)i + 1u;
        const uint32_t r = mt_temper(y) >> __clz(nn);   // getrandbits(nn.bit_length())
        if (r < nn) {
            const unsigned long long d = ((perm >> (4 * i)) ^ (perm >> (4 * r))) & 0xFull;
            perm ^= (d << (4 * i)) | (d << (4 * r));
            i--;
        }
    }
    return i < 1;
}

__global__ void __launch_bounds__(256) deal_kernel(const long long* __restrict__ seeds, long long n,
                                                   uint4* __restrict__ states, uint32_t* __restrict__ hand_order,
                                                   unsigned long long* __restrict__ deck, int zero_means_42) {
    __shared__ uint4 T4[156];
    stage_mt_table(T4);
    uint32_t lo[DEAL_WIN + 2], hi[DEAL_WIN];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        long long sd = seeds[g];
        if (sd == 0 && zero_means_42) sd = 42;   // `seed or self.seed` (mini_scopa_game.py:132, default seed 42)
        unsigned long long a = sd < 0 ? (unsigned long long)(-(sd + 1)) + 1ull : (unsigned long long)sd;
        const uint32_t key0 = (uint32_t)a, key1 = (uint32_t)(a >> 32);
        mt_seed_window<DEAL_WIN>(T4, key0, key1, lo, hi);
        unsigned long long perm;
        if (!shuffle_deck_window<DEAL_WIN>(lo, hi, perm)) perm = deal_slow(key0, key1);
        const uint32_t ord = (uint32_t)perm;   // first 8 dealt cards: 4 to player 0, 4 to player 1
        uint32_t h0 = 0u, h1 = 0u;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            h0 |= 1u << ((ord >> (4 * i)) & 0xFu);
            h1 |= 1u << ((ord >> (16 + 4 * i)) & 0xFu);
        }
        if (states) states[g] = st_make(h0, h1, 8u);
        if (hand_order) hand_order[g] = ord;
        if (deck) deck[g] = perm;
    }
}

// ------------------------------------------------------------------------------------------------
// 40-card deck: FullDeck(seed) = random.seed + random.shuffle of ids 0..39 (src/envs/full_scopa_game.py:32-35).
// Packed as four 64-bit words, ten 6-bit card ids each (position p -> word p / 10, bits 6 * (p % 10)).
constexpr int FULL_WIN = 128;   // a 40-card shuffle takes 60 draws on average (39 + rejections); P(> 128) = 8e-14

struct Deck40 {
    unsigned long long w[4];
    __device__ __forceinline__ uint32_t get(int p) const {
        const int q = p / 10;
        const unsigned long long x = q == 0 ? w[0] : (q == 1 ? w[1] : (q == 2 ? w[2] : w[3]));
        return (uint32_t)(x >> (6 * (p % 10))) & 0x3Fu;
    }
    __device__ __forceinline__ void set(int p, uint32_t c) {
        const int q = p / 10, sh = 6 * (p % 10);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k == q) w[k] = (w[k] & ~(0x3Full << sh)) | ((unsigned long long)c << sh);
    }
};

template <typename Gen>
__device__ __forceinline__ bool shuffle_deck40(Gen& gen, Deck40& d) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        d.w[q] = 0ull;
        for (int k = 0; k < 10; k++) d.w[q] |= (unsigned long long)(10 * q + k) << (6 * k);
    }
    for (int i = 39; i >= 1; i--) {
        const uint32_t nn = (uint32_t)i + 1u;
        const int kbits = 32 - __clz(nn);
        uint32_t r;
        do {
            uint32_t w;
            if (!gen.next(w)) return false;
            r = w >> (32 - kbits);
        } while (r >= nn);
        const uint32_t ci = d.get(i), cr = d.get((int)r);
        d.set(i, cr); d.set((int)r, ci);
    }
    return true;
}

// fast path: the same draws with the warp-uniform loop order of shuffle_deck_window.  The deck being shuffled
// lives in shared memory, one 32-bit word per card, column = thread (bank = lane: conflict-free whatever the
// row), so a swap is two loads and two stores instead of select chains over four packed 64-bit registers.
constexpr int FULL_DECK_THREADS = 128;
__device__ __forceinline__ bool shuffle_deck40_window(const uint32_t* lo, const uint32_t* hi, uint32_t* col, Deck40& d) {
#pragma unroll
    for (int p = 0; p < 40; p++) col[p * FULL_DECK_THREADS] = (uint32_t)p;
    int i = 39;
    uint32_t a = lo[0];
    for (int kk = 0; kk < FULL_WIN && i >= 1; kk++) {
        const uint32_t b = lo[kk + 1];
        uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
        y = hi[kk] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        a = b;
        const uint32_t nn = (uint32_t)i + 1u;
        const uint32_t r = mt_temper(y) >> __clz(nn);
        if (r < nn) {
            const uint32_t ci = col[i * FULL_DECK_THREADS], cr = col[r * FULL_DECK_THREADS];
            col[i * FULL_DECK_THREADS] = cr;
            col[r * FULL_DECK_THREADS] = ci;
            i--;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        unsigned long long w = 0ull;
#pragma unroll
        for (int k = 0; k < 10; k++) w |= (unsigned long long)col[(10 * q + k) * FULL_DECK_THREADS] << (6 * k);
        d.w[q] = w;
    }
    return i < 1;
}

__device__ __noinline__ void full_deck_slow(uint32_t key0, uint32_t key1, Deck40& d) {
    uint32_t mt[624];
    mt_seed_full(key0, key1, mt);
    MtFull gen{mt, 624};
    shuffle_deck40(gen, d);
}

__global__ void __launch_bounds__(FULL_DECK_THREADS) full_deck_kernel(const long long* __restrict__ seeds, long long n,
                                                        ulonglong4* __restrict__ decks, int zero_means_42, int force_slow) {
    __shared__ uint4 T4[156];
    __shared__ uint32_t deck_s[40 * FULL_DECK_THREADS];
    stage_mt_table(T4);
    uint32_t lo[FULL_WIN + 2], hi[FULL_WIN];
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        long long sd = seeds[g];
        if (sd == 0 && zero_means_42) sd = 42;   // `seed or self.seed` (full_scopa_game.py:244, default seed 42)
        unsigned long long a = sd < 0 ? (unsigned long long)(-(sd + 1)) + 1ull : (unsigned long long)sd;
        const uint32_t key0 = (uint32_t)a, key1 = (uint32_t)(a >> 32);
        Deck40 d;
        bool ok = false;
        if (!force_slow) {
            mt_seed_window<FULL_WIN>(T4, key0, key1, lo, hi);
            ok = shuffle_deck40_window(lo, hi, deck_s + threadIdx.x, d);
        }
        if (!ok) full_deck_slow(key0, key1, d);
        decks[g] = make_ulonglong4(d.w[0], d.w[1], d.w[2], d.w[3]);
    }
}

// test hook: every seed through the slow path (exercised by tests/test_gpu_env.py)
__global__ void __launch_bounds__(128) deal_slow_kernel(const long long* __restrict__ seeds, long long n,
                                                        uint4* __restrict__ states, uint32_t* __restrict__ hand_order) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        long long sd = seeds[g];
        if (sd == 0) sd = 42;
        unsigned long long a = sd < 0 ? (unsigned long long)(-(sd + 1)) + 1ull : (unsigned long long)sd;
        const uint32_t ord = (uint32_t)deal_slow((uint32_t)a, (uint32_t)(a >> 32));
        uint32_t h0 = 0u, h1 = 0u;
        for (int i = 0; i < 4; i++) {
            h0 |= 1u << ((ord >> (4 * i)) & 0xFu);
            h1 |= 1u << ((ord >> (16 + 4 * i)) & 0xFu);
        }
        states[g] = st_make(h0, h1, 8u);
        hand_order[g] = ord;
    }
}

// ------------------------------------------------------------------------------------------------
// K1  step / legal / capture / key
__global__ void __launch_bounds__(256) step_kernel(uint4* __restrict__ states, const uint8_t* __restrict__ actions,
                                                   float2* __restrict__ rewards, uint8_t* __restrict__ done,
                                                   long long n) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        MsState s = states[g];
        step(s, (uint32_t)actions[g]);
        states[g] = s;
        const bool term = st_terminal(s);
        if (rewards) {
            float r0 = term ? reward0(s) : 0.f;
            rewards[g] = make_float2(r0, 0.f - r0);   // r1 = s1 - mean: +0.0 on a tie, never -0.0
        }
        if (done) done[g] = term ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) legal_kernel(const uint4* __restrict__ states,
                                                    const uint32_t* __restrict__ hand_order, int player,
                                                    uint16_t* __restrict__ mask, uchar4* __restrict__ ordered,
                                                    uint8_t* __restrict__ count, uchar4* __restrict__ capture,
                                                    long long n) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        const MsState s = states[g];
        const int p = player < 0 ? st_cur(s) : player;
        uint32_t list;
        const uint32_t nl = legal_list(s, hand_order[g], p, list);
        uint32_t m = 0u;
        uint8_t o[4] = {0xFF, 0xFF, 0xFF, 0xFF}, c[4] = {0, 0, 0, 0};
        const uint32_t hand = st_hand(s, p);
        const uint32_t tset = table_set(s.y, st_table_len(s));
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if ((uint32_t)k < nl) {
                uint32_t a = (list >> (4 * k)) & 0xFu;
                m |= 1u << a;
                o[k] = (uint8_t)a;
                // the fallback action [0] on an empty hand is a pass: it captures nothing
                c[k] = ((hand >> a) & 1u) ? (uint8_t)capture_mask(s.y, st_table_len(s), a, tset) : 0;
            }
        }
        if (mask) mask[g] = (uint16_t)m;
        if (ordered) ordered[g] = make_uchar4(o[0], o[1], o[2], o[3]);
        if (count) count[g] = (uint8_t)nl;
        if (capture) capture[g] = make_uchar4(c[0], c[1], c[2], c[3]);
    }
}

__global__ void __launch_bounds__(256) capture_kernel(const uint4* __restrict__ states,
                                                      const uint8_t* __restrict__ cards,
                                                      uint8_t* __restrict__ out, long long n) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        const MsState s = states[g];
        const uint32_t card = cards[g] & 0xFu, len = st_table_len(s);
        const uint32_t tset = table_set(s.y, len);
        uint32_t m = capture_mask(s.y, len, card, tset);
        // card_in_table() of a card that itself lies on the table (never the case for a card played from
        // a hand): it has the rank too, so the first of {card, twin} in table order is taken (:72-74)
        if ((tset >> card) & 1u) {
            const uint32_t self = 1u << nibble_pos(s.y, card);
            m = ((tset >> card_twin(card)) & 1u) ? (m < self ? m : self) : self;
        }
        out[g] = (uint8_t)m;
    }
}

__global__ void __launch_bounds__(256) keys_kernel(const uint4* __restrict__ states, int player,
                                                   unsigned long long* __restrict__ keys, long long n) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        const MsState s = states[g];
        const int p = player < 0 ? st_cur(s) : player;
        keys[g] = st_terminal(s) ? 0xFFFFFFFFFFFFFFFFull : infoset_key(s, p);
    }
}

// ------------------------------------------------------------------------------------------------
// K2  fused random rollout: the whole 8-ply game in registers; 20 B in, 8 B actions + 8 B rewards
// (+16 B final state) out per game.
__global__ void __launch_bounds__(256) rollout_kernel(const uint4* __restrict__ states,
                                                      const uint32_t* __restrict__ hand_order, long long n,
                                                      uint2 key, unsigned long long game_offset,
                                                      uint2* __restrict__ actions8, float2* __restrict__ rewards,
                                                      uint4* __restrict__ final_states) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n;
         g += (long long)gridDim.x * blockDim.x) {
        MsState s = states[g];
        const uint32_t ho = hand_order[g];
        const uint32_t dealt = dealt_set(s);
        const unsigned long long gid = game_offset + (unsigned long long)g;
        // the ply loop is deliberately NOT unrolled: one copy of step() (about 6 KB of SASS) stays resident in
        // the instruction cache instead of eight (the unrolled kernel was 50 KB, beyond the 32 KB L1.5 I-cache)
        unsigned long long acts = 0ull;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        // Both hands as ordered nibble lists (deal order restricted to the cards still held: the order of
        // legal_actions()), built once and edited as cards are played -- the per-ply legal_list() compaction was the
        // largest single line of this kernel (11 % of its instructions, ncu source view).
        uint32_t hl = 0u;
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            uint32_t lst;
            MsState t = s;
            t.w &= ~(1u << 18);                                 // list the hand even if the state is terminal
            const uint32_t cnt = legal_list(t, ho, pl, lst);
            if (st_hand(s, pl) != 0u && cnt) hl |= (lst & 0xFFFFu) << (16 * pl);
        }
#pragma unroll 1
        for (int ply = 0; ply < 8; ply++) {
            if ((ply & 3) == 0)
                x = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(ply >> 2), MS_TAG_ROLL), key);
            const int q = ply & 3;
            const uint32_t xw = q == 0 ? x.x : (q == 1 ? x.y : (q == 2 ? x.z : x.w));
            const int pl = st_cur(s);
            const uint32_t nl = st_terminal(s) ? 0u : (uint32_t)__popc(st_hand(s, pl));
            uint32_t a = 0u;                                    // empty hand: legal_actions() == [0]; terminal: no-op
            if (nl) {
                const uint32_t idx = __umulhi(xw, nl);
                const uint32_t lst = (hl >> (16 * pl)) & 0xFFFFu;
                a = (lst >> (4u * idx)) & 0xFu;
                hl = (hl & ~(0xFFFFu << (16 * pl))) | (nibble_remove(lst, idx) << (16 * pl));
            }
            step(s, a, table_set_from_dealt(s, dealt));
            acts |= (unsigned long long)a << (8 * ply);
        }
        if (actions8) actions8[g] = make_uint2((uint32_t)acts, (uint32_t)(acts >> 32));
        if (rewards) {
            float r0 = st_terminal(s) ? reward0(s) : 0.f;
            rewards[g] = make_float2(r0, 0.f - r0);   // r1 = s1 - mean: +0.0 on a tie, never -0.0
        }
        if (final_states) final_states[g] = s;
    }
}

}  // namespace ms

#ifndef MS_HOST_RULES_ONLY   // tests/emu/ms_env_host.cpp compiles every kernel above for the host (CPU checks)
namespace ms {
// ------------------------------------------------------------------------------------------------
static std::mutex g_mt_mu;
static bool g_mt_done[64] = {};

// uploads the seed-independent MT19937 start table once per device; a failed upload is retried by the next call
// (a once_flag would be consumed by the failure and later launches would read an empty table)
static int ensure_mt_table() {
    int dev = 0;
    MS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(MS_ERR_ARG, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_mt_mu);
    if (g_mt_done[dev]) return MS_OK;
    uint32_t t[624];
    t[0] = 19650218u;   // init_genrand(19650218), the seed-independent start of init_by_array
    for (int i = 1; i < 624; i++) t[i] = 1812433253u * (t[i - 1] ^ (t[i - 1] >> 30)) + (uint32_t)i;
    cudaError_t err = cudaMemcpyToSymbol(g_mt_init, t, sizeof(t));
    if (err != cudaSuccess) return fail(MS_ERR_CUDA, "uploading MT table failed: %s", cudaGetErrorString(err));
    g_mt_done[dev] = true;
    return MS_OK;
}

// scratch for the host-buffer entry points (grown on demand, one per device, guarded by a mutex)
struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
};
static Scratch g_scratch[64];
std::mutex g_scratch_mu;

int scratch_get(size_t bytes, char** out, cudaStream_t* stream) {
    int dev = 0;
    MS_CUDA(cudaGetDevice(&dev));
    Scratch& sc = g_scratch[dev & 63];
    if (!sc.stream) MS_CUDA(cudaStreamCreateWithFlags(&sc.stream, cudaStreamNonBlocking));
    if (sc.bytes < bytes) {
        if (sc.p) MS_CUDA(cudaFree(sc.p));
        sc.p = nullptr; sc.bytes = 0;
        size_t want = bytes + bytes / 4 + 4096;
        MS_CUDA(cudaMalloc(&sc.p, want));
        sc.bytes = want;
    }
    *out = (char*)sc.p;
    *stream = sc.stream;
    return MS_OK;
}

constexpr int64_t MS_HOST_CHUNK_DEFAULT = 262144;
std::atomic<int64_t> g_host_chunk{MS_HOST_CHUNK_DEFAULT};   // games per pipeline stage of the *_host rollouts
int64_t host_chunk() { return g_host_chunk.load(); }
int64_t host_stage_size(int64_t lo, int64_t n) {
    const int64_t chunk = g_host_chunk.load();
    const int64_t q = ((chunk / 4 + 127) / 128) * 128;          // stages stay 128-byte aligned slices of every array
    const int64_t left = n - lo;
    if (n <= chunk + 2 * q) return left < chunk ? left : chunk;  // small calls: plain chunks
    if (lo == 0) return q;
    if (left > chunk + q) return chunk;
    if (left > q) return ((left - q + 127) / 128) * 128 < left ? ((left - q + 127) / 128) * 128 : left;
    return left;
}

int host_pipe_streams(cudaStream_t out[3]) {
    static cudaStream_t pipe[64][3] = {};
    int dev = 0;
    MS_CUDA(cudaGetDevice(&dev));
    for (int i = 0; i < 3; i++) {
        if (!pipe[dev & 63][i]) MS_CUDA(cudaStreamCreateWithFlags(&pipe[dev & 63][i], cudaStreamNonBlocking));
        out[i] = pipe[dev & 63][i];
    }
    return MS_OK;
}

}  // namespace ms

using namespace ms;

extern "C" {

int ms_abi_version(void) { return MS_ABI_VERSION; }
const char* ms_last_error(void) { return last_error_buf(); }
uint64_t ms_launch_count(void) { return g_launches.load(); }

int ms_deal_from_seeds(const int64_t* d_seeds, int64_t n, ms_state* d_states, uint32_t* d_hand_order, void* stream) {
    if (n < 0 || (n > 0 && (!d_seeds || !d_states || !d_hand_order))) return fail(MS_ERR_ARG, "ms_deal_from_seeds: bad argument");
    if (n == 0) return MS_OK;
    int rc = ensure_mt_table();
    if (rc) return rc;
    deal_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)d_seeds, (long long)n, (uint4*)d_states, d_hand_order, nullptr, 1);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

}  // extern "C"

namespace ms {
// deck with the env-level seed rule (`seed or self.seed`: 0 means 42) -- used by the team deal (ms_team.cu)
int team_deck_from_seeds(const int64_t* d_seeds, int64_t n, uint64_t* d_deck, void* stream) {
    int rc = ensure_mt_table();
    if (rc) return rc;
    deal_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)d_seeds, (long long)n, nullptr, nullptr, (unsigned long long*)d_deck, 1);
    MS_LAUNCH_CHECK();
    return MS_OK;
}
}  // namespace ms

namespace ms {
// FullDeck(seed) for n seeds (used by ms_full.cu); zero_means_42 = the env-level `seed or self.seed` rule
int full_deck_from_seeds(const int64_t* d_seeds, int64_t n, void* d_decks, int zero_means_42, int force_slow, void* stream) {
    int rc = ensure_mt_table();
    if (rc) return rc;
    full_deck_kernel<<<grid_for(n, FULL_DECK_THREADS, 8), FULL_DECK_THREADS, 0, (cudaStream_t)stream>>>(
        (const long long*)d_seeds, (long long)n, (ulonglong4*)d_decks, zero_means_42, force_slow);
    MS_LAUNCH_CHECK();
    return MS_OK;
}
}  // namespace ms

extern "C" {

int ms_deck_from_seeds(const int64_t* d_seeds, int64_t n, uint64_t* d_deck, void* stream) {
    if (n < 0 || (n > 0 && (!d_seeds || !d_deck))) return fail(MS_ERR_ARG, "ms_deck_from_seeds: bad argument");
    if (n == 0) return MS_OK;
    int rc = ensure_mt_table();
    if (rc) return rc;
    deal_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)d_seeds, (long long)n, nullptr, nullptr, (unsigned long long*)d_deck, 0);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_debug_deal_slow_path(const int64_t* d_seeds, int64_t n, ms_state* d_states, uint32_t* d_hand_order, void* stream) {
    if (n < 0 || (n > 0 && (!d_seeds || !d_states || !d_hand_order))) return fail(MS_ERR_ARG, "ms_debug_deal_slow_path: bad argument");
    if (n == 0) return MS_OK;
    int rc = ensure_mt_table();
    if (rc) return rc;
    deal_slow_kernel<<<grid_for(n, 128, 8), 128, 0, (cudaStream_t)stream>>>((const long long*)d_seeds, (long long)n,
                                                                           (uint4*)d_states, d_hand_order);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_step(ms_state* d_states, const uint8_t* d_actions, float* d_rewards, uint8_t* d_done, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!d_states || !d_actions))) return fail(MS_ERR_ARG, "ms_step: bad argument");
    if (n == 0) return MS_OK;
    step_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((uint4*)d_states, d_actions, (float2*)d_rewards,
                                                                      d_done, (long long)n);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_legal_actions(const ms_state* d_states, const uint32_t* d_hand_order, int player, uint16_t* d_mask,
                     uint8_t* d_ordered, uint8_t* d_count, uint8_t* d_capture, int64_t n, void* stream) {
    if (n < 0 || player > 1 || (n > 0 && (!d_states || !d_hand_order))) return fail(MS_ERR_ARG, "ms_legal_actions: bad argument");
    if (n == 0) return MS_OK;
    legal_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)d_states, d_hand_order, player, d_mask,
                                                                       (uchar4*)d_ordered, d_count, (uchar4*)d_capture,
                                                                       (long long)n);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_capture(const ms_state* d_states, const uint8_t* d_cards, uint8_t* d_table_pos_mask, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!d_states || !d_cards || !d_table_pos_mask))) return fail(MS_ERR_ARG, "ms_capture: bad argument");
    if (n == 0) return MS_OK;
    capture_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)d_states, d_cards, d_table_pos_mask,
                                                                         (long long)n);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_infoset_keys(const ms_state* d_states, int player, uint64_t* d_keys, int64_t n, void* stream) {
    if (n < 0 || player > 1 || (n > 0 && (!d_states || !d_keys))) return fail(MS_ERR_ARG, "ms_infoset_keys: bad argument");
    if (n == 0) return MS_OK;
    keys_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)d_states, player,
                                                                      (unsigned long long*)d_keys, (long long)n);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_rollout_random(const ms_state* d_states, const uint32_t* d_hand_order, int64_t n, uint64_t philox_seed,
                      uint64_t game_offset, uint8_t* d_actions, float* d_rewards, ms_state* d_final, void* stream) {
    if (n < 0 || (n > 0 && (!d_states || !d_hand_order))) return fail(MS_ERR_ARG, "ms_rollout_random: bad argument");
    if (n == 0) return MS_OK;
    rollout_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)d_states, d_hand_order, (long long)n, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
        (unsigned long long)game_offset, (uint2*)d_actions, (float2*)d_rewards, (uint4*)d_final);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

// ---------------------------------------------------------------------------- host-buffer forms
int ms_deal_from_seeds_host(const int64_t* h_seeds, int64_t n, ms_state* h_states, uint32_t* h_hand_order) {
    if (n < 0 || (n > 0 && (!h_seeds || !h_states || !h_hand_order))) return fail(MS_ERR_ARG, "ms_deal_from_seeds_host: bad argument");
    if (n == 0) return MS_OK;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    size_t o_seed = 0, o_st = align256(o_seed + 8 * n), o_ho = align256(o_st + 16 * n), tot = align256(o_ho + 4 * n);
    char* d; cudaStream_t st;
    int rc = scratch_get(tot, &d, &st);
    if (rc) return rc;
    MS_CUDA(cudaMemcpyAsync(d + o_seed, h_seeds, 8 * n, cudaMemcpyHostToDevice, st));
    rc = ms_deal_from_seeds((const int64_t*)(d + o_seed), n, (ms_state*)(d + o_st), (uint32_t*)(d + o_ho), st);
    if (rc) return rc;
    MS_CUDA(cudaMemcpyAsync(h_states, d + o_st, 16 * n, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaMemcpyAsync(h_hand_order, d + o_ho, 4 * n, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    return MS_OK;
}

int ms_step_host(ms_state* h_states, const uint8_t* h_actions, float* h_rewards, uint8_t* h_done, int64_t n) {
    if (n < 0 || (n > 0 && (!h_states || !h_actions))) return fail(MS_ERR_ARG, "ms_step_host: bad argument");
    if (n == 0) return MS_OK;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    size_t o_st = 0, o_a = align256(16 * n), o_r = align256(o_a + n), o_d = align256(o_r + 8 * n), tot = align256(o_d + n);
    char* d; cudaStream_t st;
    int rc = scratch_get(tot, &d, &st);
    if (rc) return rc;
    MS_CUDA(cudaMemcpyAsync(d + o_st, h_states, 16 * n, cudaMemcpyHostToDevice, st));
    MS_CUDA(cudaMemcpyAsync(d + o_a, h_actions, n, cudaMemcpyHostToDevice, st));
    rc = ms_step((ms_state*)(d + o_st), (const uint8_t*)(d + o_a), (float*)(d + o_r), (uint8_t*)(d + o_d), n, st);
    if (rc) return rc;
    MS_CUDA(cudaMemcpyAsync(h_states, d + o_st, 16 * n, cudaMemcpyDeviceToHost, st));
    if (h_rewards) MS_CUDA(cudaMemcpyAsync(h_rewards, d + o_r, 8 * n, cudaMemcpyDeviceToHost, st));
    if (h_done) MS_CUDA(cudaMemcpyAsync(h_done, d + o_d, n, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    return MS_OK;
}

int ms_legal_actions_host(const ms_state* h_states, const uint32_t* h_hand_order, int player, uint16_t* h_mask,
                          uint8_t* h_ordered, uint8_t* h_count, uint8_t* h_capture, int64_t n) {
    if (n < 0 || (n > 0 && (!h_states || !h_hand_order))) return fail(MS_ERR_ARG, "ms_legal_actions_host: bad argument");
    if (n == 0) return MS_OK;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    size_t o_st = 0, o_ho = align256(16 * n), o_m = align256(o_ho + 4 * n), o_o = align256(o_m + 2 * n),
           o_c = align256(o_o + 4 * n), o_cap = align256(o_c + n), tot = align256(o_cap + 4 * n);
    char* d; cudaStream_t st;
    int rc = scratch_get(tot, &d, &st);
    if (rc) return rc;
    MS_CUDA(cudaMemcpyAsync(d + o_st, h_states, 16 * n, cudaMemcpyHostToDevice, st));
    MS_CUDA(cudaMemcpyAsync(d + o_ho, h_hand_order, 4 * n, cudaMemcpyHostToDevice, st));
    rc = ms_legal_actions((const ms_state*)(d + o_st), (const uint32_t*)(d + o_ho), player, (uint16_t*)(d + o_m),
                          (uint8_t*)(d + o_o), (uint8_t*)(d + o_c), (uint8_t*)(d + o_cap), n, st);
    if (rc) return rc;
    if (h_mask) MS_CUDA(cudaMemcpyAsync(h_mask, d + o_m, 2 * n, cudaMemcpyDeviceToHost, st));
    if (h_ordered) MS_CUDA(cudaMemcpyAsync(h_ordered, d + o_o, 4 * n, cudaMemcpyDeviceToHost, st));
    if (h_count) MS_CUDA(cudaMemcpyAsync(h_count, d + o_c, n, cudaMemcpyDeviceToHost, st));
    if (h_capture) MS_CUDA(cudaMemcpyAsync(h_capture, d + o_cap, 4 * n, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    return MS_OK;
}

int ms_infoset_keys_host(const ms_state* h_states, int player, uint64_t* h_keys, int64_t n) {
    if (n < 0 || (n > 0 && (!h_states || !h_keys))) return fail(MS_ERR_ARG, "ms_infoset_keys_host: bad argument");
    if (n == 0) return MS_OK;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    size_t o_st = 0, o_k = align256(16 * n), tot = align256(o_k + 8 * n);
    char* d; cudaStream_t st;
    int rc = scratch_get(tot, &d, &st);
    if (rc) return rc;
    MS_CUDA(cudaMemcpyAsync(d + o_st, h_states, 16 * n, cudaMemcpyHostToDevice, st));
    rc = ms_infoset_keys((const ms_state*)(d + o_st), player, (uint64_t*)(d + o_k), n, st);
    if (rc) return rc;
    MS_CUDA(cudaMemcpyAsync(h_keys, d + o_k, 8 * n, cudaMemcpyDeviceToHost, st));
    MS_CUDA(cudaStreamSynchronize(st));
    return MS_OK;
}

int64_t ms_debug_set_host_chunk(int64_t games) {
    g_host_chunk.store(games > 0 ? ((games + 127) / 128) * 128 : MS_HOST_CHUNK_DEFAULT);
    return g_host_chunk.load();
}

int ms_rollout_random_host(const int64_t* h_seeds, int64_t n, uint64_t philox_seed, uint64_t game_offset,
                           uint8_t* h_actions, float* h_rewards) {
    if (n < 0 || (n > 0 && !h_seeds)) return fail(MS_ERR_ARG, "ms_rollout_random_host: bad argument");
    if (n == 0) return MS_OK;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    size_t o_seed = 0, o_st = align256(8 * n), o_ho = align256(o_st + 16 * n), o_a = align256(o_ho + 4 * n),
           o_r = align256(o_a + 8 * n), tot = align256(o_r + 8 * n);
    char* d; cudaStream_t st0;
    int rc = scratch_get(tot, &d, &st0);
    if (rc) return rc;
    // Chunked over three streams so that the H2D copy of chunk c+1, the kernels of chunk c and the D2H copy of
    // chunk c-1 overlap (separate copy engines per direction); the chunks are 128-byte aligned slices.
    cudaStream_t pipe[3];
    rc = host_pipe_streams(pipe);
    if (rc) return rc;
    int c = 0;
    for (int64_t lo = 0, m = 0; lo < n; lo += m, c++) {
        m = host_stage_size(lo, n);
        cudaStream_t st = pipe[c % 3];
        MS_CUDA(cudaMemcpyAsync(d + o_seed + 8 * lo, h_seeds + lo, 8 * m, cudaMemcpyHostToDevice, st));
        rc = ms_deal_from_seeds((const int64_t*)(d + o_seed) + lo, m, (ms_state*)(d + o_st) + lo, (uint32_t*)(d + o_ho) + lo, st);
        if (rc) return rc;
        rc = ms_rollout_random((const ms_state*)(d + o_st) + lo, (const uint32_t*)(d + o_ho) + lo, m, philox_seed,
                               game_offset + (uint64_t)lo, (uint8_t*)(d + o_a) + 8 * lo, (float*)(d + o_r) + 2 * lo, nullptr, st);
        if (rc) return rc;
        if (h_actions) MS_CUDA(cudaMemcpyAsync(h_actions + 8 * lo, d + o_a + 8 * lo, 8 * m, cudaMemcpyDeviceToHost, st));
        if (h_rewards) MS_CUDA(cudaMemcpyAsync(h_rewards + 2 * lo, d + o_r + 8 * lo, 8 * m, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < 3; i++) MS_CUDA(cudaStreamSynchronize(pipe[i]));
    return MS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Atomic-throughput microbenchmarks (there is no published atomic peak for this part, so the
// roofline the MCCFR kernel's atomics are compared with is measured on the box).  Addresses are pseudo-random
// over a table the size of the MCCFR delta table (738 x 4 doubles) -- the uncontended / L2-resident case.
namespace ms {
constexpr int ATOM_TABLE = 738 * 4;

__global__ void __launch_bounds__(768, 1) atom_smem_f64_kernel(double* sink, int iters) {
    __shared__ double tab[ATOM_TABLE];
    for (int i = threadIdx.x; i < ATOM_TABLE; i += blockDim.x) tab[i] = 0.0;
    __syncthreads();
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        atomicAdd(&tab[(x >> 8) % ATOM_TABLE], 1.0);
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = tab[0] + tab[ATOM_TABLE - 1];
}

__global__ void __launch_bounds__(768, 1) atom_smem_u32_kernel(double* sink, int iters) {
    __shared__ uint32_t tab[ATOM_TABLE];
    for (int i = threadIdx.x; i < ATOM_TABLE; i += blockDim.x) tab[i] = 0u;
    __syncthreads();
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        atomicAdd(&tab[(x >> 8) % ATOM_TABLE], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = (double)(tab[0] + tab[ATOM_TABLE - 1]);
}

__global__ void __launch_bounds__(768, 1) atom_global_f64_kernel(double* tab, int iters) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        atomicAdd(&tab[(x >> 8) % ATOM_TABLE], 1.0);      // result unused -> RED.E.ADD.F64
    }
}
}  // namespace ms

extern "C" int ms_debug_atomic_peaks(double h_out[3], void* stream) {
    if (!h_out) return fail(MS_ERR_ARG, "ms_debug_atomic_peaks: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    double* d = nullptr;
    MS_CUDA(cudaMalloc(&d, sizeof(double) * (ATOM_TABLE + kNumSMs)));
    MS_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (ATOM_TABLE + kNumSMs), st));
    cudaEvent_t e0, e1;
    MS_CUDA(cudaEventCreate(&e0)); MS_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, grid = kNumSMs, block = 768;
    const double ops = (double)iters * grid * block;
    for (int which = 0; which < 3; which++) {
        float ms_best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {          // first repetition is the warm-up
            MS_CUDA(cudaEventRecord(e0, st));
            if (which == 0) atom_smem_f64_kernel<<<grid, block, 0, st>>>(d + ATOM_TABLE, iters);
            else if (which == 1) atom_smem_u32_kernel<<<grid, block, 0, st>>>(d + ATOM_TABLE, iters);
            else atom_global_f64_kernel<<<grid, block, 0, st>>>(d, iters);
            MS_LAUNCH_CHECK();
            MS_CUDA(cudaEventRecord(e1, st));
            MS_CUDA(cudaEventSynchronize(e1));
            float ms_t = 0.f;
            MS_CUDA(cudaEventElapsedTime(&ms_t, e0, e1));
            if (rep > 0 && ms_t < ms_best) ms_best = ms_t;
        }
        h_out[which] = ops / (ms_best * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    MS_CUDA(cudaFree(d));
    return MS_OK;
}
#endif  // MS_HOST_RULES_ONLY
