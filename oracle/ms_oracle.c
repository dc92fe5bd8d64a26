/* oracle/ms_oracle.c -- CPU restatement of the reference's Miniscopa hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY (see ms_oracle.h).  Deliberately written with the reference's own data
 * shapes (ordered lists of cards, a dict of per-infoset arrays keyed by the info STRING, plain
 * recursion) and NOT with the packed bitboards / hash slots the CUDA product uses, so that the two
 * implementations share no code and no representation.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off matters: the reference's float64 arithmetic is separate numpy multiply/add
 * ufunc loops (no FMA), and the CFR parity test is bit-exact.
 *
 * Citations are relative to /root/reference/.
 */
#include "ms_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ================================================================================ cards */
/* src/envs/mini_scopa_game.py:17-23 -- id = suit_idx*4 + card_idx (:149-153) */
static const int RANK[16] = {2, 5, 8, 10, 2, 5, 7, 9, 3, 6, 8, 9, 3, 6, 7, 10};
static const char SUITC[4] = {'c', 'f', 'p', 'b'}; /* cuori fiori picche bello */

/* ================================================================================ MT19937 */
/* CPython Modules/_randommodule.c and numpy legacy RandomState share this generator. */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt_t;

static void mt_init_genrand(mt_t* m, uint32_t s) {
    m->mt[0] = s;
    for (int i = 1; i < 624; i++)
        m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}

static void mt_init_by_array(mt_t* m, const uint32_t* key, int len) {
    mt_init_genrand(m, 19650218u);
    int i = 1, j = 0;
    int k = 624 > len ? 624 : len;
    for (; k; k--) {
        m->mt[i] = (m->mt[i] ^ ((m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { m->mt[0] = m->mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = 623; k; k--) {
        m->mt[i] = (m->mt[i] ^ ((m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { m->mt[0] = m->mt[623]; i = 1; }
    }
    m->mt[0] = 0x80000000u;
}

static uint32_t mt_next(mt_t* m) {
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    uint32_t y;
    if (m->idx >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (m->mt[kk] & 0x80000000u) | (m->mt[kk + 1] & 0x7fffffffu);
            m->mt[kk] = m->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < 623; kk++) {
            y = (m->mt[kk] & 0x80000000u) | (m->mt[kk + 1] & 0x7fffffffu);
            m->mt[kk] = m->mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (m->mt[623] & 0x80000000u) | (m->mt[0] & 0x7fffffffu);
        m->mt[623] = m->mt[396] ^ (y >> 1) ^ mag01[y & 1u];
        m->idx = 0;
    }
    y = m->mt[m->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* random.seed(int): key = 32-bit little-endian words of abs(seed), at least one word */
static void py_random_seed(mt_t* m, int64_t seed) {
    uint64_t a = seed < 0 ? (uint64_t)(-(seed + 1)) + 1u : (uint64_t)seed;
    uint32_t key[2] = {(uint32_t)(a & 0xffffffffu), (uint32_t)(a >> 32)};
    mt_init_by_array(m, key, key[1] ? 2 : 1);
}

static int bit_length(uint32_t n) {
    int k = 0;
    while (n) { k++; n >>= 1; }
    return k;
}

/* Random._randbelow_with_getrandbits: k = n.bit_length(); r = getrandbits(k); while r >= n: retry */
static uint32_t py_randbelow(mt_t* m, uint32_t n) {
    int k = bit_length(n);
    uint32_t r = mt_next(m) >> (32 - k);
    while (r >= n) r = mt_next(m) >> (32 - k);
    return r;
}

/* MiniDeck.__init__ (src/envs/mini_scopa_game.py:25-28): cards in id order, random.seed(seed),
 * random.shuffle(cards) == for i in reversed(range(1, 16)): j = randbelow(i+1); swap */
void ora_deck(int64_t seed, int out16[16]) {
    mt_t m;
    py_random_seed(&m, seed);
    for (int i = 0; i < 16; i++) out16[i] = i;
    for (int i = 15; i >= 1; i--) {
        int j = (int)py_randbelow(&m, (uint32_t)(i + 1));
        int t = out16[i]; out16[i] = out16[j]; out16[j] = t;
    }
}

/* ================================================================================ Philox4x32-10 */
/* Counter-based stream shared BY SPECIFICATION (DESIGN.md section "Random streams") with the CUDA
 * kernels; implemented independently here. */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define TAG_ROLL 0x4C4C4F52u /* "ROLL" */
#define TAG_MCCF 0x4643434Du /* "MCCF" */
#define TAG_MCCF_SEQ (TAG_MCCF + 64u) /* the sequential stream of the static-shape kernels */
#define TAG_SDCF 0x46434453u /* "SDCF" */

/* ================================================================================ RNG facade */
struct ora_rng {
    int kind;            /* 0 numpy legacy MT19937, 1 Philox addressed by call index, 2 Philox addressed by draw index */
    mt_t mt;
    uint64_t seed;
    /* Philox addressing, set by the traversal before each draw */
    uint32_t tag; uint64_t trav; uint32_t call;
    uint32_t draw;       /* kind 2: number of uniforms this traversal has consumed so far */
};

ora_rng* ora_rng_new(int kind, uint64_t seed) {
    ora_rng* r = (ora_rng*)calloc(1, sizeof(ora_rng));
    r->kind = kind; r->seed = seed;
    if (kind == 0) mt_init_genrand(&r->mt, (uint32_t)seed); /* np.random.seed(int) -> init_genrand */
    return r;
}
void ora_rng_free(ora_rng* r) { free(r); }

/* Philox addressing of the traversal streams: one 4-word block serves two consecutive call indices:
 * ctr = (traversal id lo, hi, call >> 1, tag); call c uses words 2*(c&1) and 2*(c&1)+1. */
static void rng_philox(const ora_rng* r, uint32_t out2[2]) {
    uint32_t ctr[4] = {(uint32_t)r->trav, (uint32_t)(r->trav >> 32), r->call >> 1, r->tag};
    uint32_t key[2] = {(uint32_t)r->seed, (uint32_t)(r->seed >> 32)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    out2[0] = o[2 * (r->call & 1u)]; out2[1] = o[2 * (r->call & 1u) + 1];
}

/* kind 2 ("sequential" stream of the headline MCCFR kernel): the d-th uniform a traversal
 * consumes is word d & 3 of the block with ctr = (traversal id lo, hi, d >> 2, tag), as a 31-bit fraction
 * u = (word >> 1) / 2^31 (exactly representable, so `cdf <= u` below is the comparison numpy's searchsorted makes on
 * this u; 31 bits so that ceil(cdf * 2^31), the integer the kernel compares with, fits 32 bits even for cdf = 1). */
static double rng_seq32(ora_rng* r) {
    uint32_t ctr[4] = {(uint32_t)r->trav, (uint32_t)(r->trav >> 32), r->draw >> 2, r->tag};
    uint32_t key[2] = {(uint32_t)r->seed, (uint32_t)(r->seed >> 32)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    double u = (double)(o[r->draw & 3u] >> 1) / 2147483648.0;
    r->draw++;
    return u;
}

/* uniform double in [0,1): numpy random_sample == (a>>5, b>>6) 53-bit; same formula on Philox words */
static double rng_double(ora_rng* r) {
    uint32_t a, b;
    if (r->kind == 2) return rng_seq32(r);
    if (r->kind == 0) { a = mt_next(&r->mt) >> 5; b = mt_next(&r->mt) >> 6; }
    else { uint32_t o[2]; rng_philox(r, o); a = o[0] >> 5; b = o[1] >> 6; }
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* np.random.choice(n items, p): cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(u, 'right') */
static int rng_choice_p(ora_rng* r, const double* p, int n) {
    if (r->kind == 2 && n == 1) return 0;   /* sequential stream: a forced move consumes no uniform */
    double cdf[16];
    double acc = 0.0;
    for (int i = 0; i < n; i++) { acc += p[i]; cdf[i] = acc; }
    double last = cdf[n - 1];
    for (int i = 0; i < n; i++) cdf[i] /= last;
    double u = rng_double(r);
    int idx = 0;
    while (idx < n && cdf[idx] <= u) idx++;
    if (idx >= n) idx = n - 1;
    return idx;
}

/* np.random.choice(n items) without p == legacy randint(0, n): masked rejection on 32-bit words,
 * no draw when n == 1.  Philox: mulhi(out[0], n). */
static int rng_choice_uniform(ora_rng* r, int n) {
    if (r->kind == 0) {
        uint32_t rng = (uint32_t)(n - 1);
        if (rng == 0) return 0;
        uint32_t mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = mt_next(&r->mt) & mask; } while (v > rng);
        return (int)v;
    }
    uint32_t o[2]; rng_philox(r, o);
    return (int)(((uint64_t)o[0] * (uint64_t)n) >> 32);
}

/* ================================================================================ game rules */
/* MiniScopaGame.card_in_table (src/envs/mini_scopa_game.py:66-91).  Returns the number of captured
 * table cards and writes their table positions (in the order the reference lists them). */
int ora_card_in_table(const int* table, int ntable, int card, int* pos_out) {
    int target = RANK[card];
    if (target <= 0 || ntable == 0) return 0;
    for (int i = 0; i < ntable; i++)                       /* :72-74 first exact match */
        if (RANK[table[i]] == target) { pos_out[0] = i; return 1; }
    /* :76-85  comb_sums[s] = tuple of indices whose ranks sum to s (first found) */
    int have[11]; int comb[11][8]; int clen[11];
    for (int s = 0; s <= target; s++) { have[s] = 0; clen[s] = 0; }
    have[0] = 1;
    for (int idx = 0; idx < ntable; idx++) {
        int r = RANK[table[idx]];
        for (int s = target; s >= r; s--) {
            if (!have[s] && have[s - r]) {
                have[s] = 1;
                clen[s] = clen[s - r] + 1;
                for (int q = 0; q < clen[s - r]; q++) comb[s][q] = comb[s - r][q];
                comb[s][clen[s - r]] = idx;
            }
        }
    }
    if (!have[target]) return 0;
    for (int q = 0; q < clen[target]; q++) pos_out[q] = comb[target][q];
    return clen[target];
}

/* MiniScopaGame.play_card (:93-104) */
static void play_card(ora_env* e, int pl, int hand_pos) {
    int card = e->hand[pl][hand_pos];
    int pos[8];
    int ncap = ora_card_in_table(e->table, e->ntable, card, pos);
    if (ncap > 0) {
        int captured[8];
        for (int q = 0; q < ncap; q++) captured[q] = e->table[pos[q]];
        /* self.table.remove(c) for each captured card (cards are unique objects) */
        int keep[16]; int nk = 0;
        for (int i = 0; i < e->ntable; i++) {
            int gone = 0;
            for (int q = 0; q < ncap; q++) if (pos[q] == i) gone = 1;
            if (!gone) keep[nk++] = e->table[i];
        }
        for (int i = 0; i < nk; i++) e->table[i] = keep[i];
        e->ntable = nk;
        for (int q = 0; q < ncap; q++) e->caps[pl][e->ncaps[pl]++] = captured[q]; /* captured + [card] */
        e->caps[pl][e->ncaps[pl]++] = card;
        if (e->ntable == 0) e->scopas[pl] += 1;
    } else {
        e->table[e->ntable++] = card;
    }
    for (int i = hand_pos; i + 1 < e->nhand[pl]; i++) e->hand[pl][i] = e->hand[pl][i + 1];
    e->nhand[pl]--;
}

/* evaluate_game (:106-114) */
void ora_env_evaluate(const ora_env* e, double out[2]) {
    int r0 = e->ncaps[0] + 2 * e->scopas[0];
    int r1 = e->ncaps[1] + 2 * e->scopas[1];
    int total = r0 + r1;
    if (total == 0) { out[0] = 0; out[1] = 0; return; }
    double mean = (double)total / 2;
    out[0] = r0 - mean; out[1] = r1 - mean;
}

/* MiniScopaGame.reset(seed) (:56-64) + MiniScopaEnv.reset (:131-138) */
void ora_env_reset(ora_env* e, int64_t seed, int has_seed) {
    int64_t s = (has_seed && seed != 0) ? seed : e->seed;   /* `seed or self.seed` (:132) */
    int deck[16];
    ora_deck(s, deck);
    e->ntable = 0;
    for (int p = 0; p < 2; p++) {
        e->ncaps[p] = 0; e->scopas[p] = 0; e->nhand[p] = 4;
        for (int i = 0; i < 4; i++) e->hand[p][i] = deck[p * 4 + i];
        e->rewards[p] = 0; e->term[p] = 0;
    }
    e->agent = 0;
    e->step_count = 0;
}

void ora_env_init(ora_env* e, int64_t seed) {
    memset(e, 0, sizeof(*e));
    e->max_steps = 2 * 4;          /* :128 */
    e->seed = seed;
    ora_env_reset(e, seed, 1);     /* :130 self.reset(seed) */
}

/* MiniScopaEnv.step (:140-167) */
void ora_env_step(ora_env* e, int action) {
    if (e->term[e->agent]) return;                 /* :141-143 dead step (no-op with the AECEnv shim) */
    int pl = e->agent;
    int hp = -1;
    if (action >= 0 && action < 16)                /* :149-155 action -> (rank, suit) -> card in hand */
        for (int i = 0; i < e->nhand[pl]; i++) if (e->hand[pl][i] == action) { hp = i; break; }
    if (hp >= 0) play_card(e, pl, hp);             /* else: silent pass (:156-157) */
    e->step_count += 1;
    int all_empty = (e->nhand[0] == 0 && e->nhand[1] == 0);
    if (all_empty || e->step_count >= e->max_steps) {
        double r[2]; ora_env_evaluate(e, r);
        e->rewards[0] = r[0]; e->rewards[1] = r[1];
        e->term[0] = e->term[1] = 1;
    }
    e->agent = (pl + 1) % 2;
}

/* ---- OpenSpiel wrapper (src/envs/openspiel_mini_scopa.py) */
void ora_state_init(ora_state* s, int64_t seed) {
    memset(s, 0, sizeof(*s));
    ora_env_init(&s->env, seed);          /* MiniScopaEnv(num_players) then env.reset() (:11-13) */
    ora_env_reset(&s->env, 0, 0);
    s->is_terminal = 0; s->nhist = 0;
}
void ora_state_clone(const ora_state* s, ora_state* out) {
    *out = *s;
    out->env.max_steps = 16;              /* :108 */
}
void ora_state_apply(ora_state* s, int action) {   /* :49-53 */
    if (s->nhist < 40) s->history[s->nhist++] = action;
    ora_env_step(&s->env, action);
    s->is_terminal = s->env.term[0] && s->env.term[1];
}
int ora_state_current_player(const ora_state* s) { return s->is_terminal ? -4 : s->env.agent; }

int ora_state_legal(const ora_state* s, int player, int* out) {  /* :22-47 */
    if (s->is_terminal) return 0;
    if (player < 0) player = ora_state_current_player(s);
    int n = s->env.nhand[player];
    for (int i = 0; i < n; i++) out[i] = s->env.hand[player][i];
    if (n == 0) { out[0] = 0; return 1; }   /* :47 fallback */
    return n;
}

static int fmt_cards(char* buf, const int* cards, int n) {
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (i) buf[k++] = '-';
        k += sprintf(buf + k, "%d%c", RANK[cards[i]], SUITC[cards[i] / 4]);
    }
    buf[k] = 0;
    return k;
}

int ora_state_info_string(const ora_state* s, int player, char* buf, int buflen) {  /* :86-95 */
    if (player == -100) player = ora_state_current_player(s);
    if (s->is_terminal || player < 0) return snprintf(buf, buflen, "TERMINAL");
    char h[64], t[64];
    fmt_cards(h, s->env.hand[player], s->env.nhand[player]);
    fmt_cards(t, s->env.table, s->env.ntable);
    return snprintf(buf, buflen, "P%d:H[%s]_T[%s]", player, h, t);
}

void ora_state_rewards(const ora_state* s, double out[2]) {   /* :78-81 */
    if (!s->is_terminal) { out[0] = 0; out[1] = 0; return; }
    out[0] = s->env.rewards[0]; out[1] = s->env.rewards[1];
}

int ora_state_history_str(const ora_state* s, char* buf, int buflen) {   /* :70-76 */
    char h[256]; int k = 0; h[0] = 0;
    for (int i = 0; i < s->nhist; i++) k += sprintf(h + k, i ? "-%d" : "%d", s->history[i]);
    if (s->is_terminal) {
        double r[2]; ora_state_rewards(s, r);
        return snprintf(buf, buflen, "TERMINAL:%s:%.2f,%.2f", h, r[0], r[1]);
    }
    return snprintf(buf, buflen, "H:%s:P%d", h, ora_state_current_player(s));
}

/* ================================================================================ batched helpers */
void ora_batch_deal(const int64_t* seeds, int64_t n, int* hands) {
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < n; g++) {
        int deck[16];
        int64_t s = seeds[g] != 0 ? seeds[g] : 42;  /* env-level `seed or self.seed` */
        ora_deck(s, deck);
        for (int i = 0; i < 8; i++) hands[g * 8 + i] = deck[i];
    }
}

void ora_rollout_random(const int64_t* seeds, int64_t n, uint64_t philox_seed, uint64_t game_offset,
                        uint8_t* actions, float* rewards, uint8_t* scopas, uint8_t* ncaps, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < n; g++) {
        ora_env e;
        memset(&e, 0, sizeof(e));
        e.max_steps = 8; e.seed = 42;
        ora_env_reset(&e, seeds[g], 1);
        uint32_t key[2] = {(uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)};
        for (int ply = 0; ply < 8; ply++) {
            int pl = e.agent;
            int nl = e.nhand[pl];
            /* "ROLL" stream: one Philox block serves 4 plies: ctr = (game lo, hi, ply/4, tag), word ply%4 */
            uint64_t gid = game_offset + (uint64_t)g;
            uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(ply >> 2), TAG_ROLL};
            uint32_t o[4];
            philox4x32_10(ctr, key, o);
            int a = nl > 0 ? e.hand[pl][(int)(((uint64_t)o[ply & 3] * (uint64_t)nl) >> 32)] : 0;
            actions[g * 8 + ply] = (uint8_t)a;
            ora_env_step(&e, a);
        }
        rewards[g * 2 + 0] = (float)e.rewards[0]; rewards[g * 2 + 1] = (float)e.rewards[1];
        if (scopas) { scopas[g * 2] = (uint8_t)e.scopas[0]; scopas[g * 2 + 1] = (uint8_t)e.scopas[1]; }
        if (ncaps) { ncaps[g * 2] = (uint8_t)e.ncaps[0]; ncaps[g * 2 + 1] = (uint8_t)e.ncaps[1]; }
    }
}

/* ================================================================================ info-set table */
typedef struct {
    char key[48];
    int nlegal; int legal[4];
    double regret[4], strategy[4], local[4];   /* InfoNode (vanilla_cfr.py:8-21 / mc_cfr.py:9-18) */
} ora_node;

struct ora_table {
    ora_node* nodes; int n, cap;
    int* hidx; int hcap;
};

static uint32_t fnv1a(const char* s) {
    uint32_t h = 2166136261u;
    while (*s) { h ^= (uint8_t)*s++; h *= 16777619u; }
    return h;
}

ora_table* ora_table_new(void) {
    ora_table* t = (ora_table*)calloc(1, sizeof(ora_table));
    t->cap = 4096; t->nodes = (ora_node*)calloc(t->cap, sizeof(ora_node));
    t->hcap = 16384; t->hidx = (int*)malloc(sizeof(int) * t->hcap);
    for (int i = 0; i < t->hcap; i++) t->hidx[i] = -1;
    return t;
}
void ora_table_free(ora_table* t) { if (t) { free(t->nodes); free(t->hidx); free(t); } }
int ora_table_size(const ora_table* t) { return t->n; }
const char* ora_table_key(const ora_table* t, int i) { return t->nodes[i].key; }
int ora_table_nlegal(const ora_table* t, int i) { return t->nodes[i].nlegal; }
const int* ora_table_legal(const ora_table* t, int i) { return t->nodes[i].legal; }
double* ora_table_regret(ora_table* t, int i) { return t->nodes[i].regret; }
double* ora_table_strategy(ora_table* t, int i) { return t->nodes[i].strategy; }

int ora_table_find(const ora_table* t, const char* key) {
    uint32_t h = fnv1a(key) & (uint32_t)(t->hcap - 1);
    while (t->hidx[h] >= 0) {
        if (strcmp(t->nodes[t->hidx[h]].key, key) == 0) return t->hidx[h];
        h = (h + 1) & (uint32_t)(t->hcap - 1);
    }
    return -1;
}

/* _get_or_create_node (vanilla_cfr.py:51-54) / _get_node (mc_cfr.py:32-35) */
static ora_node* table_get(ora_table* t, const char* key, const int* legal, int nlegal) {
    uint32_t h = fnv1a(key) & (uint32_t)(t->hcap - 1);
    while (t->hidx[h] >= 0) {
        if (strcmp(t->nodes[t->hidx[h]].key, key) == 0) return &t->nodes[t->hidx[h]];
        h = (h + 1) & (uint32_t)(t->hcap - 1);
    }
    if (t->n >= t->cap) { fprintf(stderr, "ora_table full\n"); abort(); }
    ora_node* nd = &t->nodes[t->n];
    memset(nd, 0, sizeof(*nd));
    strncpy(nd->key, key, sizeof(nd->key) - 1);
    nd->nlegal = nlegal;
    for (int i = 0; i < nlegal; i++) { nd->legal[i] = legal[i]; nd->local[i] = 1.0 / nlegal; }
    t->hidx[h] = t->n++;
    return nd;
}

/* InfoNode.get_strategy (vanilla_cfr.py:23-30) == current_strategy (mc_cfr.py:20-24) */
static void regret_matching(const double* regret, int n, double* out) {
    double pos[4]; double norm = 0.0;
    for (int i = 0; i < n; i++) { pos[i] = regret[i] > 0 ? regret[i] : 0.0; }
    for (int i = 0; i < n; i++) norm += pos[i];
    if (norm > 0) for (int i = 0; i < n; i++) out[i] = pos[i] / norm;
    else for (int i = 0; i < n; i++) out[i] = 1.0 / n;
}

/* ================================================================================ vanilla CFR */
/* CFRTrainer._cfr_recursive (src/algorithms/vanilla_cfr.py:56-99) */
double ora_cfr_recursive(ora_table* t, const ora_state* s, int tp, double r0, double r1) {
    if (s->is_terminal) { double r[2]; ora_state_rewards(s, r); return r[tp]; }
    int cp = ora_state_current_player(s);
    char key[64];
    ora_state_info_string(s, cp, key, sizeof key);
    int legal[4]; int n = ora_state_legal(s, -1, legal);
    ora_node* nd = table_get(t, key, legal, n);
    double au[4];
    for (int i = 0; i < n; i++) {
        ora_state c; ora_state_clone(s, &c); ora_state_apply(&c, legal[i]);
        if (cp == 0) au[i] = ora_cfr_recursive(t, &c, tp, r0 * nd->local[i], r1);
        else         au[i] = ora_cfr_recursive(t, &c, tp, r0, r1 * nd->local[i]);
    }
    double util = 0.0;                                 /* np.sum(local_strategy * action_utils) */
    for (int i = 0; i < n; i++) util += nd->local[i] * au[i];
    if (cp == tp) {
        double reach = tp == 0 ? r0 : r1;
        double opp = tp == 0 ? r1 : r0;
        for (int i = 0; i < n; i++) {
            double regret = au[i] - util;
            nd->regret[i] += opp * regret;
            nd->strategy[i] += reach * nd->local[i];
        }
    }
    regret_matching(nd->regret, n, nd->local);         /* :97 refreshed after EVERY visit */
    return util;
}

/* CFRTrainer.train (:105-110) */
void ora_cfr_train(ora_table* t, int64_t seed, int iters) {
    for (int it = 0; it < iters; it++)
        for (int p = 0; p < 2; p++) {
            ora_state s; ora_state_init(&s, seed);
            ora_cfr_recursive(t, &s, p, 1.0, 1.0);
        }
}

/* ================================================================================ MCCFR */
typedef struct {
    ora_table* t; ora_rng* rng; int tp;
    /* batch mode: deltas (indexed like t->nodes) instead of in-place updates */
    double (*dreg)[4]; double (*dstr)[4];
    int64_t n_updates, n_visits;
} mccfr_ctx;

static void mccfr_key(const ora_state* s, int player, char* key, int len) {
    char info[64];
    ora_state_info_string(s, player, info, sizeof info);
    snprintf(key, len, "%d|%s", player, info);
}

/* MCCFRTrainer._sample (src/algorithms/mc_cfr.py:37-86) */
static double mccfr_sample(mccfr_ctx* c, const ora_state* s, const double reach[2], const double samp[2]) {
    uint32_t my_call = c->rng->call++;                 /* Philox addressing: index of this call */
    c->n_visits++;
    if (s->is_terminal) { double r[2]; ora_state_rewards(s, r); return r[c->tp]; }
    int player = ora_state_current_player(s);
    char key[72]; mccfr_key(s, player, key, sizeof key);
    int legal[4]; int n = ora_state_legal(s, player, legal);
    ora_node* nd = table_get(c->t, key, legal, n);
    int slot = (int)(nd - c->t->nodes);
    double sigma[4];
    regret_matching(nd->regret, n, sigma);             /* :54 */
    uint32_t saved = c->rng->call; c->rng->call = my_call;
    int ai = rng_choice_p(c->rng, sigma, n);           /* :55 np.random.choice(legal, p=sigma) */
    c->rng->call = saved;
    ora_state nx; ora_state_clone(s, &nx); ora_state_apply(&nx, legal[ai]);
    double nreach[2] = {reach[0], reach[1]}, nsamp[2] = {samp[0], samp[1]};
    if (player == c->tp) nsamp[player] *= sigma[ai];
    else { nreach[player] *= sigma[ai]; nsamp[player] *= sigma[ai]; }
    double util = mccfr_sample(c, &nx, nreach, nsamp); /* :67 */
    if (player == c->tp) {
        double cfv[4];
        for (int i = 0; i < n; i++) {                  /* :71-78 */
            ora_state tmp; ora_state_clone(s, &tmp); ora_state_apply(&tmp, legal[i]);
            double tsamp[2] = {samp[0], samp[1]};
            tsamp[player] *= sigma[i];
            cfv[i] = mccfr_sample(c, &tmp, reach, tsamp);
        }
        double v = 0.0;                                /* np.dot(sigma, cfv_all) */
        for (int i = 0; i < n; i++) v += sigma[i] * cfv[i];
        double opp = reach[1 - player];                /* np.prod over p != player */
        double w = samp[player] > 0 ? opp / samp[player] : 0.0;
        nd = &c->t->nodes[slot];
        for (int i = 0; i < n; i++) {
            double dr = w * (cfv[i] - v);
            double ds = reach[player] * sigma[i];
            if (c->dreg) { c->dreg[slot][i] += dr; c->dstr[slot][i] += ds; }
            else { nd->regret[i] += dr; nd->strategy[i] += ds; }
        }
        c->n_updates++;
    }
    return util;
}

/* MCCFRTrainer.iteration (:88-92) x iters, in-place like the reference */
void ora_mccfr_iterate(ora_table* t, int64_t seed, int iters, ora_rng* rng, uint64_t first_iter) {
    for (int it = 0; it < iters; it++)
        for (int p = 0; p < 2; p++) {
            ora_state s; ora_state_init(&s, seed);
            mccfr_ctx c = {t, rng, p, NULL, NULL, 0, 0};
            rng->tag = TAG_MCCF + (uint32_t)p; rng->trav = (first_iter + (uint64_t)it); rng->call = 0;
            double one[2] = {1.0, 1.0};
            mccfr_sample(&c, &s, one, one);
        }
}

static void populate_rec(ora_table* t, const ora_state* s) {
    if (s->is_terminal) return;
    int player = ora_state_current_player(s);
    char key[72]; mccfr_key(s, player, key, sizeof key);
    int legal[4]; int n = ora_state_legal(s, player, legal);
    table_get(t, key, legal, n);
    for (int i = 0; i < n; i++) {
        ora_state c; ora_state_clone(s, &c); ora_state_apply(&c, legal[i]);
        populate_rec(t, &c);
    }
}
void ora_mccfr_populate(ora_table* t, int64_t seed) {
    ora_state s; ora_state_init(&s, seed);
    populate_rec(t, &s);
}

static void mccfr_batch_impl(ora_table* t, int64_t seed, int player, uint64_t philox_seed, uint64_t first_trav,
                             int64_t ntrav, int64_t* n_updates, int64_t* n_visits, int rng_kind, uint32_t tag_base) {
    int n0 = t->n;
    double (*dreg)[4] = calloc((size_t)t->cap, sizeof(double[4]));
    double (*dstr)[4] = calloc((size_t)t->cap, sizeof(double[4]));
    ora_rng* rng = ora_rng_new(rng_kind, philox_seed);
    mccfr_ctx c = {t, rng, player, dreg, dstr, 0, 0};
    for (int64_t k = 0; k < ntrav; k++) {
        ora_state s; ora_state_init(&s, seed);
        rng->tag = tag_base + (uint32_t)player; rng->trav = first_trav + (uint64_t)k; rng->call = 0; rng->draw = 0;
        double one[2] = {1.0, 1.0};
        mccfr_sample(&c, &s, one, one);
    }
    if (t->n != n0) { fprintf(stderr, "ora_mccfr_batch: table was not pre-populated\n"); abort(); }
    for (int i = 0; i < t->n; i++)
        for (int a = 0; a < 4; a++) { t->nodes[i].regret[a] += dreg[i][a]; t->nodes[i].strategy[a] += dstr[i][a]; }
    if (n_updates) *n_updates = c.n_updates;
    if (n_visits) *n_visits = c.n_visits;
    ora_rng_free(rng); free(dreg); free(dstr);
}

void ora_mccfr_batch(ora_table* t, int64_t seed, int player, uint64_t philox_seed,
                     uint64_t first_trav, int64_t ntrav, int64_t* n_updates, int64_t* n_visits) {
    mccfr_batch_impl(t, seed, player, philox_seed, first_trav, ntrav, n_updates, n_visits, 1, TAG_MCCF);
}

/* the same frozen-sigma batch on the sequential 32-bit stream (what ms_mccfr_batch's static-shape kernel consumes) */
void ora_mccfr_batch_seq(ora_table* t, int64_t seed, int player, uint64_t philox_seed,
                         uint64_t first_trav, int64_t ntrav, int64_t* n_updates, int64_t* n_visits) {
    mccfr_batch_impl(t, seed, player, philox_seed, first_trav, ntrav, n_updates, n_visits, 2, TAG_MCCF_SEQ);
}

/* ================================================================================ exploitability */
/* Restated open_spiel best response (third-party, unpinned -> parity unpinned); mirrors
 * oracle/ms_exploit.py which documents the algorithm. */
typedef struct {
    ora_state st; int parent; int nchild; int child[4]; int legal[4];
    int info;            /* index into per-player infoset list (for the BR player) */
    double cf;           /* counterfactual reach for the BR player */
    double value; int has_value;
    double prob[4];      /* policy probs at this node */
} br_node;

typedef struct { char key[48]; int first; int n; int* members; int br; int has_br; } br_info;

typedef struct {
    br_node* nodes; int n, cap;
    br_info* infos; int ninfo;
    int b;
} br_ctx;

static void policy_probs(const ora_table* t, int kind, const ora_state* s, int n, const int* legal, double* out) {
    int player = ora_state_current_player(s);
    char info[64], key[72];
    ora_state_info_string(s, player, info, sizeof info);
    if (kind == 0) {                       /* LearnedCFRPolicy.action_probabilities (vanilla_cfr.py:128-144) */
        int i = ora_table_find(t, info);
        if (i >= 0) {
            const ora_node* nd = &t->nodes[i];
            double norm = 0.0;
            for (int a = 0; a < nd->nlegal; a++) norm += nd->strategy[a];
            for (int a = 0; a < n; a++) out[a] = norm > 0 ? nd->strategy[a] / norm : 1.0 / nd->nlegal;
            return;
        }
    } else if (kind == 1) {                /* ScopaLearnedPolicy.action_probabilities (mc_cfr.py:110-130) */
        snprintf(key, sizeof key, "%d|%s", player, info);
        int i = ora_table_find(t, key);
        if (i >= 0) {
            const ora_node* nd = &t->nodes[i];
            double tot = 0.0;
            for (int a = 0; a < nd->nlegal; a++) tot += nd->strategy[a];
            for (int a = 0; a < n; a++) out[a] = tot > 1e-12 ? nd->strategy[a] / tot : 1.0 / nd->nlegal;
            return;
        }
    }
    for (int a = 0; a < n; a++) out[a] = 1.0 / n;
}

static int br_build(br_ctx* c, const ora_table* t, int kind, const ora_state* s, int parent) {
    int id = c->n++;
    br_node* nd = &c->nodes[id];
    memset(nd, 0, sizeof(*nd));
    nd->st = *s; nd->parent = parent; nd->info = -1;
    if (s->is_terminal) return id;
    int n = ora_state_legal(s, -1, nd->legal);
    nd->nchild = n;
    policy_probs(t, kind, s, n, nd->legal, nd->prob);
    for (int i = 0; i < n; i++) {
        ora_state ch; ora_state_clone(s, &ch); ora_state_apply(&ch, c->nodes[id].legal[i]);
        int cid = br_build(c, t, kind, &ch, id);
        c->nodes[id].child[i] = cid;
    }
    return id;
}

static double br_value(br_ctx* c, int id);

static int br_action(br_ctx* c, int info) {
    br_info* I = &c->infos[info];
    if (I->has_br) return I->br;
    int nA = c->nodes[I->members[0]].nchild;
    int best = 0; double bestq = 0;
    for (int a = 0; a < nA; a++) {
        double q = 0.0;
        for (int m = 0; m < I->n; m++) {
            br_node* nd = &c->nodes[I->members[m]];
            q += nd->cf * br_value(c, nd->child[a]);
        }
        if (a == 0 || q > bestq) { best = a; bestq = q; }
    }
    I->br = best; I->has_br = 1;
    return best;
}

static double br_value(br_ctx* c, int id) {
    br_node* nd = &c->nodes[id];
    if (nd->has_value) return nd->value;
    double v;
    if (nd->st.is_terminal) { double r[2]; ora_state_rewards(&nd->st, r); v = r[c->b]; }
    else if (ora_state_current_player(&nd->st) == c->b) v = br_value(c, nd->child[br_action(c, nd->info)]);
    else {
        v = 0.0;
        for (int i = 0; i < nd->nchild; i++) if (nd->prob[i] > 0.0) v += nd->prob[i] * br_value(c, nd->child[i]);
    }
    nd->value = v; nd->has_value = 1;
    return v;
}

double ora_exploitability(const ora_table* t, int policy_kind, int64_t seed, double br_values[2]) {
    double total = 0.0;
    for (int b = 0; b < 2; b++) {
        br_ctx c; memset(&c, 0, sizeof c);
        c.cap = 4096; c.nodes = (br_node*)calloc(c.cap, sizeof(br_node)); c.b = b;
        ora_state root; ora_state_init(&root, seed);
        br_build(&c, t, policy_kind, &root, -1);
        /* counterfactual reach: product of the other player's policy probs, b's own edges weigh 1 */
        c.infos = (br_info*)calloc(c.n, sizeof(br_info));
        for (int i = 0; i < c.n; i++) {
            br_node* nd = &c.nodes[i];
            /* python multiplies from the node upward: ((1.0 * p_k) * p_{k-1}) ... */
            double cf = 1.0; int cur = i;
            while (c.nodes[cur].parent >= 0) {
                int par = c.nodes[cur].parent;
                br_node* pn = &c.nodes[par];
                int which = 0; for (int q = 0; q < pn->nchild; q++) if (pn->child[q] == cur) which = q;
                double p = ora_state_current_player(&pn->st) == b ? 1.0 : pn->prob[which];
                cf = cf * p;
                cur = par;
            }
            nd->cf = cf;
            if (!nd->st.is_terminal && ora_state_current_player(&nd->st) == b) {
                char key[64]; ora_state_info_string(&nd->st, b, key, sizeof key);
                int f = -1;
                for (int q = 0; q < c.ninfo; q++) if (strcmp(c.infos[q].key, key) == 0) { f = q; break; }
                if (f < 0) {
                    f = c.ninfo++;
                    strncpy(c.infos[f].key, key, sizeof(c.infos[f].key) - 1);
                    c.infos[f].members = (int*)malloc(sizeof(int) * 64); c.infos[f].n = 0;
                }
                if (c.infos[f].n < 64) c.infos[f].members[c.infos[f].n++] = i;
                else { fprintf(stderr, "infoset too large\n"); abort(); }
                nd->info = f;
            }
        }
        double v = br_value(&c, 0);
        if (br_values) br_values[b] = v;
        total += v;
        for (int q = 0; q < c.ninfo; q++) free(c.infos[q].members);
        free(c.infos); free(c.nodes);
    }
    return (total - 0.0) / 2;
}

/* ================================================================================ SDCFR */
/* DeepCFR._state_to_features (deep_cfr.py:213-275) + _get_legal_actions_mask (:277-282).  The
 * reference parses the info STRING back into one-hots by action id; the information is the
 * player's hand, the table (order dropped) and [player == current_player, 0.0]. */
void ora_features(const ora_state* s, int player, float feat[34], float mask[16]) {
    for (int i = 0; i < 34; i++) feat[i] = 0.f;
    for (int i = 0; i < 16; i++) mask[i] = 0.f;
    if (s->is_terminal || player < 0) return;      /* info "TERMINAL" has no H[/T[ -> zeros (:269-270) */
    for (int i = 0; i < s->env.nhand[player]; i++) feat[s->env.hand[player][i]] = 1.f;
    for (int i = 0; i < s->env.ntable; i++) feat[16 + s->env.table[i]] = 1.f;
    feat[32] = (player == ora_state_current_player(s)) ? 1.f : 0.f;
    feat[33] = 0.f;
    int legal[4]; int n = ora_state_legal(s, player, legal);
    for (int i = 0; i < n; i++) mask[legal[i]] = 1.f;
}

static void linear_relu(const float* w, const float* b, const float* x, int nin, int nout, int relu, float* y) {
    for (int o = 0; o < nout; o++) {
        float acc = b[o];
        for (int i = 0; i < nin; i++) acc += w[o * nin + i] * x[i];
        y[o] = (relu && acc < 0.f) ? 0.f : acc;
    }
}

void ora_mlp_forward(const ora_mlp* net, const float feat[34], float out[16]) {
    float h1[128], h2[64];
    linear_relu(net->w1, net->b1, feat, 34, 128, 1, h1);
    linear_relu(net->w2, net->b2, h1, 128, 64, 1, h2);
    linear_relu(net->w3, net->b3, h2, 64, 16, 0, out);
}

/* AdvantageNetwork.get_advantages (:54-68) + positive_regret_policy (nets.py:93-101) */
void ora_advantages_policy(const ora_mlp* net, const float feat[34], const float mask[16],
                           float adv[16], float pol[16]) {
    float raw[16];
    ora_mlp_forward(net, feat, raw);
    float z = 0.f;
    for (int i = 0; i < 16; i++) {
        adv[i] = raw[i] * mask[i] - 1e6f * (1.f - mask[i]);
        float pos = (adv[i] > 0.f ? adv[i] : 0.f) * mask[i];
        pol[i] = pos; z += pos;
    }
    if (z < 1e-8f) z = 1e-8f;
    for (int i = 0; i < 16; i++) pol[i] = pol[i] / z;
}

typedef struct {
    const ora_mlp* nets; ora_rng* rng; int player;
    float* out_feat; float* out_target; float* out_mask; int cap; int n;
} sd_ctx;

/* DeepCFR._external_sampling_cfr (:284-365); float32 arithmetic as numpy>=2 (NEP 50) evaluates it */
static float sdcfr_rec(sd_ctx* c, const ora_state* s) {
    uint32_t my_call = c->rng->call++;
    if (s->is_terminal) { double r[2]; ora_state_rewards(s, r); return (float)r[c->player]; }
    int cp = ora_state_current_player(s);
    float feat[34], mask[16], adv[16], pol[16];
    ora_features(s, cp, feat, mask);
    ora_advantages_policy(&c->nets[cp], feat, mask, adv, pol);
    int legal[4]; int n = ora_state_legal(s, cp, legal);
    if (cp == c->player) {
        float value = 0.f; float cfv[16];
        for (int i = 0; i < 16; i++) cfv[i] = 0.f;
        for (int i = 0; i < n; i++) {
            ora_state ch; ora_state_clone(s, &ch); ora_state_apply(&ch, legal[i]);
            float av = sdcfr_rec(c, &ch);
            value += pol[legal[i]] * av;
            cfv[legal[i]] = av;
        }
        float reg[16]; float mx = 0.f;
        for (int i = 0; i < 16; i++) { reg[i] = cfv[i] - value; float a = fabsf(reg[i]); if (a > mx) mx = a; }
        if (mx > 0.f) { float d = mx + 1e-8f; for (int i = 0; i < 16; i++) reg[i] = reg[i] / d; }  /* :70-75 */
        if (c->n < c->cap) {
            memcpy(c->out_feat + (size_t)c->n * 34, feat, sizeof feat);
            memcpy(c->out_target + (size_t)c->n * 16, reg, sizeof reg);
            memcpy(c->out_mask + (size_t)c->n * 16, mask, sizeof mask);
        }
        c->n++;
        return value;
    }
    /* opponent: sample one action (:347-365) */
    float ap[4]; float sum = 0.f;
    for (int i = 0; i < n; i++) { ap[i] = pol[legal[i]]; }
    for (int i = 0; i < n; i++) sum += ap[i];
    int ai;
    uint32_t saved = c->rng->call; c->rng->call = my_call;
    if (sum == 0.f) ai = rng_choice_uniform(c->rng, n);
    else {
        double p[4];
        for (int i = 0; i < n; i++) p[i] = (double)(ap[i] / sum);   /* float32 divide, then float64 cumsum */
        ai = rng_choice_p(c->rng, p, n);
    }
    c->rng->call = saved;
    ora_state nx; ora_state_clone(s, &nx); ora_state_apply(&nx, legal[ai]);
    return sdcfr_rec(c, &nx);
}

float ora_sdcfr_traverse(const ora_mlp nets[2], int64_t seed, int player, ora_rng* rng, uint64_t trav_id,
                         float* out_feat, float* out_target, float* out_mask, int cap, int* n_out) {
    sd_ctx c = {nets, rng, player, out_feat, out_target, out_mask, cap, 0};
    rng->tag = TAG_SDCF + (uint32_t)player; rng->trav = trav_id; rng->call = 0;
    ora_state s; ora_state_init(&s, seed);
    float v = sdcfr_rec(&c, &s);
    if (n_out) *n_out = c.n;
    return v;
}

/* ================================================================================ CPU baseline */
/* bench.py's cpu_baseline / --impl reference leg: `nthreads` independent workers, each with its own
 * pre-populated table, each running ntrav_per_thread traversal pairs (player 0 then player 1 against
 * the frozen table) of the reference estimator.  Returns totals over all workers. */
void ora_mccfr_bench(int64_t seed, int64_t ntrav_per_thread, int nthreads, uint64_t philox_seed,
                     int64_t* updates, int64_t* visits) {
    int64_t tu = 0, tv = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : tu, tv)
    for (int w = 0; w < nthreads; w++) {
        ora_table* t = ora_table_new();
        ora_mccfr_populate(t, seed);
        int64_t u = 0, v = 0;
        for (int p = 0; p < 2; p++) {
            int64_t uu = 0, vv = 0;
            ora_mccfr_batch(t, seed, p, philox_seed, (uint64_t)w * (uint64_t)ntrav_per_thread, ntrav_per_thread, &uu, &vv);
            u += uu; v += vv;
        }
        tu += u; tv += v;
        ora_table_free(t);
    }
    if (updates) *updates = tu;
    if (visits) *visits = tv;
}

/* ================================================================================ textbook MCCFR */
/* Opt-in estimators that the reference does NOT implement (its MCCFRTrainer is the hybrid estimator above):
 * external sampling and outcome sampling as published (Lanctot et al. 2009; same update rules as
 * open_spiel/python/algorithms/external_sampling_mccfr.py with AverageType.SIMPLE and
 * outcome_sampling_mccfr.py with epsilon = 0.6, baseline 0).  Batch semantics only: sigma frozen for the batch,
 * deltas applied at the end.  Philox "MCCF" stream, call index = order of recursive invocations. */
#define OS_EPSILON 0.6

static double es_rec(mccfr_ctx* c, const ora_state* s) {
    uint32_t my_call = c->rng->call++;
    c->n_visits++;
    if (s->is_terminal) { double r[2]; ora_state_rewards(s, r); return r[c->tp]; }
    int player = ora_state_current_player(s);
    char key[72]; mccfr_key(s, player, key, sizeof key);
    int legal[4]; int n = ora_state_legal(s, player, legal);
    ora_node* nd = table_get(c->t, key, legal, n);
    int slot = (int)(nd - c->t->nodes);
    double sigma[4];
    regret_matching(nd->regret, n, sigma);
    if (player != c->tp) {                 /* opponent: sample one action, average strategy updated here */
        uint32_t saved = c->rng->call; c->rng->call = my_call;
        int ai = rng_choice_p(c->rng, sigma, n);
        c->rng->call = saved;
        for (int i = 0; i < n; i++) c->dstr[slot][i] += sigma[i];
        ora_state nx; ora_state_clone(s, &nx); ora_state_apply(&nx, legal[ai]);
        return es_rec(c, &nx);
    }
    double cv[4], value = 0.0;
    for (int i = 0; i < n; i++) {
        ora_state nx; ora_state_clone(s, &nx); ora_state_apply(&nx, legal[i]);
        cv[i] = es_rec(c, &nx);
        value += sigma[i] * cv[i];
    }
    for (int i = 0; i < n; i++) c->dreg[slot][i] += cv[i] - value;
    c->n_updates++;
    return value;
}

static double os_rec(mccfr_ctx* c, const ora_state* s, double my_reach, double opp_reach, double sample_reach) {
    uint32_t my_call = c->rng->call++;
    c->n_visits++;
    if (s->is_terminal) { double r[2]; ora_state_rewards(s, r); return r[c->tp]; }
    int player = ora_state_current_player(s);
    char key[72]; mccfr_key(s, player, key, sizeof key);
    int legal[4]; int n = ora_state_legal(s, player, legal);
    ora_node* nd = table_get(c->t, key, legal, n);
    int slot = (int)(nd - c->t->nodes);
    double sigma[4], sp[4];
    regret_matching(nd->regret, n, sigma);
    for (int i = 0; i < n; i++)
        sp[i] = (player == c->tp) ? OS_EPSILON * (1.0 / n) + (1.0 - OS_EPSILON) * sigma[i] : sigma[i];
    uint32_t saved = c->rng->call; c->rng->call = my_call;
    int ai = rng_choice_p(c->rng, sp, n);
    c->rng->call = saved;
    double nmy = my_reach, nopp = opp_reach;
    if (player == c->tp) nmy = my_reach * sigma[ai]; else nopp = opp_reach * sigma[ai];
    ora_state nx; ora_state_clone(s, &nx); ora_state_apply(&nx, legal[ai]);
    double child = os_rec(c, &nx, nmy, nopp, sample_reach * sp[ai]);
    double est_a = child / sp[ai];                       /* estimated value of the sampled action, 0 for the others */
    double value_estimate = sigma[ai] * est_a;
    if (player == c->tp) {
        double w = opp_reach / sample_reach;
        double cf_value = value_estimate * w;
        for (int i = 0; i < n; i++) {
            double cf_action = (i == ai ? est_a : 0.0) * w;
            c->dreg[slot][i] += cf_action - cf_value;
            c->dstr[slot][i] += my_reach * sigma[i] / sample_reach;
        }
        c->n_updates++;
    }
    return value_estimate;
}

/* mode 1 = external sampling, 2 = outcome sampling; same calling convention as ora_mccfr_batch */
void ora_mccfr_batch_mode(ora_table* t, int64_t seed, int mode, int player, uint64_t philox_seed,
                          uint64_t first_trav, int64_t ntrav, int64_t* n_updates, int64_t* n_visits) {
    double (*dreg)[4] = calloc((size_t)t->cap, sizeof(double[4]));
    double (*dstr)[4] = calloc((size_t)t->cap, sizeof(double[4]));
    ora_rng* rng = ora_rng_new(1, philox_seed);
    mccfr_ctx c = {t, rng, player, dreg, dstr, 0, 0};
    int n0 = t->n;
    for (int64_t k = 0; k < ntrav; k++) {
        ora_state s; ora_state_init(&s, seed);
        rng->tag = TAG_MCCF + (uint32_t)player + 16u * (uint32_t)mode; rng->trav = first_trav + (uint64_t)k; rng->call = 0;
        if (mode == 1) es_rec(&c, &s); else os_rec(&c, &s, 1.0, 1.0, 1.0);
    }
    if (t->n != n0) { fprintf(stderr, "ora_mccfr_batch_mode: table was not pre-populated\n"); abort(); }
    for (int i = 0; i < t->n; i++)
        for (int a = 0; a < 4; a++) { t->nodes[i].regret[a] += dreg[i][a]; t->nodes[i].strategy[a] += dstr[i][a]; }
    if (n_updates) *n_updates = c.n_updates;
    if (n_visits) *n_visits = c.n_visits;
    ora_rng_free(rng); free(dreg); free(dstr);
}

/* ================================================================================ team Miniscopa (2v2) */
/* src/envs/team_mini_scopa_game.py:44-243, list-based like the rest of this file. */
void ora_team_reset(ora_team_env* e, int64_t seed, int has_seed) {
    int64_t s = (has_seed && seed != 0) ? seed : e->seed;        /* `seed or self.seed` (:168) */
    int deck[16];
    ora_deck(s, deck);
    memset(e->ncaps, 0, sizeof e->ncaps);
    memset(e->scopas, 0, sizeof e->scopas);
    e->ntable = 0; e->last_capture_team = -1;
    for (int p = 0; p < 4; p++) {
        e->nhand[p] = 4;
        for (int i = 0; i < 4; i++) e->hand[p][i] = deck[4 * p + i];
        e->rewards[p] = 0; e->term[p] = 0;
    }
    e->agent = 0; e->step_count = 0;
}

void ora_team_init(ora_team_env* e, int64_t seed) {
    memset(e, 0, sizeof *e);
    e->max_steps = 16; e->seed = seed;
    ora_team_reset(e, seed, 1);
}

static void team_evaluate(ora_team_env* e, double out[4]) {       /* evaluate_game (:118-148) */
    if (e->ntable > 0 && e->last_capture_team >= 0) {
        int p = 2 * e->last_capture_team;                         /* first player of that team */
        for (int i = 0; i < e->ntable; i++) e->caps[p][e->ncaps[p]++] = e->table[i];
    }
    int sc[2] = {0, 0};
    for (int p = 0; p < 4; p++) sc[p / 2] += e->ncaps[p] + 2 * e->scopas[p];
    int total = sc[0] + sc[1];
    if (total == 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    double mean = (double)total / 2;
    out[0] = out[1] = sc[0] - mean; out[2] = out[3] = sc[1] - mean;
}

void ora_team_step(ora_team_env* e, int action) {                 /* step (:171-205) + play_card (:101-116) */
    if (e->term[e->agent]) return;
    int pl = e->agent, hp = -1;
    if (action >= 0 && action < 16)
        for (int i = 0; i < e->nhand[pl]; i++) if (e->hand[pl][i] == action) { hp = i; break; }
    if (hp >= 0) {
        int card = e->hand[pl][hp], pos[8];
        int ncap = ora_card_in_table(e->table, e->ntable, card, pos);
        if (ncap > 0) {
            int keep[16], nk = 0;
            for (int q = 0; q < ncap; q++) e->caps[pl][e->ncaps[pl]++] = e->table[pos[q]];
            e->caps[pl][e->ncaps[pl]++] = card;
            for (int i = 0; i < e->ntable; i++) {
                int gone = 0;
                for (int q = 0; q < ncap; q++) if (pos[q] == i) gone = 1;
                if (!gone) keep[nk++] = e->table[i];
            }
            for (int i = 0; i < nk; i++) e->table[i] = keep[i];
            e->ntable = nk;
            e->last_capture_team = pl / 2;
            if (nk == 0) e->scopas[pl] += 1;
        } else {
            e->table[e->ntable++] = card;
        }
        for (int i = hp; i + 1 < e->nhand[pl]; i++) e->hand[pl][i] = e->hand[pl][i + 1];
        e->nhand[pl]--;
    }
    e->step_count += 1;
    int empty = 1;
    for (int p = 0; p < 4; p++) if (e->nhand[p]) empty = 0;
    if (empty || e->step_count >= e->max_steps) {
        double r[4]; team_evaluate(e, r);
        for (int p = 0; p < 4; p++) { e->rewards[p] = r[p]; e->term[p] = 1; }
    }
    e->agent = (pl + 1) % 4;
}

#define TAG_TEAM 0x4D414554u
void ora_team_rollout_random(const int64_t* seeds, int64_t n, uint64_t philox_seed, uint64_t game_offset,
                             uint8_t* actions, float* rewards, uint8_t* scopas, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < n; g++) {
        ora_team_env e;
        ora_team_init(&e, 42);
        ora_team_reset(&e, seeds[g], 1);
        uint32_t key[2] = {(uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)};
        for (int ply = 0; ply < 16; ply++) {
            uint64_t gid = game_offset + (uint64_t)g;
            uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(ply >> 2), TAG_TEAM}, o[4];
            philox4x32_10(ctr, key, o);
            int pl = e.agent, nl = e.term[pl] ? 0 : e.nhand[pl];
            int a = nl > 0 ? e.hand[pl][(int)(((uint64_t)o[ply & 3] * (uint64_t)nl) >> 32)] : 0;
            actions[g * 16 + ply] = (uint8_t)a;
            ora_team_step(&e, a);
        }
        for (int p = 0; p < 4; p++) { rewards[g * 4 + p] = (float)e.rewards[p]; if (scopas) scopas[g * 4 + p] = (uint8_t)e.scopas[p]; }
    }
}

/* ================================================================================ multi-deal MCCFR */
/* No reference solver plays more than the seed-42 deal, so this section is "parity unpinned" beyond its
 * D = 1 case: it is MCCFRTrainer._sample (src/algorithms/mc_cfr.py:37-86) run from a deal drawn uniformly from
 * a list of seeds (Philox "DEAL" stream), with the info_sets dict (:28-35) keyed by the information itself
 * -- "player|hand ids ascending|table ids in order" -- instead of the string whose hand order depends on the
 * deal.  regret / strategy arrays are indexed like the reference's (by card id), compacted to the cards in
 * hand.  Batch semantics as in ora_mccfr_batch: strategies frozen while a batch runs. */
#define TAG_DEAL 0x4C414544u /* "DEAL" */

typedef struct {
    char key[64];
    int player, nh, hand[4], nt, table[8];
    double regret[4], strategy[4], dreg[4];
    int64_t dcnt;
} md_node;

struct ora_mdtable {
    md_node* nodes; int64_t n, cap;
    int64_t* hidx; int64_t hcap;
};

ora_mdtable* ora_md_new(void) {
    ora_mdtable* t = (ora_mdtable*)calloc(1, sizeof(*t));
    t->cap = 1024; t->nodes = (md_node*)calloc((size_t)t->cap, sizeof(md_node));
    t->hcap = 4096; t->hidx = (int64_t*)malloc(sizeof(int64_t) * (size_t)t->hcap);
    for (int64_t i = 0; i < t->hcap; i++) t->hidx[i] = -1;
    return t;
}
void ora_md_free(ora_mdtable* t) { if (t) { free(t->nodes); free(t->hidx); free(t); } }
int64_t ora_md_size(const ora_mdtable* t) { return t->n; }
const char* ora_md_key(const ora_mdtable* t, int64_t i) { return t->nodes[i].key; }
int ora_md_nlegal(const ora_mdtable* t, int64_t i) { return t->nodes[i].nh; }
double* ora_md_regret(ora_mdtable* t, int64_t i) { return t->nodes[i].regret; }
double* ora_md_strategy(ora_mdtable* t, int64_t i) { return t->nodes[i].strategy; }
/* the same information in the packed form the CUDA library uses for its keys (export format only) */
uint64_t ora_md_packed_key(const ora_mdtable* t, int64_t i) {
    const md_node* nd = &t->nodes[i];
    uint64_t hand = 0, table = 0;
    for (int k = 0; k < nd->nh; k++) hand |= 1ull << nd->hand[k];
    for (int k = 0; k < nd->nt; k++) table |= (uint64_t)nd->table[k] << (4 * k);
    return ((uint64_t)nd->player << 52) | (hand << 36) | ((uint64_t)nd->nt << 32) | table;
}

static void md_rehash(ora_mdtable* t) {
    free(t->hidx);
    t->hcap *= 2;
    t->hidx = (int64_t*)malloc(sizeof(int64_t) * (size_t)t->hcap);
    for (int64_t i = 0; i < t->hcap; i++) t->hidx[i] = -1;
    for (int64_t i = 0; i < t->n; i++) {
        uint32_t h = fnv1a(t->nodes[i].key) & (uint32_t)(t->hcap - 1);
        while (t->hidx[h] >= 0) h = (h + 1) & (uint32_t)(t->hcap - 1);
        t->hidx[h] = i;
    }
}

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

static int64_t md_get(ora_mdtable* t, const ora_state* s, int player) {
    int hand[4]; int nh = s->env.nhand[player];
    for (int k = 0; k < nh; k++) hand[k] = s->env.hand[player][k];
    qsort(hand, (size_t)nh, sizeof(int), cmp_int);
    char key[64]; int o = snprintf(key, sizeof key, "%d|", player);
    for (int k = 0; k < nh; k++) o += snprintf(key + o, sizeof key - (size_t)o, k ? ",%d" : "%d", hand[k]);
    o += snprintf(key + o, sizeof key - (size_t)o, "|");
    for (int k = 0; k < s->env.ntable; k++) o += snprintf(key + o, sizeof key - (size_t)o, k ? ",%d" : "%d", s->env.table[k]);
    uint32_t h = fnv1a(key) & (uint32_t)(t->hcap - 1);
    while (t->hidx[h] >= 0) {
        if (strcmp(t->nodes[t->hidx[h]].key, key) == 0) return t->hidx[h];
        h = (h + 1) & (uint32_t)(t->hcap - 1);
    }
    if (t->n >= t->cap) {
        t->cap *= 2;
        t->nodes = (md_node*)realloc(t->nodes, (size_t)t->cap * sizeof(md_node));
    }
    md_node* nd = &t->nodes[t->n];
    memset(nd, 0, sizeof(*nd));
    strcpy(nd->key, key);
    nd->player = player; nd->nh = nh; nd->nt = s->env.ntable;
    for (int k = 0; k < nh; k++) nd->hand[k] = hand[k];
    for (int k = 0; k < nd->nt; k++) nd->table[k] = s->env.table[k];
    t->hidx[h] = t->n++;
    if (t->n * 2 > t->hcap) md_rehash(t);
    return t->n - 1;
}

typedef struct { ora_mdtable* t; ora_rng* rng; int tp; int64_t n_updates, n_visits; } md_ctx;

/* column of a card inside a node: its rank among the node's (ascending) hand ids */
static int md_column(const md_node* nd, int card) {
    for (int k = 0; k < nd->nh; k++) if (nd->hand[k] == card) return k;
    fprintf(stderr, "md_column: card not in hand\n"); abort();
}

static double md_sample(md_ctx* c, const ora_state* s, const double reach[2], const double samp[2]) {
    uint32_t my_call = c->rng->call++;
    c->n_visits++;
    if (s->is_terminal) { double r[2]; ora_state_rewards(s, r); return r[c->tp]; }
    int player = ora_state_current_player(s);
    int legal[4]; int n = ora_state_legal(s, player, legal);      /* hand (= deal) order, like the reference */
    /* infosets with a single legal action are not stored: sigma = [1.0], regret stays 0, and the reference's
     * strategy_sum there is its visit count, which no consumer reads */
    int64_t slot = n > 1 ? md_get(c->t, s, player) : -1;
    int col[4] = {0, 0, 0, 0}; double by_col[4], sigma[4] = {1.0, 0.0, 0.0, 0.0};
    if (slot >= 0) {
        regret_matching(c->t->nodes[slot].regret, n, by_col);     /* :54, over the table's columns */
        for (int i = 0; i < n; i++) { col[i] = md_column(&c->t->nodes[slot], legal[i]); sigma[i] = by_col[col[i]]; }
    }
    uint32_t saved = c->rng->call; c->rng->call = my_call;
    int ai = rng_choice_p(c->rng, sigma, n);                      /* :55 */
    c->rng->call = saved;
    ora_state nx; ora_state_clone(s, &nx); ora_state_apply(&nx, legal[ai]);
    double nreach[2] = {reach[0], reach[1]}, nsamp[2] = {samp[0], samp[1]};
    if (player == c->tp) nsamp[player] *= sigma[ai];
    else { nreach[player] *= sigma[ai]; nsamp[player] *= sigma[ai]; }
    double util = md_sample(c, &nx, nreach, nsamp);               /* :67 */
    if (player == c->tp) {
        double cfv[4];
        for (int i = 0; i < n; i++) {                             /* :71-78 */
            ora_state tmp; ora_state_clone(s, &tmp); ora_state_apply(&tmp, legal[i]);
            double tsamp[2] = {samp[0], samp[1]};
            tsamp[player] *= sigma[i];
            cfv[i] = md_sample(c, &tmp, reach, tsamp);
        }
        double v = 0.0;
        for (int i = 0; i < n; i++) v += sigma[i] * cfv[i];
        double opp = reach[1 - player];
        double w = samp[player] > 0 ? opp / samp[player] : 0.0;
        if (slot >= 0) {
            md_node* nd = &c->t->nodes[slot];                     /* the array may have moved: re-index */
            for (int i = 0; i < n; i++) nd->dreg[col[i]] += w * (cfv[i] - v);
            nd->dcnt += 1;                                        /* strategy delta = reach[player] (= 1.0) * sigma */
        }
        c->n_updates++;
    }
    return util;
}

void ora_md_batch(ora_mdtable* t, const int64_t* seeds, int64_t n_deals, int player, uint64_t philox_seed,
                  uint64_t first_trav, int64_t ntrav, int64_t* n_updates, int64_t* n_visits) {
    ora_rng* rng = ora_rng_new(1, philox_seed);
    md_ctx c = {t, rng, 0, 0, 0};
    const uint32_t key[2] = {(uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)};
    for (int64_t k = 0; k < ntrav; k++) {
        uint64_t trav = first_trav + (uint64_t)k;
        uint32_t ctr[4] = {(uint32_t)trav, (uint32_t)(trav >> 32), 0u, TAG_DEAL}, o[4];
        philox4x32_10(ctr, key, o);
        int64_t deal = (int64_t)(((uint64_t)o[0] * (uint64_t)n_deals) >> 32);
        for (int tp = 0; tp < 2; tp++) {
            if (player < 2 && tp != player) continue;
            ora_state s; ora_state_init(&s, seeds[deal]);
            c.tp = tp;
            rng->tag = TAG_MCCF + (uint32_t)tp; rng->trav = trav; rng->call = 0;
            double one[2] = {1.0, 1.0};
            md_sample(&c, &s, one, one);
        }
    }
    if (n_updates) *n_updates = c.n_updates;
    if (n_visits) *n_visits = c.n_visits;
    ora_rng_free(rng);
}

void ora_md_apply(ora_mdtable* t) {
    for (int64_t i = 0; i < t->n; i++) {
        md_node* nd = &t->nodes[i];
        if (nd->dcnt == 0) continue;
        double sigma[4];
        regret_matching(nd->regret, nd->nh, sigma);
        for (int a = 0; a < nd->nh; a++) {
            nd->strategy[a] += (double)nd->dcnt * sigma[a];
            nd->regret[a] += nd->dreg[a];
            nd->dreg[a] = 0.0;
        }
        nd->dcnt = 0;
    }
}

/* ================================================================================ 40-card Scopa */
/* FullDeck / FullScopaGame / FullScopaEnv (src/envs/full_scopa_game.py:21-342), two players.
 * Card id = suit_idx * 10 + (rank - 1), suits = denari, coppe, spade, bastoni (:23-24, :262-266). */
#define TAG_FULL 0x4C4C5546u /* "FULL" */

static int full_rank(int c) { return c % 10 + 1; }

void ora_full_deck(int64_t seed, int out40[40]) {           /* FullDeck.__init__ (:32-35) */
    mt_t m; py_random_seed(&m, seed);
    for (int i = 0; i < 40; i++) out40[i] = i;
    for (int i = 39; i >= 1; i--) {                          /* random.shuffle */
        int j = (int)py_randbelow(&m, (uint32_t)i + 1u);
        int t = out40[i]; out40[i] = out40[j]; out40[j] = t;
    }
}

static void full_deal_hands(ora_full_env* e) {               /* 3 cards to each player, player 0 first */
    for (int p = 0; p < 2; p++) {
        e->nhand[p] = 3;
        for (int i = 0; i < 3; i++) e->hand[p][i] = e->deck[e->deck_pos++];
    }
}

void ora_full_reset(ora_full_env* e, int64_t seed, int has_seed) {   /* FullScopaEnv.reset (:243-250) + game.reset (:68-86) */
    int64_t sd = (has_seed && seed != 0) ? seed : e->seed;   /* `seed or self.seed` */
    ora_full_deck(sd, e->deck);
    e->deck_pos = 0;
    e->ntable = 4;
    for (int i = 0; i < 4; i++) e->table[i] = e->deck[e->deck_pos++];
    for (int p = 0; p < 2; p++) { e->ncaps[p] = 0; e->scopas[p] = 0; }
    full_deal_hands(e);
    e->last_capture = -1; e->round_number = 0;
    e->agent = 0; e->step_count = 0;
    e->rewards[0] = e->rewards[1] = 0.0; e->term[0] = e->term[1] = 0;
}

void ora_full_init(ora_full_env* e, int64_t seed) {          /* FullScopaEnv.__init__ (:231-241) */
    memset(e, 0, sizeof(*e));
    e->seed = seed; e->max_steps = 200;
    ora_full_reset(e, seed, 1);
}

/* find_capture_combinations (:101-128), first combination only (play_card takes combinations[0], :137-141):
 * the first table card of equal rank, else the subset with the smallest position mask whose ranks sum to the
 * target.  Returns the number of captured table positions written to pos_out (ascending). */
int ora_full_capture(const int* table, int ntable, int card, int* pos_out) {
    int target = full_rank(card);
    for (int i = 0; i < ntable; i++)
        if (full_rank(table[i]) == target) { pos_out[0] = i; return 1; }
    for (uint32_t mask = 1; mask < (1u << ntable); mask++) {
        int sum = 0;
        for (int i = 0; i < ntable; i++) if (mask & (1u << i)) sum += full_rank(table[i]);
        if (sum == target) {
            int n = 0;
            for (int i = 0; i < ntable; i++) if (mask & (1u << i)) pos_out[n++] = i;
            return n;
        }
    }
    return 0;
}

static int primiera_value(int rank) {                        /* :27-30 */
    static const int v[11] = {0, 16, 12, 13, 14, 15, 18, 21, 10, 10, 10};
    return v[rank];
}

static int full_primiera(const int* caps, int n) {           /* calculate_primiera_score (:160-172) */
    int best[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        int s = caps[i] / 10, v = primiera_value(full_rank(caps[i]));
        if (v > best[s]) best[s] = v;
    }
    for (int s = 0; s < 4; s++) if (!best[s]) return 0;
    return best[0] + best[1] + best[2] + best[3];
}

static void full_evaluate(ora_full_env* e, double out[2]) {  /* evaluate_game (:174-226) */
    int scores[2] = {0, 0};
    if (e->ntable > 0 && e->last_capture >= 0)               /* sweep: the table is NOT cleared (:187-188) */
        for (int i = 0; i < e->ntable; i++) e->caps[e->last_capture][e->ncaps[e->last_capture]++] = e->table[i];
    if (e->ncaps[0] != e->ncaps[1]) scores[e->ncaps[0] > e->ncaps[1] ? 0 : 1] += 1;          /* carte */
    int den[2] = {0, 0}, sette = -1, prim[2];
    for (int p = 0; p < 2; p++) {
        for (int i = 0; i < e->ncaps[p]; i++) {
            if (e->caps[p][i] < 10) den[p]++;
            if (e->caps[p][i] == 6 && sette < 0) sette = p;
        }
        prim[p] = full_primiera(e->caps[p], e->ncaps[p]);
    }
    if (den[0] != den[1]) scores[den[0] > den[1] ? 0 : 1] += 1;                              /* denari */
    if (sette >= 0) scores[sette] += 1;                                                      /* sette bello */
    if ((prim[0] > 0 || prim[1] > 0) && prim[0] != prim[1]) scores[prim[0] > prim[1] ? 0 : 1] += 1;
    scores[0] += e->scopas[0]; scores[1] += e->scopas[1];
    int total = scores[0] + scores[1];
    if (total == 0) { out[0] = out[1] = 0.0; return; }
    double mean = total / 2.0;
    out[0] = scores[0] - mean; out[1] = scores[1] - mean;
}

void ora_full_step(ora_full_env* e, int action) {            /* FullScopaEnv.step (:252-296) */
    if (e->term[e->agent]) return;                           /* dead step */
    int pl = e->agent, hp = -1;
    for (int i = 0; i < e->nhand[pl]; i++) if (e->hand[pl][i] == action) { hp = i; break; }
    if (hp >= 0 && action >= 0 && action < 40) {             /* play_card (:130-158) */
        int pos[20];
        int n = ora_full_capture(e->table, e->ntable, action, pos);
        if (n > 0) {
            for (int i = 0; i < n; i++) e->caps[pl][e->ncaps[pl]++] = e->table[pos[i]];
            e->caps[pl][e->ncaps[pl]++] = action;
            for (int i = n - 1; i >= 0; i--) {
                for (int k = pos[i]; k + 1 < e->ntable; k++) e->table[k] = e->table[k + 1];
                e->ntable--;
            }
            e->last_capture = pl;
            if (e->ntable == 0) e->scopas[pl]++;
        } else {
            e->table[e->ntable++] = action;
        }
        for (int k = hp; k + 1 < e->nhand[pl]; k++) e->hand[pl][k] = e->hand[pl][k + 1];
        e->nhand[pl]--;
    }
    e->step_count++;
    if (e->nhand[0] == 0 && e->nhand[1] == 0) {
        if (40 - e->deck_pos >= 6) { full_deal_hands(e); e->round_number++; }
        else { full_evaluate(e, e->rewards); e->term[0] = e->term[1] = 1; }
    }
    if (e->step_count >= e->max_steps) { full_evaluate(e, e->rewards); e->term[0] = e->term[1] = 1; }
    e->agent = 1 - e->agent;
}

/* random-policy games on the "FULL" Philox stream: ctr = (game id lo, hi, ply / 4, tag), word ply % 4,
 * action = hand[mulhi32(x, |hand|)]; 36 plies.  actions [n][36] u8, rewards [n][2] f32, scopas [n][2] u8,
 * ncaps [n][2] u8, maxtable [n] u8 (longest table seen). */
void ora_full_rollout_random(const int64_t* seeds, int64_t n, uint64_t philox_seed, uint64_t game_offset,
                             uint8_t* actions, float* rewards, uint8_t* scopas, uint8_t* ncaps, uint8_t* maxtable,
                             int nthreads) {
    const uint32_t key[2] = {(uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)};
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t g = 0; g < n; g++) {
        ora_full_env e; ora_full_init(&e, 42);
        ora_full_reset(&e, seeds[g], 1);
        int mt = e.ntable;
        for (int ply = 0; ply < 36; ply++) {
            uint64_t gid = game_offset + (uint64_t)g;
            uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(ply >> 2), TAG_FULL}, o[4];
            philox4x32_10(ctr, key, o);
            int pl = e.agent, nl = e.nhand[pl];
            int a = nl > 0 ? e.hand[pl][(int)(((uint64_t)o[ply & 3] * (uint64_t)nl) >> 32)] : 0;
            actions[g * 36 + ply] = (uint8_t)a;
            ora_full_step(&e, a);
            if (e.ntable > mt) mt = e.ntable;
        }
        rewards[g * 2] = (float)e.rewards[0]; rewards[g * 2 + 1] = (float)e.rewards[1];
        if (scopas) { scopas[g * 2] = (uint8_t)e.scopas[0]; scopas[g * 2 + 1] = (uint8_t)e.scopas[1]; }
        if (ncaps) { ncaps[g * 2] = (uint8_t)e.ncaps[0]; ncaps[g * 2 + 1] = (uint8_t)e.ncaps[1]; }
        if (maxtable) maxtable[g] = (uint8_t)mt;
    }
}

/* deal-blocked form of the multi-deal estimator: a visit draws ONE deal (Philox "DEAL" stream, counter word 2 = 1)
 * and runs `pairs` traversals on it, global ids visit * pairs + i; every deal's infosets with more than one action
 * exist in the table from the start (ora_md_populate). */
static void md_populate_rec(ora_mdtable* t, const ora_state* s) {
    if (s->is_terminal) return;
    int player = ora_state_current_player(s);
    int legal[4]; int n = ora_state_legal(s, player, legal);
    if (n > 1) md_get(t, s, player);
    for (int i = 0; i < n; i++) {
        ora_state c; ora_state_clone(s, &c); ora_state_apply(&c, legal[i]);
        md_populate_rec(t, &c);
    }
}
void ora_md_populate(ora_mdtable* t, const int64_t* seeds, int64_t n_deals) {
    for (int64_t d = 0; d < n_deals; d++) { ora_state s; ora_state_init(&s, seeds[d]); md_populate_rec(t, &s); }
}

void ora_md_batch_blocked(ora_mdtable* t, const int64_t* seeds, int64_t n_deals, int player, uint64_t philox_seed,
                          uint64_t first_visit, int64_t n_visits, int64_t pairs, int64_t* n_updates, int64_t* n_visits_out) {
    ora_rng* rng = ora_rng_new(2, philox_seed);               /* the sequential stream, like ora_mccfr_batch_seq */
    md_ctx c = {t, rng, 0, 0, 0};
    const uint32_t key[2] = {(uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)};
    for (int64_t v = 0; v < n_visits; v++) {
        uint64_t visit = first_visit + (uint64_t)v;
        uint32_t ctr[4] = {(uint32_t)visit, (uint32_t)(visit >> 32), 1u, TAG_DEAL}, o[4];
        philox4x32_10(ctr, key, o);
        int64_t deal = (int64_t)(((uint64_t)o[0] * (uint64_t)n_deals) >> 32);
        for (int64_t i = 0; i < pairs; i++) {
            uint64_t trav = visit * (uint64_t)pairs + (uint64_t)i;
            for (int tp = 0; tp < 2; tp++) {
                if (player < 2 && tp != player) continue;
                ora_state s; ora_state_init(&s, seeds[deal]);
                c.tp = tp;
                rng->tag = TAG_MCCF_SEQ + (uint32_t)tp; rng->trav = trav; rng->call = 0; rng->draw = 0;
                double one[2] = {1.0, 1.0};
                md_sample(&c, &s, one, one);
            }
        }
    }
    if (n_updates) *n_updates = c.n_updates;
    if (n_visits_out) *n_visits_out = c.n_visits;
    ora_rng_free(rng);
}
