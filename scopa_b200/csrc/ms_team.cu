// scopa_b200/csrc/ms_team.cu -- 2v2 team Miniscopa on the device (a "next" row): four players,
// all 16 cards dealt, 16 plies, teams {0,1} vs {2,3}; same capture rule as the 1v1 game plus the last-capturer
// sweep and team scoring.  Replaces /root/reference/src/envs/team_mini_scopa_game.py:44-243
// (TeamMiniScopaGame.reset / card_in_table / play_card / evaluate_game, TeamMiniScopaEnv.reset / step).
//
// Packed state, 32 bytes (two 128-bit words per lane):
//   w0: hand[0] | hand[1] << 16      w1: hand[2] | hand[3] << 16      (bit = card id)
//   w2: ordered table, 8 nibbles (a card is only ever placed when no equal-rank card lies on the table, so the
//       table holds distinct ranks: never more than 8 cards)
//   w3: table_len (4) | step_count (5) << 4 | current player (2) << 9 | terminal << 11 |
//       last_capture_team + 1 (2) << 12 | max_steps (5) << 14
//   w4: captures[0] | captures[1] << 16   w5: captures[2] | captures[3] << 16
//   w6: scopas, 4 bits per player         w7: unused
#include "ms_common.cuh"
#include "ms_state.cuh"

namespace ms {

struct __align__(16) TeamState { uint32_t w[8]; };

__device__ __forceinline__ uint32_t tm_hand(const TeamState& s, int p) { return (s.w[p >> 1] >> (16 * (p & 1))) & 0xFFFFu; }
__device__ __forceinline__ uint32_t tm_caps(const TeamState& s, int p) { return (s.w[4 + (p >> 1)] >> (16 * (p & 1))) & 0xFFFFu; }
__device__ __forceinline__ uint32_t tm_len(const TeamState& s) { return s.w[3] & 0xFu; }
__device__ __forceinline__ uint32_t tm_step_count(const TeamState& s) { return (s.w[3] >> 4) & 0x1Fu; }
__device__ __forceinline__ int tm_cur(const TeamState& s) { return (int)((s.w[3] >> 9) & 3u); }
__device__ __forceinline__ bool tm_terminal(const TeamState& s) { return (s.w[3] >> 11) & 1u; }
__device__ __forceinline__ int tm_lct(const TeamState& s) { return (int)((s.w[3] >> 12) & 3u) - 1; }   // -1 = None
__device__ __forceinline__ uint32_t tm_max_steps(const TeamState& s) { return (s.w[3] >> 14) & 0x1Fu; }
__device__ __forceinline__ uint32_t tm_scopas(const TeamState& s, int p) { return (s.w[6] >> (4 * p)) & 0xFu; }

// evaluate_game (team_mini_scopa_game.py:118-148): the cards left on the table go to the FIRST player of the team
// that captured last (the table itself is not cleared), team score = sum over its players of captures + 2 scopas,
// rewards are the scores minus their mean, [t0, t0, t1, t1]; all zero when nothing was scored.
__device__ __forceinline__ void tm_finish(TeamState& s, float* r4) {
    const int lct = tm_lct(s);
    const uint32_t len = tm_len(s);
    if (len && lct >= 0) s.w[4 + lct] |= table_set(s.w[2], len);      // first player of team t is player 2t: low half of w[4+t]
    int sc[2] = {0, 0};
#pragma unroll
    for (int p = 0; p < 4; p++) sc[p >> 1] += __popc(tm_caps(s, p)) + 2 * (int)tm_scopas(s, p);
    const float r0 = 0.5f * (float)(sc[0] - sc[1]);
    if (r4) { r4[0] = r0; r4[1] = r0; r4[2] = 0.f - r0; r4[3] = 0.f - r0; }
}

// TeamMiniScopaEnv.step (:171-205) + play_card (:101-116).  Illegal action = pass, terminal state = no-op.
__device__ __forceinline__ void tm_step(TeamState& s, uint32_t action, float* r4) {
    if (tm_terminal(s)) return;
    const int p = tm_cur(s);
    const uint32_t hand = tm_hand(s, p);
    if (action < 16u && ((hand >> action) & 1u)) {
        const uint32_t len = tm_len(s);
        uint32_t order = s.w[2];
        const uint32_t capm = capture_mask(order, len, action, table_set(order, len));
        if (capm) {
            uint32_t taken = 1u << action, k = len, m = capm;
            while (m) {
                const uint32_t i = 31u - (uint32_t)__clz((int)m);
                m ^= 1u << i;
                taken |= 1u << ((order >> (4u * i)) & 0xFu);
                order = nibble_remove(order, i);
                k--;
            }
            s.w[2] = order;
            s.w[4 + (p >> 1)] |= taken << (16 * (p & 1));
            s.w[3] = (s.w[3] & ~0xFu & ~(3u << 12)) | k | ((uint32_t)((p >> 1) + 1) << 12);   // last_capture_team
            if (k == 0u) s.w[6] += 1u << (4 * p);
        } else {
            s.w[2] = order | (action << (4u * len));
            s.w[3] += 1u;
        }
        s.w[p >> 1] &= ~((1u << action) << (16 * (p & 1)));
    }
    s.w[3] += 1u << 4;                                            // step_count++
    const bool term = ((s.w[0] | s.w[1]) == 0u) || (tm_step_count(s) >= tm_max_steps(s));
    s.w[3] = (s.w[3] & ~(3u << 9)) | ((uint32_t)((p + 1) & 3) << 9);
    if (term) { s.w[3] |= 1u << 11; tm_finish(s, r4); }
}

__device__ __forceinline__ uint32_t tm_legal_list(const TeamState& s, unsigned long long hand_order, int p, uint32_t& list) {
    list = 0u;
    if (tm_terminal(s)) return 0u;
    uint32_t hand = tm_hand(s, p), n = 0u;
    const uint32_t ord = (uint32_t)(hand_order >> (16 * p)) & 0xFFFFu;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t c = (ord >> (4 * i)) & 0xFu;
        if ((hand >> c) & 1u) { list |= c << (4u * n); n++; hand &= ~(1u << c); }
    }
    return n;
}

__device__ __forceinline__ void tm_load(const uint4* p, long long g, TeamState& s) {
    const uint4 a = p[2 * g], b = p[2 * g + 1];
    s.w[0] = a.x; s.w[1] = a.y; s.w[2] = a.z; s.w[3] = a.w; s.w[4] = b.x; s.w[5] = b.y; s.w[6] = b.z; s.w[7] = b.w;
}
__device__ __forceinline__ void tm_store(uint4* p, long long g, const TeamState& s) {
    p[2 * g] = make_uint4(s.w[0], s.w[1], s.w[2], s.w[3]);
    p[2 * g + 1] = make_uint4(s.w[4], s.w[5], s.w[6], s.w[7]);
}

// from the shuffled deck (nibble i = i-th card): four hands of four, empty table, player 0 to move
__global__ void __launch_bounds__(256) team_init_kernel(const unsigned long long* __restrict__ deck, long long n,
                                                        uint4* __restrict__ states) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        const unsigned long long d = deck[g];
        TeamState s;
#pragma unroll
        for (int i = 0; i < 8; i++) s.w[i] = 0u;
#pragma unroll
        for (int i = 0; i < 16; i++) s.w[i >> 3] |= (1u << (uint32_t)((d >> (4 * i)) & 0xFull)) << (16 * ((i >> 2) & 1));
        s.w[3] = 16u << 14;                                        // max_steps = 4 players x 4 cards (:165)
        tm_store(states, g, s);
    }
}

__global__ void __launch_bounds__(256) team_step_kernel(uint4* __restrict__ states, const uint8_t* __restrict__ actions,
                                                        float4* __restrict__ rewards, uint8_t* __restrict__ done, long long n) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        TeamState s;
        tm_load(states, g, s);
        const bool was = tm_terminal(s);
        float r4[4] = {0.f, 0.f, 0.f, 0.f};
        tm_step(s, (uint32_t)actions[g], r4);
        tm_store(states, g, s);
        if (was) {                                                 // rewards of a finished game stay what they were
            TeamState t = s;
            t.w[3] &= ~(3u << 12);                                 // no second sweep: score the piles as they are
            tm_finish(t, r4);
        }
        if (rewards) rewards[g] = make_float4(r4[0], r4[1], r4[2], r4[3]);
        if (done) done[g] = tm_terminal(s) ? 1 : 0;
    }
}

#define MS_TAG_TEAM 0x4D414554u   // "TEAM"
__global__ void __launch_bounds__(256) team_rollout_kernel(const uint4* __restrict__ states,
                                                           const unsigned long long* __restrict__ hand_order, long long n,
                                                           uint2 key, unsigned long long game_offset,
                                                           uint4* __restrict__ actions16, float4* __restrict__ rewards,
                                                           uint4* __restrict__ final_states) {
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        TeamState s;
        tm_load(states, g, s);
        const unsigned long long ho = hand_order[g], gid = game_offset + (unsigned long long)g;
        uint32_t acts[4] = {0u, 0u, 0u, 0u};
        float r4[4] = {0.f, 0.f, 0.f, 0.f};
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
        for (int ply = 0; ply < 16; ply++) {
            if ((ply & 3) == 0)
                x = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(ply >> 2), MS_TAG_TEAM), key);
            const int q = ply & 3;
            const uint32_t xw = q == 0 ? x.x : (q == 1 ? x.y : (q == 2 ? x.z : x.w));
            uint32_t list;
            const uint32_t nl = tm_legal_list(s, ho, tm_cur(s), list);
            const uint32_t a = nl ? (list >> (4u * __umulhi(xw, nl))) & 0xFu : 0u;
            tm_step(s, a, r4);
            acts[ply >> 2] |= a << (8 * (ply & 3));
        }
        if (actions16) actions16[g] = make_uint4(acts[0], acts[1], acts[2], acts[3]);
        if (rewards) rewards[g] = make_float4(r4[0], r4[1], r4[2], r4[3]);
        if (final_states) tm_store(final_states, g, s);
    }
}

// the deck kernel lives in ms_env.cu
int team_deck_from_seeds(const int64_t* d_seeds, int64_t n, uint64_t* d_deck, void* stream);

}  // namespace ms

#ifndef MS_HOST_RULES_ONLY   // tests/emu/ms_team_host.cpp compiles everything above for the host (CPU checks of the rules)

using namespace ms;

extern "C" {

int ms_team_deal_from_seeds(const int64_t* d_seeds, int64_t n, ms_team_state* d_states, uint64_t* d_hand_order, void* stream) {
    if (n < 0 || (n > 0 && (!d_seeds || !d_states || !d_hand_order))) return fail(MS_ERR_ARG, "ms_team_deal_from_seeds: bad argument");
    if (n == 0) return MS_OK;
    int rc = team_deck_from_seeds(d_seeds, n, d_hand_order, stream);     // the deck IS the hand order (all 16 cards are dealt)
    if (rc) return rc;
    team_init_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)d_hand_order, (long long)n,
                                                                           (uint4*)d_states);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_team_step(ms_team_state* d_states, const uint8_t* d_actions, float* d_rewards, uint8_t* d_done, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!d_states || !d_actions))) return fail(MS_ERR_ARG, "ms_team_step: bad argument");
    if (n == 0) return MS_OK;
    team_step_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>((uint4*)d_states, d_actions, (float4*)d_rewards,
                                                                           d_done, (long long)n);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_team_rollout_random(const ms_team_state* d_states, const uint64_t* d_hand_order, int64_t n, uint64_t philox_seed,
                           uint64_t game_offset, uint8_t* d_actions, float* d_rewards, ms_team_state* d_final, void* stream) {
    if (n < 0 || (n > 0 && (!d_states || !d_hand_order))) return fail(MS_ERR_ARG, "ms_team_rollout_random: bad argument");
    if (n == 0) return MS_OK;
    team_rollout_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)d_states, (const unsigned long long*)d_hand_order, (long long)n,
        make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)), (unsigned long long)game_offset, (uint4*)d_actions,
        (float4*)d_rewards, (uint4*)d_final);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_team_step_host(ms_team_state* h_states, const uint8_t* h_actions, float* h_rewards, uint8_t* h_done, int64_t n) {
    if (n < 0 || (n > 0 && (!h_states || !h_actions))) return fail(MS_ERR_ARG, "ms_team_step_host: bad argument");
    if (n == 0) return MS_OK;
    char* d = nullptr;
    const size_t o_a = 32 * (size_t)n, o_r = (o_a + n + 255) & ~(size_t)255, o_d = o_r + 16 * (size_t)n, tot = o_d + n + 256;
    MS_CUDA(cudaMalloc(&d, tot));
    MS_CUDA(cudaMemcpy(d, h_states, 32 * (size_t)n, cudaMemcpyHostToDevice));
    MS_CUDA(cudaMemcpy(d + o_a, h_actions, (size_t)n, cudaMemcpyHostToDevice));
    int rc = ms_team_step((ms_team_state*)d, (const uint8_t*)(d + o_a), (float*)(d + o_r), (uint8_t*)(d + o_d), n, nullptr);
    if (rc) { cudaFree(d); return rc; }
    MS_CUDA(cudaMemcpy(h_states, d, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    if (h_rewards) MS_CUDA(cudaMemcpy(h_rewards, d + o_r, 16 * (size_t)n, cudaMemcpyDeviceToHost));
    if (h_done) MS_CUDA(cudaMemcpy(h_done, d + o_d, (size_t)n, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaFree(d));
    return MS_OK;
}

int ms_team_deal_from_seeds_host(const int64_t* h_seeds, int64_t n, ms_team_state* h_states, uint64_t* h_hand_order) {
    if (n < 0 || (n > 0 && (!h_seeds || !h_states || !h_hand_order))) return fail(MS_ERR_ARG, "ms_team_deal_from_seeds_host: bad argument");
    if (n == 0) return MS_OK;
    char* d = nullptr;
    const size_t o_s = 8 * (size_t)n, o_h = o_s + 32 * (size_t)n;
    MS_CUDA(cudaMalloc(&d, o_h + 8 * (size_t)n));
    MS_CUDA(cudaMemcpy(d, h_seeds, 8 * (size_t)n, cudaMemcpyHostToDevice));
    int rc = ms_team_deal_from_seeds((const int64_t*)d, n, (ms_team_state*)(d + o_s), (uint64_t*)(d + o_h), nullptr);
    if (rc) { cudaFree(d); return rc; }
    MS_CUDA(cudaMemcpy(h_states, d + o_s, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaMemcpy(h_hand_order, d + o_h, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    MS_CUDA(cudaFree(d));
    return MS_OK;
}

}  // extern "C"
#endif  // MS_HOST_RULES_ONLY
