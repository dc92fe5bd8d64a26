// tests/emu/ms_state_host.cpp -- the PRODUCT's Miniscopa rule code (scopa_b200/csrc/ms_state.cuh: capture_mask, step,
// legal_list, infoset_key, reward0 -- the __device__ functions every env / solver kernel calls) compiled for the host,
// so that `pytest -m "not gpu"` checks the shipped rules against the reference-generated fixtures on machines without a
// GPU.  Test infrastructure.  The few device intrinsics the header uses get their one-lane meaning here:
// __reduce_max_sync over a "warp" of one lane is the lane's own value (the header uses it only as a shared loop bound).
// The thin loops below restate what step_kernel / legal_kernel / capture_kernel / keys_kernel do around those functions
// (scopa_b200/csrc/ms_env.cu:368-445): per-row load, rule call, store.
#include <cstdint>
#include <cuda_runtime.h>   // vector types (uint4, make_uint4) for the host compiler

#define MS_HOST_ONE_THREAD
#include "host_intrinsics.h"

#include "../../scopa_b200/csrc/ms_state.cuh"

using namespace ms;

extern "C" {

// step_kernel
void host_step(uint32_t* states, const uint8_t* actions, float* rewards, uint8_t* done, long long n) {
    for (long long g = 0; g < n; ++g) {
        MsState s = make_uint4(states[4 * g], states[4 * g + 1], states[4 * g + 2], states[4 * g + 3]);
        step(s, (uint32_t)actions[g]);
        states[4 * g] = s.x; states[4 * g + 1] = s.y; states[4 * g + 2] = s.z; states[4 * g + 3] = s.w;
        const bool term = st_terminal(s);
        const float r0 = term ? reward0(s) : 0.f;
        rewards[2 * g] = r0;
        rewards[2 * g + 1] = 0.f - r0;
        done[g] = term ? 1 : 0;
    }
}

// legal_kernel (player < 0: the player to move)
void host_legal(const uint32_t* states, const uint32_t* hand_order, int player, uint16_t* mask, uint8_t* ordered,
                uint8_t* count, uint8_t* capture, long long n) {
    for (long long g = 0; g < n; ++g) {
        const MsState s = make_uint4(states[4 * g], states[4 * g + 1], states[4 * g + 2], states[4 * g + 3]);
        const int p = player < 0 ? st_cur(s) : player;
        uint32_t list;
        const uint32_t nl = legal_list(s, hand_order[g], p, list);
        const uint32_t hand = st_hand(s, p), tset = table_set(s.y, st_table_len(s));
        uint32_t m = 0u;
        for (int k = 0; k < 4; ++k) {
            ordered[4 * g + k] = 0xFF;
            capture[4 * g + k] = 0;
            if ((uint32_t)k < nl) {
                const uint32_t a = (list >> (4 * k)) & 0xFu;
                m |= 1u << a;
                ordered[4 * g + k] = (uint8_t)a;
                capture[4 * g + k] = ((hand >> a) & 1u) ? (uint8_t)capture_mask(s.y, st_table_len(s), a, tset) : 0;
            }
        }
        mask[g] = (uint16_t)m;
        count[g] = (uint8_t)nl;
    }
}

// capture_kernel
void host_capture(const uint32_t* states, const uint8_t* cards, uint8_t* out, long long n) {
    for (long long g = 0; g < n; ++g) {
        const MsState s = make_uint4(states[4 * g], states[4 * g + 1], states[4 * g + 2], states[4 * g + 3]);
        const uint32_t card = cards[g] & 0xFu, len = st_table_len(s);
        const uint32_t tset = table_set(s.y, len);
        uint32_t m = capture_mask(s.y, len, card, tset);
        if ((tset >> card) & 1u) {                    // the test hook's extra case, as in capture_kernel (ms_env.cu:428-431)
            const uint32_t self = 1u << nibble_pos(s.y, card);
            m = ((tset >> card_twin(card)) & 1u) ? (m < self ? m : self) : self;
        }
        out[g] = (uint8_t)m;
    }
}

// keys_kernel
void host_keys(const uint32_t* states, int player, uint64_t* keys, long long n) {
    for (long long g = 0; g < n; ++g) {
        const MsState s = make_uint4(states[4 * g], states[4 * g + 1], states[4 * g + 2], states[4 * g + 3]);
        const int p = player < 0 ? st_cur(s) : player;
        keys[g] = st_terminal(s) ? 0xFFFFFFFFFFFFFFFFull : infoset_key(s, p);
    }
}

}  // extern "C"
