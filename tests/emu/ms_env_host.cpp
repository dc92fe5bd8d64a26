// tests/emu/ms_env_host.cpp -- the PRODUCT's env kernels (scopa_b200/csrc/ms_env.cu: deal_kernel with the windowed
// MT19937 seeding and the warp-uniform shuffle, deal_slow_kernel, full_deck_kernel, step_kernel, legal_kernel,
// capture_kernel, keys_kernel, rollout_kernel) compiled for the host.  Every kernel is a one-thread-per-game grid-stride
// loop; the only block-wide step is staging the seed-independent MT table in shared memory (a static array here, the
// barrier a no-op for a block of one thread), and the __constant__ table is a plain global filled by host_env_init().
// So "a grid of one block of one thread" runs every row in order.  Test infrastructure.
#include <cstdint>
#include <cuda_runtime.h>

#define MS_HOST_ONE_THREAD
#include "host_intrinsics.h"

#define MS_HOST_RULES_ONLY
#include "../../scopa_b200/csrc/ms_env.cu"

extern "C" {
// init_genrand(19650218): what ensure_mt_table() uploads into g_mt_init on a device
void host_env_init() {
    uint32_t* t = ms::g_mt_init;
    t[0] = 19650218u;
    for (int i = 1; i < 624; i++) t[i] = 1812433253u * (t[i - 1] ^ (t[i - 1] >> 30)) + (uint32_t)i;
}
void host_deal(const long long* seeds, long long n, uint32_t* states, uint32_t* hand_order, unsigned long long* deck, int zero_means_42) {
    ms::deal_kernel(seeds, n, (uint4*)states, hand_order, deck, zero_means_42);
}
void host_deal_slow(const long long* seeds, long long n, uint32_t* states, uint32_t* hand_order) {
    ms::deal_slow_kernel(seeds, n, (uint4*)states, hand_order);
}
void host_full_deck(const long long* seeds, long long n, unsigned long long* decks, int zero_means_42, int force_slow) {
    ms::full_deck_kernel(seeds, n, (ulonglong4*)decks, zero_means_42, force_slow);
}
// same signatures as tests/emu/ms_state_host.cpp (which restates the loops around the rule header); here the kernels run
void host_step(uint32_t* states, const uint8_t* actions, float* rewards, uint8_t* done, long long n) {
    ms::step_kernel((uint4*)states, actions, (float2*)rewards, done, n);
}
void host_legal(const uint32_t* states, const uint32_t* hand_order, int player, uint16_t* mask, uint8_t* ordered, uint8_t* count,
                uint8_t* capture, long long n) {
    ms::legal_kernel((const uint4*)states, hand_order, player, mask, (uchar4*)ordered, count, (uchar4*)capture, n);
}
void host_capture(const uint32_t* states, const uint8_t* cards, uint8_t* out, long long n) {
    ms::capture_kernel((const uint4*)states, cards, out, n);
}
void host_keys(const uint32_t* states, int player, uint64_t* keys, long long n) {
    ms::keys_kernel((const uint4*)states, player, (unsigned long long*)keys, n);
}
void host_rollout(const uint32_t* states, const uint32_t* hand_order, long long n, unsigned long long philox_seed,
                  unsigned long long game_offset, uint8_t* actions8, float* rewards, uint32_t* final_states) {
    ms::rollout_kernel((const uint4*)states, hand_order, n, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
                       game_offset, (uint2*)actions8, (float2*)rewards, (uint4*)final_states);
}
}
