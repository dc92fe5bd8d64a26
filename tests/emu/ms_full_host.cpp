// tests/emu/ms_full_host.cpp -- the PRODUCT's 40-card Scopa device code (scopa_b200/csrc/ms_full.cu: fs_step,
// fs_capture_mask, fs_evaluate, fs_legal_list AND the kernels full_init_kernel / full_step_kernel / full_legal_kernel /
// full_evaluate_kernel / full_rollout_kernel) compiled for the host.  The kernels have no barriers and walk their rows
// with a grid-stride loop; their shared-memory scratch is one column per thread (a static array here), and the warp
// votes that only share loop bounds between lanes reduce to the lane's own value.  So "a grid of one block of one
// thread" runs every row in order.  Test infrastructure.
#include <cstdint>
#include <cuda_runtime.h>

#define MS_HOST_ONE_THREAD
#include "host_intrinsics.h"

#define MS_HOST_RULES_ONLY
#include "../../scopa_b200/csrc/ms_full.cu"

extern "C" {
void host_full_init(const unsigned long long* decks, long long n, uint32_t* states) {
    ms::full_init_kernel((const ulonglong4*)decks, n, (uint4*)states);
}
unsigned host_full_step(uint32_t* states, const unsigned long long* decks, const uint8_t* actions, float* rewards, uint8_t* done,
                        long long n) {
    unsigned overflow = 0;
    ms::full_step_kernel((uint4*)states, (const ulonglong4*)decks, actions, (float2*)rewards, done, n, &overflow);
    return overflow;
}
void host_full_legal(const uint32_t* states, const unsigned long long* decks, int player, uint8_t* ordered, uint8_t* count,
                     long long n) {
    ms::full_legal_kernel((const uint4*)states, (const ulonglong4*)decks, player, ordered, count, n);
}
void host_full_evaluate(uint32_t* states, float* rewards, int* detail, long long n) {
    ms::full_evaluate_kernel((uint4*)states, (float2*)rewards, detail, n);
}
unsigned host_full_rollout(const uint32_t* states, const unsigned long long* decks, long long n, unsigned long long philox_seed,
                           unsigned long long game_offset, uint8_t* actions36, float* rewards, uint32_t* final_states) {
    unsigned overflow = 0;
    ms::full_rollout_kernel((const uint4*)states, (const ulonglong4*)decks, n,
                            make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)), game_offset, actions36,
                            (float2*)rewards, (uint4*)final_states, &overflow);
    return overflow;
}
}
