// tests/emu/ms_team_host.cpp -- the PRODUCT's 2v2 team-Miniscopa device code (scopa_b200/csrc/ms_team.cu: tm_step,
// tm_finish, tm_legal_list AND the three kernels team_init_kernel / team_step_kernel / team_rollout_kernel) compiled for
// the host.  The kernels have no shared memory or barriers and walk their rows with a grid-stride loop, so "a grid of one
// block of one thread" (blockIdx = threadIdx = 0, blockDim = gridDim = 1) runs every row in order.  Test infrastructure.
#include <cstdint>
#include <cuda_runtime.h>

#define MS_HOST_ONE_THREAD
#include "host_intrinsics.h"

#define MS_HOST_RULES_ONLY
#include "../../scopa_b200/csrc/ms_team.cu"

extern "C" {
void host_team_init(const unsigned long long* deck, long long n, uint32_t* states) { ms::team_init_kernel(deck, n, (uint4*)states); }
void host_team_step(uint32_t* states, const uint8_t* actions, float* rewards, uint8_t* done, long long n) {
    ms::team_step_kernel((uint4*)states, actions, (float4*)rewards, done, n);
}
void host_team_rollout(const uint32_t* states, const unsigned long long* hand_order, long long n, unsigned long long philox_seed,
                       unsigned long long game_offset, uint8_t* actions16, float* rewards, uint32_t* final_states) {
    ms::team_rollout_kernel((const uint4*)states, hand_order, n, make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32)),
                            game_offset, (uint4*)actions16, (float4*)rewards, (uint4*)final_states);
}
}
