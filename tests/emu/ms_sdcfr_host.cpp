// tests/emu/ms_sdcfr_host.cpp -- the PRODUCT's SDCFR kernels, fp32 path (scopa_b200/csrc/ms_sdcfr.cu: sd_mlp_kernel<0>,
// and the level-batched external-sampling traversal sd_init_kernel / sd_level_mlp_kernel<0> + sd_expand_kernel / sd_forced_kernel /
// sd_terminal_kernel / sd_backward_kernel / sd_root_value_kernel) executed on the host by the CTA emulator of
// tests/emu/cta_emu.h.  The tcgen05 / TMEM / mbarrier path (PREC = 1) is inline PTX and has no host meaning: its
// helpers are parsed but never instantiated here.  __syncwarp is a barrier over the 32 emulated threads of a warp
// (sd_backward_kernel stages rows per warp); fp32 arithmetic is separate mul / add like the device build
// (nvcc --fmad=false; this file: -ffp-contract=off).  The launch sequence of ms_sdcfr_traverse is restated below.
// Test infrastructure.
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <vector>

#include "host_intrinsics.h"
#include "cta_emu.h"
static pthread_barrier_t emu_warp_barrier[64];
static inline void __syncwarp() { pthread_barrier_wait(&emu_warp_barrier[threadIdx.x >> 5]); }

#define MS_HOST_RULES_ONLY
#include "../../scopa_b200/csrc/ms_sdcfr.cu"

namespace ms {   // declared in ms_common.cuh, defined in ms_env.cu in the library
std::atomic<uint64_t> g_launches{0};
char* last_error_buf() { static thread_local char buf[512]; return buf; }
}

namespace {
using namespace ms;

unsigned grid_of(long long n, int block, int per_sm) {
    long long need = (n + block - 1) / block, cap = 148LL * per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

// kernel<<<grid, threads>>> with per-warp barriers available to __syncwarp
template <class K, class A>
int launch(K k, const A& a, unsigned grid, unsigned threads) {
    for (unsigned w = 0; w < threads / 32; w++) pthread_barrier_init(&emu_warp_barrier[w], nullptr, 32);
    const int rc = emu_launch_grid(k, a, grid, threads);
    for (unsigned w = 0; w < threads / 32; w++) pthread_barrier_destroy(&emu_warp_barrier[w]);
    return rc;
}

struct MlpArgs { const float* net; const float* feat; const float* mask; float* adv; float* pol; long long n; };
void mlp_entry(MlpArgs a) { sd_mlp_kernel<0>(a.net, nullptr, a.feat, a.mask, a.adv, a.pol, a.n); }
struct LvlArgs { SdArgs a; int d; };
void level_mlp_entry(LvlArgs x) { sd_level_mlp_kernel<0>(x.a, x.d); }
void expand_entry(LvlArgs x) { sd_expand_kernel(x.a, x.d); }
void forced_entry(LvlArgs x) { sd_forced_kernel(x.a, x.d); }
void backward_entry(LvlArgs x) { sd_backward_kernel(x.a, x.d); }
struct InitArgs { SdArgs a; uint4 root; };
void init_entry(InitArgs x) { sd_init_kernel(x.a, x.root); }
void terminal_entry(SdArgs a) { sd_terminal_kernel(a); }
void root_value_entry(SdArgs a) { sd_root_value_kernel(a); }
}  // namespace

extern "C" {

int host_sd_samples_per_traversal(int player) {
    SdShape sh;
    sd_shape(player & 1, sh);
    return sh.samples;
}

// ms_mlp_forward, precision 0
int host_sd_mlp_forward(const float* net, const float* feat, const float* mask, float* adv, float* pol, long long n) {
    if (SD_SMEM_FP32 > EMU_SMEM_BYTES) return -4;
    MlpArgs a{net, feat, mask, adv, pol, n};
    return launch(mlp_entry, a, grid_of(n, SD_TILE, 1), SD_TILE);
}

// ms_sdcfr_traverse, precision 0
int host_sd_traverse(const uint32_t* root4, uint32_t hand_order, int player, const float* net0, const float* net1,
                     long long n_trav, unsigned long long philox_seed, unsigned long long first_trav, float* feat,
                     float* target, float* mask, float* root_value) {
    SdArgs a{};
    const size_t need = sd_workspace(n_trav, player, nullptr, nullptr);
    std::vector<char> ws(need + 256);
    char* base = (char*)(((uintptr_t)ws.data() + 255) & ~(uintptr_t)255);
    sd_workspace(n_trav, player, &a, base);
    a.net[0] = net0; a.net[1] = net1;
    a.hand_order = hand_order;
    a.pkey = make_uint2((uint32_t)philox_seed, (uint32_t)(philox_seed >> 32));
    a.first_trav = first_trav; a.n_trav = n_trav;
    a.out_feat = feat; a.out_target = target; a.out_mask = mask; a.out_value = root_value;
    InitArgs ia{a, make_uint4(root4[0], root4[1], root4[2], root4[3])};
    if (launch(init_entry, ia, grid_of(n_trav, 256, 4), 256)) return -1;
    for (int d = 0; d < 8; d++) {
        const bool forced_opp = ((d & 1) != player) && (4 - d / 2 == 1);   // the opponent's last card
        LvlArgs la{a, d};
        if (forced_opp) {
            if (launch(forced_entry, la, grid_of(n_trav * a.sh.n[d], 256, 8), 256)) return -1;
            continue;
        }
        if (launch(level_mlp_entry, la, grid_of(n_trav * a.sh.n[d], SD_TILE, 1), SD_TILE)) return -1;
        if (launch(expand_entry, la, grid_of(n_trav * a.sh.n[d], 256, 8), 256)) return -1;
    }
    if (launch(terminal_entry, a, grid_of(n_trav * a.sh.n[8], 256, 8), 256)) return -1;
    for (int d = 7; d >= 0; d--) {
        LvlArgs la{a, d};
        if (launch(backward_entry, la, grid_of(n_trav * a.sh.n[d], 256, 8), 256)) return -1;
    }
    if (root_value && launch(root_value_entry, a, grid_of(n_trav, 256, 4), 256)) return -1;
    return 0;
}

}  // extern "C"
