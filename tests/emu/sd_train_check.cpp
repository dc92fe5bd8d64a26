// tests/emu/sd_train_check.cpp -- torch-free GPU check of ms_sdcfr_train: the same seeded problem is run through the
// C ABI (libscopa_b200.so, sd_train_kernel on the device) and through the host emulation of the same kernel source
// (libsd_train_emu.so); parameters, Adam moments and losses must agree BIT FOR BIT.  Then times the launch.
// usage: sd_train_check [batch] [epochs] [n_rows] [timing_reps]        exit 0 = identical
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <vector>

#include "../../include/scopa_b200.h"

extern "C" int emu_ms_sdcfr_average_policy(const float*, const float*, int, const float*, const float*, long long, float*,
                                           void*, size_t, void*);
extern "C" int emu_ms_sdcfr_sample_rows(int*, int, int, long long, unsigned long long, unsigned long long, void*);
extern "C" int emu_ms_sdcfr_train(float*, float*, float*, long long, const float*, const float*, const float*, long long,
                                  const int*, int, int, double, double, double, double, double, float*, void*, size_t, void*);
extern "C" int emu_ms_sdcfr_train_cluster(float*, float*, float*, long long, const float*, const float*, const float*,
                                          long long, const int*, int, int, double, double, double, double, double, float*,
                                          void*, size_t, void*);

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(rng_state >> 33);
}
static float uni() { return (rnd() & 0xFFFFFF) / 16777216.0f; }

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

// sd_train_check avgpol [n_nets] [n_rows] [timing_reps]: ms_sdcfr_average_policy on the device against the emulation
static int check_avgpol(int K, int n, int reps) {
    const int NF = 13776;
    std::vector<float> nets((size_t)K * NF), w(K), feat((size_t)n * 34), mask((size_t)n * 16, 0.f), pol_e((size_t)n * 16),
        ws_e((size_t)K * n * 16);
    for (auto& x : nets) x = (2 * uni() - 1) * 0.2f;
    double tot = 0;
    for (int k = 0; k < K; ++k) tot += k + 2;
    for (int k = 0; k < K; ++k) w[k] = (float)((k + 2) / tot);
    for (auto& x : feat) x = uni() < 0.3f ? 1.f : 0.f;
    for (int r = 0; r < n; ++r)
        for (int j = 0, nl = 1 + rnd() % 4; j < nl; ++j) mask[(size_t)r * 16 + rnd() % 16] = 1.f;
    const bool timing_only = getenv("SD_CHECK_TIMING_ONLY") != nullptr;     // skip the (slow) emulation of big cases
    if (!timing_only &&
        emu_ms_sdcfr_average_policy(nets.data(), w.data(), K, feat.data(), mask.data(), n, pol_e.data(), ws_e.data(),
                                    ws_e.size() * 4, nullptr)) {
        printf("emulation failed\n");
        return 2;
    }
    if (getenv("SD_TRAIN_EMU_ONLY")) {
        double s = 0;
        for (float x : pol_e) s += x;
        printf("emu avgpol: sum of policies %.6f over %d rows\n", s, n);
        return 0;
    }
    float *d_nets, *d_w, *d_feat, *d_mask, *d_pol;
    void* d_ws;
    size_t wsb = ms_sdcfr_average_policy_workspace_bytes(K, n);
    CK(cudaMalloc(&d_nets, nets.size() * 4)); CK(cudaMalloc(&d_w, K * 4)); CK(cudaMalloc(&d_feat, feat.size() * 4));
    CK(cudaMalloc(&d_mask, mask.size() * 4)); CK(cudaMalloc(&d_pol, pol_e.size() * 4)); CK(cudaMalloc(&d_ws, wsb));
    CK(cudaMemcpy(d_nets, nets.data(), nets.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w.data(), K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_feat, feat.data(), feat.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_mask, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice));
    int rc = ms_sdcfr_average_policy(d_nets, d_w, K, d_feat, d_mask, n, d_pol, d_ws, wsb, nullptr);
    if (rc) {
        printf("ms_sdcfr_average_policy failed: %d %s\n", rc, ms_last_error());
        return 2;
    }
    CK(cudaDeviceSynchronize());
    std::vector<float> pol_g(pol_e.size());
    CK(cudaMemcpy(pol_g.data(), d_pol, pol_g.size() * 4, cudaMemcpyDeviceToHost));
    size_t d = 0;
    double s = 0;
    for (size_t i = 0; i < pol_g.size(); ++i) {
        if (!timing_only) d += memcmp(&pol_g[i], &pol_e[i], 4) != 0;
        s += pol_g[i];
    }
    printf("avgpol nets %d rows %d: differing words %zu of %zu%s; sum of policies %.6f\n", K, n, d, pol_g.size(),
           timing_only ? " (NOT compared: timing only)" : "", s);
    if (reps > 0) {
        cudaEvent_t t0, t1;
        CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        for (int i = 0; i < 3; ++i) ms_sdcfr_average_policy(d_nets, d_w, K, d_feat, d_mask, n, d_pol, d_ws, wsb, nullptr);
        CK(cudaEventRecord(t0));
        for (int i = 0; i < reps; ++i) ms_sdcfr_average_policy(d_nets, d_w, K, d_feat, d_mask, n, d_pol, d_ws, wsb, nullptr);
        CK(cudaEventRecord(t1));
        CK(cudaEventSynchronize(t1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("timing: %.2f us per call (%d nets x %d rows, two launches)\n", 1e3 * ms / reps, K, n);
    }
    return d ? 1 : 0;
}

// sd_train_check sample [batch] [epochs] [n_rows] [timing_reps]: ms_sdcfr_sample_rows on the device against the emulation
static int check_sample(int batch, int epochs, long long n_rows, int reps) {
    std::vector<int> e((size_t)epochs * batch), g((size_t)epochs * batch);
    const unsigned long long seed = 0x1234567890ull, first = 17;
    if (emu_ms_sdcfr_sample_rows(e.data(), batch, epochs, n_rows, seed, first, nullptr)) {
        printf("emulation failed\n");
        return 2;
    }
    int* d_idx;
    CK(cudaMalloc(&d_idx, g.size() * 4));
    int rc = ms_sdcfr_sample_rows(d_idx, batch, epochs, n_rows, seed, first, nullptr);
    if (rc) {
        printf("ms_sdcfr_sample_rows failed: %d %s\n", rc, ms_last_error());
        return 2;
    }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(g.data(), d_idx, g.size() * 4, cudaMemcpyDeviceToHost));
    size_t d = 0;
    for (size_t i = 0; i < g.size(); ++i) d += g[i] != e[i];
    printf("sample batch %d epochs %d rows %lld: differing words %zu of %zu; first rows %d %d %d\n", batch, epochs, n_rows, d,
           g.size(), g[0], g.size() > 1 ? g[1] : -1, g.size() > 2 ? g[2] : -1);
    if (reps > 0) {
        cudaEvent_t t0, t1;
        CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        for (int i = 0; i < 3; ++i) ms_sdcfr_sample_rows(d_idx, batch, epochs, n_rows, seed, first, nullptr);
        CK(cudaEventRecord(t0));
        for (int i = 0; i < reps; ++i) ms_sdcfr_sample_rows(d_idx, batch, epochs, n_rows, seed, first + i, nullptr);
        CK(cudaEventRecord(t1));
        CK(cudaEventSynchronize(t1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("timing: %.2f us per call (%d epochs x %d rows of %lld)\n", 1e3 * ms / reps, epochs, batch, n_rows);
    }
    return d ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "sample"))
        return check_sample(argc > 2 ? atoi(argv[2]) : 128, argc > 3 ? atoi(argv[3]) : 10, argc > 4 ? atoll(argv[4]) : 100000,
                            argc > 5 ? atoi(argv[5]) : 0);
    if (argc > 1 && !strcmp(argv[1], "avgpol"))
        return check_avgpol(argc > 2 ? atoi(argv[2]) : 8, argc > 3 ? atoi(argv[3]) : 70, argc > 4 ? atoi(argv[4]) : 0);
    // sd_train_check [cluster] [batch] [epochs] [n_rows] [timing_reps]: "cluster" selects ms_sdcfr_train_cluster
    const bool cluster = argc > 1 && !strcmp(argv[1], "cluster");
    if (cluster) { --argc; ++argv; }
    const int batch = argc > 1 ? atoi(argv[1]) : 128, epochs = argc > 2 ? atoi(argv[2]) : 6;
    const int n_rows = argc > 3 ? atoi(argv[3]) : 4096, reps = argc > 4 ? atoi(argv[4]) : 50;
    typedef int (*train_fn)(float*, float*, float*, int64_t, const float*, const float*, const float*, int64_t, const int32_t*,
                            int32_t, int32_t, double, double, double, double, double, float*, void*, size_t, void*);
    typedef int (*emu_fn)(float*, float*, float*, long long, const float*, const float*, const float*, long long, const int*,
                          int, int, double, double, double, double, double, float*, void*, size_t, void*);
    const train_fn dev_train = cluster ? ms_sdcfr_train_cluster : ms_sdcfr_train;
    const emu_fn emu_train = cluster ? emu_ms_sdcfr_train_cluster : emu_ms_sdcfr_train;
    const char* which = cluster ? "cluster kernel" : "one-CTA kernel";
    const int NF = 13776;
    std::vector<float> net(NF), m(NF, 0.f), v(NF, 0.f), feat((size_t)n_rows * 34), target((size_t)n_rows * 16),
        mask((size_t)n_rows * 16), loss_e(epochs), grad(NF);
    // xavier-like weights, bias 0.1
    const int fan[3][2] = {{128, 34}, {64, 128}, {16, 64}};
    int off = 0;
    for (auto& f : fan) {
        float lim = sqrtf(6.0f / (f[0] + f[1]));
        for (int i = 0; i < f[0] * f[1]; ++i) net[off++] = (2 * uni() - 1) * lim;
        for (int i = 0; i < f[0]; ++i) net[off++] = 0.1f;
    }
    for (auto& x : feat) x = uni() < 0.25f ? 1.f : 0.f;
    for (int r = 0; r < n_rows; ++r) {
        int nl = 1 + rnd() % 4;
        for (int j = 0; j < nl; ++j) mask[(size_t)r * 16 + rnd() % 16] = 1.f;
        for (int j = 0; j < 16; ++j) target[(size_t)r * 16 + j] = (2 * uni() - 1) * mask[(size_t)r * 16 + j] * (r % 3 ? 1.f : 5.f);
    }
    std::vector<int> idx((size_t)epochs * batch);
    for (int e = 0; e < epochs; ++e) {          // distinct rows per epoch
        std::vector<int> perm(n_rows);
        for (int i = 0; i < n_rows; ++i) perm[i] = i;
        for (int i = 0; i < batch; ++i) {
            int j = i + rnd() % (n_rows - i);
            std::swap(perm[i], perm[j]);
            idx[(size_t)e * batch + i] = perm[i];
        }
    }
    if (getenv("SD_TRAIN_EMU_ONLY")) {
        std::vector<float> h = net;
        emu_train(h.data(), m.data(), v.data(), 0, feat.data(), target.data(), mask.data(), n_rows, idx.data(), batch,
                  epochs, 5e-4, 0.9, 0.999, 1e-8, 1.0, loss_e.data(), grad.data(), grad.size() * 4, nullptr);
        for (int e = 0; e < epochs; ++e) printf("emu loss[%d] %.9g (%a)\n", e, loss_e[e], loss_e[e]);
        return 0;
    }
    // ---- device run through the C ABI
    float *d_net, *d_m, *d_v, *d_feat, *d_target, *d_mask, *d_loss;
    int* d_idx;
    void* d_ws;
    size_t ws = ms_sdcfr_train_workspace_bytes();
    CK(cudaMalloc(&d_net, NF * 4)); CK(cudaMalloc(&d_m, NF * 4)); CK(cudaMalloc(&d_v, NF * 4));
    CK(cudaMalloc(&d_feat, feat.size() * 4)); CK(cudaMalloc(&d_target, target.size() * 4));
    CK(cudaMalloc(&d_mask, mask.size() * 4)); CK(cudaMalloc(&d_loss, epochs * 4));
    CK(cudaMalloc(&d_idx, idx.size() * 4)); CK(cudaMalloc(&d_ws, ws));
    CK(cudaMemcpy(d_net, net.data(), NF * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_m, 0, NF * 4)); CK(cudaMemset(d_v, 0, NF * 4));
    CK(cudaMemcpy(d_feat, feat.data(), feat.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_target, target.data(), target.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_mask, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
    // two calls (steps_done 0, then epochs/2) to cover the bias-correction hand-over
    const int e1 = epochs / 2, e2 = epochs - e1;
    int rc = dev_train(d_net, d_m, d_v, 0, d_feat, d_target, d_mask, n_rows, d_idx, batch, e1, 5e-4, 0.9, 0.999, 1e-8,
                            1.0, d_loss, d_ws, ws, nullptr);
    if (rc == 0)
        rc = dev_train(d_net, d_m, d_v, e1, d_feat, d_target, d_mask, n_rows, d_idx + (size_t)e1 * batch, batch, e2,
                            5e-4, 0.9, 0.999, 1e-8, 1.0, d_loss + e1, d_ws, ws, nullptr);
    if (rc) {
        printf("%s failed: %d %s\n", which, rc, ms_last_error());
        return 2;
    }
    CK(cudaDeviceSynchronize());
    std::vector<float> g_net(NF), g_m(NF), g_v(NF), g_loss(epochs);
    CK(cudaMemcpy(g_net.data(), d_net, NF * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g_m.data(), d_m, NF * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g_v.data(), d_v, NF * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g_loss.data(), d_loss, epochs * 4, cudaMemcpyDeviceToHost));
    // ---- host emulation of the same kernel source
    std::vector<float> h_net = net;
    if (emu_train(h_net.data(), m.data(), v.data(), 0, feat.data(), target.data(), mask.data(), n_rows, idx.data(), batch, e1,
                  5e-4, 0.9, 0.999, 1e-8, 1.0, loss_e.data(), grad.data(), grad.size() * 4, nullptr) ||
        emu_train(h_net.data(), m.data(), v.data(), e1, feat.data(), target.data(), mask.data(), n_rows,
                  idx.data() + (size_t)e1 * batch, batch, e2, 5e-4, 0.9, 0.999, 1e-8, 1.0, loss_e.data() + e1, grad.data(),
                  grad.size() * 4, nullptr)) {
        printf("emulation failed\n");
        return 2;
    }
    auto ndiff = [](const std::vector<float>& a, const std::vector<float>& b) {
        size_t d = 0;
        for (size_t i = 0; i < a.size(); ++i) d += memcmp(&a[i], &b[i], 4) != 0;
        return d;
    };
    size_t dn = ndiff(g_net, h_net), dm = ndiff(g_m, m), dv = ndiff(g_v, v), dl = ndiff(g_loss, loss_e);
    double moved = 0;
    for (int i = 0; i < NF; ++i) moved = fmax(moved, fabs((double)g_net[i] - net[i]));
    printf("%s, batch %d epochs %d rows %d: differing words net %zu m %zu v %zu loss %zu (of %d / %d); max |param change| %.3g; "
           "loss[0] %.6f loss[last] %.6f\n", which, batch, epochs, n_rows, dn, dm, dv, dl, NF, epochs, moved, g_loss[0],
           g_loss[epochs - 1]);
    // ---- timing: `reps` launches of `epochs` steps
    if (reps > 0) {
        cudaEvent_t t0, t1;
        CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        for (int w = 0; w < 3; ++w)
            dev_train(d_net, d_m, d_v, epochs, d_feat, d_target, d_mask, n_rows, d_idx, batch, epochs, 5e-4, 0.9, 0.999,
                           1e-8, 1.0, d_loss, d_ws, ws, nullptr);
        CK(cudaEventRecord(t0));
        for (int r = 0; r < reps; ++r)
            dev_train(d_net, d_m, d_v, epochs, d_feat, d_target, d_mask, n_rows, d_idx, batch, epochs, 5e-4, 0.9, 0.999,
                           1e-8, 1.0, d_loss, d_ws, ws, nullptr);
        CK(cudaEventRecord(t1));
        CK(cudaEventSynchronize(t1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("timing: %.2f us per launch of %d steps = %.2f us per optimiser step (batch %d)\n", 1e3 * ms / reps, epochs,
               1e3 * ms / reps / epochs, batch);
    }
    return (dn | dm | dv | dl) ? 1 : 0;
}
