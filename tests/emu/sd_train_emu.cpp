// tests/emu/sd_train_emu.cpp -- sd_train_kernel (scopa_b200/csrc/ms_sd_train.cuh) executed on the host through
// cta_emu.h.  Built by tests/test_sd_train_emu.py with g++ -O2 -ffp-contract=off -mfma (fmaf = one rounding, as FFMA).
#include "cta_emu.h"
// libscopa_b200.so exports a host stub with the kernel's mangled name; when both libraries sit in one process the
// dynamic linker would bind the emulation's calls to that stub.  The emulation therefore lives in its own namespace.
#define ms ms_emu
#include "../../scopa_b200/csrc/ms_sd_train.cuh"
#include "../../scopa_b200/csrc/ms_sd_avgpol.cuh"
#include "../../scopa_b200/csrc/ms_sd_train_cluster.cuh"
#include "../../scopa_b200/csrc/ms_sd_sample.cuh"

extern "C" int emu_sd_train(float* net, float* adam_m, float* adam_v, long long steps_done, const float* feat,
                            const float* target, const float* mask, long long n_rows, const int* idx, int batch,
                            int epochs, double lr, double beta1, double beta2, double eps, double max_norm, float* loss,
                            float* grad) {
    ms::SdTrainArgs a;
    a.net = net; a.adam_m = adam_m; a.adam_v = adam_v;
    a.feat = feat; a.target = target; a.mask = mask; a.n_rows = n_rows;
    a.idx = idx; a.batch = batch; a.epochs = epochs;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
    a.b1pow = std::pow(beta1, (double)steps_done);
    a.b2pow = std::pow(beta2, (double)steps_done);
    a.loss = loss; a.grad = grad;
    return emu_launch_cta(ms::sd_train_kernel, a, ms::sdt::kThreads);
}

extern "C" int emu_sd_train_smem_bytes() { return ms::sdt::kSmemBytes; }

// Same signature as the C ABI's ms_sdcfr_train (include/scopa_b200.h), host pointers instead of device pointers:
// lets tests drive the Python binding (scopa_b200.sdcfr.FusedAdam) end to end on a machine without a GPU.
extern "C" int emu_ms_sdcfr_train(float* net, float* adam_m, float* adam_v, long long steps_done, const float* feat,
                                  const float* target, const float* mask, long long n_rows, const int* idx, int batch,
                                  int epochs, double lr, double beta1, double beta2, double eps, double max_norm,
                                  float* loss, void* workspace, size_t workspace_bytes, void* /*stream*/) {
    if (workspace_bytes < sizeof(float) * ms::sdt::kNetFloats || batch < 1 || batch > ms::sdt::kMaxBatch) return -2;
    if (epochs == 0) return 0;
    return emu_sd_train(net, adam_m, adam_v, steps_done, feat, target, mask, n_rows, idx, batch, epochs, lr, beta1, beta2,
                        eps, max_norm, loss, static_cast<float*>(workspace));
}

// Same signature as ms_sdcfr_average_policy (include/scopa_b200.h) with host pointers.  `grid` CTAs like the device
// launch; the emulation runs them one after another.
extern "C" int emu_ms_sdcfr_average_policy(const float* nets, const float* weights, int n_nets, const float* feat,
                                           const float* mask, long long n_rows, float* policy, void* workspace,
                                           size_t workspace_bytes, void* /*stream*/) {
    if (n_nets < 1 || n_rows < 0) return -2;
    if (n_rows == 0) return 0;
    if (workspace_bytes < (size_t)n_nets * (size_t)n_rows * 16 * sizeof(float)) return -2;
    ms::SdAvgPolArgs a;
    a.nets = nets; a.weights = weights; a.n_nets = n_nets; a.feat = feat; a.mask = mask; a.n_rows = n_rows;
    a.scratch = static_cast<float*>(workspace); a.policy = policy;
    const unsigned gx = n_nets < 148 ? n_nets : 148;                    // the device launch's grid (ms_sd_train.cu)
    const long long n_chunks = (n_rows + ms::sda::kRows - 1) / ms::sda::kRows;
    long long gy = (2 * 148 + gx - 1) / gx;
    if (gy > n_chunks) gy = n_chunks;
    int rc = emu_launch_grid(ms::sd_avgpol_kernel, a, gx, ms::sda::kPolThreads, (unsigned)gy);
    if (rc) return rc;
    return emu_launch_grid(ms::sd_avgpol_reduce_kernel, a, 2, 256);
}

// ms_sdcfr_train_cluster's signature; the 8 CTAs of the cluster run concurrently (8 x 256 host threads)
extern "C" int emu_ms_sdcfr_train_cluster(float* net, float* adam_m, float* adam_v, long long steps_done, const float* feat,
                                          const float* target, const float* mask, long long n_rows, const int* idx,
                                          int batch, int epochs, double lr, double beta1, double beta2, double eps,
                                          double max_norm, float* loss, void* workspace, size_t workspace_bytes,
                                          void* /*stream*/) {
    if (workspace_bytes < sizeof(float) * ms::sdt::kNetFloats || batch < 1 || batch > ms::sdt::kMaxBatch) return -2;
    if (epochs == 0) return 0;
    ms::SdTrainArgs a;
    a.net = net; a.adam_m = adam_m; a.adam_v = adam_v;
    a.feat = feat; a.target = target; a.mask = mask; a.n_rows = n_rows;
    a.idx = idx; a.batch = batch; a.epochs = epochs;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
    a.b1pow = std::pow(beta1, (double)steps_done);
    a.b2pow = std::pow(beta2, (double)steps_done);
    a.loss = loss; a.grad = static_cast<float*>(workspace);
    return emu_launch_cluster(ms::sd_train_cluster_kernel, a, ms::sdc::kCluster, ms::sdc::kCThreads);
}

// ms_sdcfr_sample_rows' signature with a host pointer
extern "C" int emu_ms_sdcfr_sample_rows(int* idx, int batch, int epochs, long long n_rows, unsigned long long seed,
                                        unsigned long long first_epoch, void* /*stream*/) {
    if (!idx || batch < 1 || batch > ms::sds::kSampleThreads || epochs < 0 || n_rows < batch || n_rows >= (1ll << 31)) return -2;
    if (epochs == 0) return 0;
    ms::SdSampleArgs a;
    a.idx = idx; a.batch = batch; a.epochs = epochs; a.n_rows = n_rows; a.seed = seed; a.first_epoch = first_epoch;
    return emu_launch_grid(ms::sd_sample_rows_kernel, a, epochs < 1024 ? epochs : 1024, ms::sds::kSampleThreads);
}
