// scopa_b200/csrc/ms_sd_train.cu -- C ABI of the fused advantage-net optimiser (kernel in ms_sd_train.cuh) and of the
// all-nets average policy (kernels in ms_sd_avgpol.cuh).
#include <cmath>

#include "ms_common.cuh"
#include "ms_sd_avgpol.cuh"
#include "ms_sd_sample.cuh"
#include "ms_sd_train.cuh"
#include "ms_sd_train_cluster.cuh"

extern "C" {

size_t ms_sdcfr_train_workspace_bytes(void) { return (size_t)ms::sdt::kNetFloats * sizeof(float); }

static int sd_train_args(const char* who, float* d_net, float* d_adam_m, float* d_adam_v, int64_t steps_done,
                         const float* d_feat, const float* d_target, const float* d_mask, int64_t n_rows, const int32_t* d_idx,
                         int32_t batch, int32_t epochs, double lr, double beta1, double beta2, double eps, double max_norm,
                         float* d_loss, void* d_workspace, size_t workspace_bytes, ms::SdTrainArgs* out) {
    using namespace ms;
    if (!d_net || !d_adam_m || !d_adam_v || !d_feat || !d_target || !d_mask || !d_idx || !d_loss || !d_workspace)
        return fail(MS_ERR_ARG, "%s: null pointer", who);
    if (batch < 1 || batch > sdt::kMaxBatch) return fail(MS_ERR_ARG, "%s: batch %d not in 1..128", who, batch);
    if (epochs < 0 || steps_done < 0 || n_rows < 1)
        return fail(MS_ERR_ARG, "%s: epochs %d, steps_done %lld, n_rows %lld", who, epochs, (long long)steps_done,
                    (long long)n_rows);
    if (!(lr > 0) || !(beta1 >= 0 && beta1 < 1) || !(beta2 >= 0 && beta2 < 1) || !(eps >= 0) || !(max_norm > 0))
        return fail(MS_ERR_ARG, "%s: lr %g, betas (%g, %g), eps %g, max_norm %g", who, lr, beta1, beta2, eps, max_norm);
    if (workspace_bytes < ms_sdcfr_train_workspace_bytes())
        return fail(MS_ERR_ARG, "%s: workspace of %zu bytes, need %zu", who, workspace_bytes, ms_sdcfr_train_workspace_bytes());
    SdTrainArgs& a = *out;
    a.net = d_net; a.adam_m = d_adam_m; a.adam_v = d_adam_v;
    a.feat = d_feat; a.target = d_target; a.mask = d_mask; a.n_rows = n_rows;
    a.idx = d_idx; a.batch = batch; a.epochs = epochs;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
    a.b1pow = std::pow(beta1, (double)steps_done);
    a.b2pow = std::pow(beta2, (double)steps_done);
    a.loss = d_loss;
    a.grad = static_cast<float*>(d_workspace);
    return MS_OK;
}

int ms_sdcfr_train(float* d_net, float* d_adam_m, float* d_adam_v, int64_t steps_done, const float* d_feat,
                   const float* d_target, const float* d_mask, int64_t n_rows, const int32_t* d_idx, int32_t batch,
                   int32_t epochs, double lr, double beta1, double beta2, double eps, double max_norm, float* d_loss,
                   void* d_workspace, size_t workspace_bytes, void* stream) {
    using namespace ms;
    SdTrainArgs a;
    int rc = sd_train_args("ms_sdcfr_train", d_net, d_adam_m, d_adam_v, steps_done, d_feat, d_target, d_mask, n_rows, d_idx,
                           batch, epochs, lr, beta1, beta2, eps, max_norm, d_loss, d_workspace, workspace_bytes, &a);
    if (rc) return rc;
    if (epochs == 0) return MS_OK;
    // per device (function attributes belong to the context), so set on every call like the other entry points
    MS_CUDA(cudaFuncSetAttribute(sd_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sdt::kSmemBytes));
    sd_train_kernel<<<1, sdt::kThreads, sdt::kSmemBytes, static_cast<cudaStream_t>(stream)>>>(a);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_sdcfr_train_cluster(float* d_net, float* d_adam_m, float* d_adam_v, int64_t steps_done, const float* d_feat,
                           const float* d_target, const float* d_mask, int64_t n_rows, const int32_t* d_idx, int32_t batch,
                           int32_t epochs, double lr, double beta1, double beta2, double eps, double max_norm,
                           float* d_loss, void* d_workspace, size_t workspace_bytes, void* stream) {
    using namespace ms;
    SdTrainArgs a;
    int rc = sd_train_args("ms_sdcfr_train_cluster", d_net, d_adam_m, d_adam_v, steps_done, d_feat, d_target, d_mask, n_rows,
                           d_idx, batch, epochs, lr, beta1, beta2, eps, max_norm, d_loss, d_workspace, workspace_bytes, &a);
    if (rc) return rc;
    if (epochs == 0) return MS_OK;
    MS_CUDA(cudaFuncSetAttribute(sd_train_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sdc::kCSmemBytes));
    // one cluster of 8 CTAs (compile-time __cluster_dims__); the partial gradients live in shared memory, d_workspace is unused
    sd_train_cluster_kernel<<<sdc::kCluster, sdc::kCThreads, sdc::kCSmemBytes, static_cast<cudaStream_t>(stream)>>>(a);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

size_t ms_sdcfr_average_policy_workspace_bytes(int32_t n_nets, int64_t n_rows) {
    if (n_nets < 0 || n_rows < 0) return 0;
    return (size_t)n_nets * (size_t)n_rows * ms::sdt::kOut * sizeof(float);
}

int ms_sdcfr_average_policy(const float* d_nets, const float* d_weights, int32_t n_nets, const float* d_feat,
                            const float* d_mask, int64_t n_rows, float* d_policy, void* d_workspace,
                            size_t workspace_bytes, void* stream) {
    using namespace ms;
    if (n_nets < 1 || n_rows < 0) return fail(MS_ERR_ARG, "ms_sdcfr_average_policy: n_nets %d, n_rows %lld", n_nets, (long long)n_rows);
    if (n_rows == 0) return MS_OK;
    if (!d_nets || !d_weights || !d_feat || !d_mask || !d_policy || !d_workspace)
        return fail(MS_ERR_ARG, "ms_sdcfr_average_policy: null pointer");
    if (workspace_bytes < ms_sdcfr_average_policy_workspace_bytes(n_nets, n_rows))
        return fail(MS_ERR_ARG, "ms_sdcfr_average_policy: workspace of %zu bytes, need %zu", workspace_bytes,
                    ms_sdcfr_average_policy_workspace_bytes(n_nets, n_rows));
    SdAvgPolArgs a;
    a.nets = d_nets; a.weights = d_weights; a.n_nets = n_nets;
    a.feat = d_feat; a.mask = d_mask; a.n_rows = n_rows;
    a.scratch = static_cast<float*>(d_workspace);
    a.policy = d_policy;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MS_CUDA(cudaFuncSetAttribute(sd_avgpol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sda::kPolSmemBytes));
    // grid.x = nets (one net per CTA, 1 CTA per SM at 119 KB of shared memory); grid.y splits a net's 64-row chunks over
    // more CTAs until about two waves of the 148 SMs are in flight (matters while the buffer holds few nets)
    const int gx = n_nets < kNumSMs ? n_nets : kNumSMs;
    const int64_t n_chunks = (n_rows + sda::kRows - 1) / sda::kRows;
    int64_t gy = (2 * kNumSMs + gx - 1) / gx;
    if (gy > n_chunks) gy = n_chunks;
    if (gy > 65535) gy = 65535;
    sd_avgpol_kernel<<<dim3((unsigned)gx, (unsigned)gy), sda::kPolThreads, sda::kPolSmemBytes, st>>>(a);
    MS_LAUNCH_CHECK();
    sd_avgpol_reduce_kernel<<<grid_for(n_rows * sdt::kOut, 256, 8), 256, 0, st>>>(a);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

int ms_sdcfr_sample_rows(int32_t* d_idx, int32_t batch, int32_t epochs, int64_t n_rows, uint64_t seed, uint64_t first_epoch,
                         void* stream) {
    using namespace ms;
    if (!d_idx) return fail(MS_ERR_ARG, "ms_sdcfr_sample_rows: null pointer");
    if (batch < 1 || batch > sds::kSampleThreads || epochs < 0)
        return fail(MS_ERR_ARG, "ms_sdcfr_sample_rows: batch %d (1..128), epochs %d", batch, epochs);
    if (n_rows < batch || n_rows >= (int64_t)1 << 31)
        return fail(MS_ERR_ARG, "ms_sdcfr_sample_rows: n_rows %lld must lie in [batch, 2^31)", (long long)n_rows);
    if (epochs == 0) return MS_OK;
    SdSampleArgs a;
    a.idx = d_idx; a.batch = batch; a.epochs = epochs; a.n_rows = n_rows; a.seed = seed; a.first_epoch = first_epoch;
    sd_sample_rows_kernel<<<epochs < 1024 ? epochs : 1024, sds::kSampleThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

}  // extern "C"
