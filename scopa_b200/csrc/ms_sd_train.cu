// scopa_b200/csrc/ms_sd_train.cu -- C ABI of the fused advantage-net optimiser (kernel in ms_sd_train.cuh).
#include <cmath>

#include "ms_common.cuh"
#include "ms_sd_train.cuh"

extern "C" {

size_t ms_sdcfr_train_workspace_bytes(void) { return (size_t)ms::sdt::kNetFloats * sizeof(float); }

int ms_sdcfr_train(float* d_net, float* d_adam_m, float* d_adam_v, int64_t steps_done, const float* d_feat,
                   const float* d_target, const float* d_mask, int64_t n_rows, const int32_t* d_idx, int32_t batch,
                   int32_t epochs, double lr, double beta1, double beta2, double eps, double max_norm, float* d_loss,
                   void* d_workspace, size_t workspace_bytes, void* stream) {
    using namespace ms;
    if (!d_net || !d_adam_m || !d_adam_v || !d_feat || !d_target || !d_mask || !d_idx || !d_loss || !d_workspace)
        return fail(MS_ERR_ARG, "ms_sdcfr_train: null pointer");
    if (batch < 1 || batch > sdt::kMaxBatch) return fail(MS_ERR_ARG, "ms_sdcfr_train: batch %d not in 1..128", batch);
    if (epochs < 0 || steps_done < 0 || n_rows < 1)
        return fail(MS_ERR_ARG, "ms_sdcfr_train: epochs %d, steps_done %lld, n_rows %lld", epochs, (long long)steps_done,
                    (long long)n_rows);
    if (!(lr > 0) || !(beta1 >= 0 && beta1 < 1) || !(beta2 >= 0 && beta2 < 1) || !(eps >= 0) || !(max_norm > 0))
        return fail(MS_ERR_ARG, "ms_sdcfr_train: lr %g, betas (%g, %g), eps %g, max_norm %g", lr, beta1, beta2, eps, max_norm);
    if (workspace_bytes < ms_sdcfr_train_workspace_bytes())
        return fail(MS_ERR_ARG, "ms_sdcfr_train: workspace of %zu bytes, need %zu", workspace_bytes,
                    ms_sdcfr_train_workspace_bytes());
    if (epochs == 0) return MS_OK;
    // per device (function attributes belong to the context), so set on every call like the other entry points
    MS_CUDA(cudaFuncSetAttribute(sd_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sdt::kSmemBytes));
    SdTrainArgs a;
    a.net = d_net; a.adam_m = d_adam_m; a.adam_v = d_adam_v;
    a.feat = d_feat; a.target = d_target; a.mask = d_mask; a.n_rows = n_rows;
    a.idx = d_idx; a.batch = batch; a.epochs = epochs;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm;
    a.b1pow = std::pow(beta1, (double)steps_done);
    a.b2pow = std::pow(beta2, (double)steps_done);
    a.loss = d_loss;
    a.grad = static_cast<float*>(d_workspace);
    sd_train_kernel<<<1, sdt::kThreads, sdt::kSmemBytes, static_cast<cudaStream_t>(stream)>>>(a);
    MS_LAUNCH_CHECK();
    return MS_OK;
}

}  // extern "C"
