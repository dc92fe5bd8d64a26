"""CPU restatement (numpy, float32) of the SDCFR optimiser step and average policy -- TEST INFRASTRUCTURE ONLY (imported
by tests/; the product never calls it).

Restates, with explicit matrix algebra instead of autograd:
  * AdvantageNetwork.train's step (/root/reference/src/algorithms/deep_cfr/deep_cfr.py:77-110): forward of the
    34 -> 128 -> 64 -> 16 ReLU MLP (nets.py:151-235, :296-331), loss = nn.MSELoss()(pred * mask, target * mask) (:101),
    clip_grad_norm_(max_norm=1.0) (:105), optim.Adam(lr=5e-4) (:47, torch defaults betas (0.9, 0.999), eps 1e-8);
  * StrategyBuffer.get_average_policy (:136-160) with positive_regret_policy (nets.py:93-101).
Pinned in tests/test_sd_train_emu.py against torch itself (the reference's own arithmetic) on seeded inputs; the
emulated CUDA kernels are then compared with both.  Parameters travel as ONE float32 blob in nn.Linear order
(w1[128][34] b1[128] w2[64][128] b2[64] w3[16][64] b3[16]), like the C ABI.
"""
import numpy as np

SHAPES = ((128, 34), (128,), (64, 128), (64,), (16, 64), (16,))
NET_FLOATS = 13776
f32 = np.float32


def split(blob):
    out, off = [], 0
    for shp in SHAPES:
        n = int(np.prod(shp))
        out.append(blob[off:off + n].reshape(shp))
        off += n
    return out


def forward(blob, x):
    """FlexibleNet(mode='mlp').forward: head(backbone(x)), MLPBlock = act(fc(x)) (nets.py:39-41, :60-61)."""
    w1, b1, w2, b2, w3, b3 = split(blob)
    h1 = np.maximum(x @ w1.T + b1, f32(0))
    h2 = np.maximum(h1 @ w2.T + b2, f32(0))
    return h1, h2, h2 @ w3.T + b3


def train_steps(blob, exp_avg, exp_avg_sq, steps_done, feat, target, mask, idx, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-8,
                max_norm=1.0):
    """In place on blob / exp_avg / exp_avg_sq; idx [epochs, batch] rows -> per-epoch losses (float32)."""
    losses = np.zeros(len(idx), f32)
    for ep, rows in enumerate(idx):
        x, t, m = feat[rows], target[rows], mask[rows]
        w1, b1, w2, b2, w3, b3 = split(blob)
        h1, h2, out = forward(blob, x)
        diff = out * m - t * m                                     # (:101) both operands masked
        n = f32(diff.size)                                         # MSELoss: mean over batch x 16
        losses[ep] = np.sum(diff * diff, dtype=f32) / n
        d_out = (f32(2) * diff / n) * m
        g_w3, g_b3 = d_out.T @ h2, d_out.sum(0, dtype=f32)
        d_h2 = (d_out @ w3) * (h2 > 0)
        g_w2, g_b2 = d_h2.T @ h1, d_h2.sum(0, dtype=f32)
        d_h1 = (d_h2 @ w2) * (h1 > 0)
        g_w1, g_b1 = d_h1.T @ x, d_h1.sum(0, dtype=f32)
        grad = np.concatenate([g.reshape(-1) for g in (g_w1, g_b1, g_w2, g_b2, g_w3, g_b3)]).astype(f32)
        # clip_grad_norm_: total_norm over all parameters, coef = max_norm / (norm + 1e-6) clamped to 1 (:105)
        norm = np.sqrt(np.sum(grad.astype(np.float64) ** 2)).astype(f32)
        grad = grad * min(f32(max_norm) / (norm + f32(1e-6)), f32(1.0))
        # torch.optim.Adam, no weight decay, no amsgrad: bias corrections computed in double from the step count
        step = steps_done + ep + 1
        exp_avg += (grad - exp_avg) * f32(1.0 - beta1)             # lerp_
        exp_avg_sq *= f32(beta2)
        exp_avg_sq += f32(1.0 - beta2) * grad * grad               # addcmul_
        step_size = f32(lr / (1.0 - beta1 ** step))
        bc2_sqrt = f32(np.sqrt(1.0 - beta2 ** step))
        blob -= step_size * (exp_avg / (np.sqrt(exp_avg_sq) / bc2_sqrt + f32(eps)))
    return losses


def positive_regret_policy(adv, mask, eps=1e-8):
    """nets.py:93-101."""
    pos = np.maximum(adv, f32(0)) * mask
    z = np.maximum(pos.sum(-1, keepdims=True, dtype=f32), f32(eps))
    return pos / z


def average_policy(blobs, weights, feat, mask):
    """StrategyBuffer.get_average_policy for a batch (:143-160): sum_k RM(net_k(x)) * (weight_k / total_weight)."""
    total = sum(weights)
    policy = np.zeros((feat.shape[0], 16), f32)
    for blob, w in zip(blobs, weights):
        policy += positive_regret_policy(forward(blob, feat)[2], mask) * f32(w / total)
    return policy
