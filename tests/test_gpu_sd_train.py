"""ms_sdcfr_train (sd_train_kernel, csrc/ms_sd_train.cuh) on the device.

1. Bit parity: tests/emu/sd_train_check runs a seeded problem through the C ABI and through the host emulation of the
   same kernel source; parameters, Adam moments and losses must be identical words.  (The emulation is pinned to the
   reference's arithmetic -- torch -- in tests/test_sd_train_emu.py.)
2. The Python binding on CUDA tensors against torch.optim.Adam on the same minibatches (fp32 tolerance: torch's sgemm
   summation order differs).
3. DeepCFR(optimizer="fused") trains through the drop-in API.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_build  # noqa: E402

from scopa_b200 import sdcfr  # noqa: E402
from scopa_b200.algorithms.deep_cfr.deep_cfr import AdvantageNetwork, DeepCFR  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch,epochs,n_rows", [(128, 6, 4096), (128, 1, 300), (32, 5, 100), (7, 3, 7), (1, 2, 1), (100, 4, 300)])
def test_device_kernel_equals_emulation_bit_for_bit(batch, epochs, n_rows):
    from scopa_b200 import _lib
    _lib.load()                                       # libscopa_b200.so must exist before the checker links against it
    exe = emu_build.build_check()
    res = subprocess.run([exe, str(batch), str(epochs), str(n_rows), "0"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "differing words net 0 m 0 v 0 loss 0" in res.stdout, res.stdout


def test_binding_matches_torch_adam_on_cuda():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    fused, plain = AdvantageNetwork(34, 16, device="cuda", optimizer="fused"), AdvantageNetwork(34, 16, device="cuda")
    plain.net.load_state_dict(fused.net.state_dict())
    n = 3000
    feat = torch.from_numpy((rng.random((n, 34)) < 0.25).astype(np.float32)).cuda()
    mask = torch.zeros((n, 16), device="cuda")
    mask[torch.arange(n, device="cuda"), torch.from_numpy(rng.integers(0, 16, n)).cuda()] = 1
    mask[torch.arange(n, device="cuda"), torch.from_numpy(rng.integers(0, 16, n)).cuda()] = 1
    target = torch.from_numpy(rng.uniform(-1, 1, (n, 16)).astype(np.float32)).cuda() * mask
    for adv in (fused, plain):
        adv.buffer.add_batch(feat, target, mask)
    rows = fused._sample_rows(128, 8)
    assert rows.is_cuda and rows.shape == (8, 128)
    fused._sample_rows = lambda batch_size, epochs: rows
    it = iter(rows.long())

    def same_rows(batch_size):
        r = next(it)
        return plain.buffer.feat[r], plain.buffer.target[r], plain.buffer.mask[r]

    plain.buffer.sample = same_rows
    before = sdcfr.flatten_net(plain.net).clone()
    lf, lp = fused.train(batch_size=128, epochs=8), plain.train(batch_size=128, epochs=8)
    assert abs(lf - lp) < 1e-5 * abs(lp), (lf, lp)
    a, b = sdcfr.flatten_net(fused.net), sdcfr.flatten_net(plain.net)
    assert float((b - before).abs().max()) > 1e-3                     # eight Adam steps of 5e-4
    # Adam turns a gradient into a step of about lr * g / (|g| + eps): an element whose gradient is within fp32 noise of
    # zero (~1e-9 here) can take a different step under another summation order (cuBLAS vs our k-ascending fmaf chain),
    # so: essentially all parameters agree to 5e-6, stragglers are bounded by the total movement of eight steps
    diff = (a - b).abs()
    assert float(diff.median()) < 1e-7 and float((diff > 5e-6).float().mean()) < 2e-3 and float(diff.max()) < 8 * 5e-4 + 1e-5
    assert fused._fused.steps_done == 8 and fused.blob() is fused._fused.blob
    # inference reads the blob the optimiser just updated
    adv_f, _ = sdcfr.mlp_forward(fused.blob(), feat[:64], mask[:64], sdcfr.FP32)
    adv_p, _ = sdcfr.mlp_forward(plain.blob(), feat[:64], mask[:64], sdcfr.FP32)
    assert float(((adv_f - adv_p) * mask[:64]).abs().max()) < 5e-3


def test_bad_rows_are_reported_not_applied():
    torch.manual_seed(1)
    net = AdvantageNetwork(34, 16, device="cuda", optimizer="fused")
    opt = net._fused
    feat, target, mask = (torch.rand((10, w), device="cuda") for w in (34, 16, 16))
    idx = torch.tensor([[0, 1, 2, 3], [4, 5, 10, 6]], dtype=torch.int32, device="cuda")      # row 10 does not exist
    w0 = opt.blob.clone()
    loss = opt.step(feat, target, mask, idx[1:].contiguous())
    assert torch.isnan(loss).all() and torch.equal(opt.blob, w0)
    loss = opt.step(feat, target, mask, idx[:1].contiguous())
    assert torch.isfinite(loss).all() and not torch.equal(opt.blob, w0)


def test_deepcfr_trains_with_the_fused_optimiser():
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401  (registers mini_scopa)
    torch.manual_seed(0)
    game = pyspiel.load_game("mini_scopa")
    d = DeepCFR(game, device="cuda", traversals_per_iteration=256, seed=3, optimizer="fused")
    w0 = [a.blob().clone() for a in d.advantage_nets]
    d.train(iterations=3, advantage_epochs=6, eval_freq=10, eval_episodes=0)
    for p in (0, 1):
        assert len(d.training_history["losses"][p]) == 3 and all(np.isfinite(d.training_history["losses"][p]))
        assert d.advantage_nets[p]._fused.steps_done == 18
        assert not torch.equal(d.advantage_nets[p].blob(), w0[p])
        assert len(d.strategy_buffers[p].strategies) == 2
        # snapshots are copies, not views of the live blob
        snap = sdcfr.flatten_net(d.strategy_buffers[p].strategies[0])
        assert snap.data_ptr() != d.advantage_nets[p].blob().data_ptr()
    s = game.new_initial_state()
    pol = d.get_policy(s, 0)
    assert pol.shape == (16,) and (pol >= 0).all() and pol.sum() < 1 + 1e-5


@pytest.mark.parametrize("n_nets,n_rows", [(3, 1), (8, 70), (20, 200), (2, 300)])
def test_average_policy_kernels_equal_emulation_bit_for_bit(n_nets, n_rows):
    from scopa_b200 import _lib
    _lib.load()
    exe = emu_build.build_check()
    res = subprocess.run([exe, "avgpol", str(n_nets), str(n_rows), "0"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "differing words 0 of" in res.stdout, res.stdout


def test_strategy_buffer_average_policy_on_cuda():
    """StrategyBuffer.get_average_policy / average_policy_batch (two launches for all nets) against the reference's loop
    of per-net forwards, here torch on the same device."""
    from scopa_b200.algorithms.deep_cfr.deep_cfr import HIDDEN, StrategyBuffer
    from scopa_b200.algorithms.deep_cfr.nets import FlexibleNet, positive_regret_policy
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    buf = StrategyBuffer(max_size=100)
    for it in range(1, 13):
        net = FlexibleNet(mode="mlp", input_shape=(34,), output_dim=16, mlp_hidden=HIDDEN, mlp_act="relu", mlp_norm="none").cuda()
        for l in net.modules():
            if isinstance(l, nn.Linear):
                nn.init.xavier_uniform_(l.weight)
                nn.init.normal_(l.bias, 0.0, 0.3)
        buf.add_strategy(net, it)
    n = 200
    feat = (rng.random((n, 34)) < 0.3).astype(np.float32)
    mask = np.zeros((n, 16), np.float32)
    for r in range(n):
        mask[r, rng.choice(16, rng.integers(1, 5), replace=False)] = 1
    x, m = torch.from_numpy(feat).cuda(), torch.from_numpy(mask).cuda()
    want = torch.zeros_like(m)
    tot = sum(buf.weights)
    with torch.no_grad():
        for net, w in zip(buf.strategies, buf.weights):
            want += positive_regret_policy(net(x), m) * (w / tot)
    got = buf.average_policy_batch(x, m)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)       # wiring check; bit parity is the test above
    one = buf.get_average_policy(feat[3], mask[3])
    assert one.dtype == np.float32 and one.shape == (16,) and np.all(one[mask[3] == 0] == 0)
    np.testing.assert_allclose(one, want[3].cpu().numpy(), rtol=1e-4, atol=1e-5)
    # the cache follows the lists
    buf.strategies.pop(0)
    buf.weights.pop(0)
    got2 = buf.average_policy_batch(x, m)
    assert not torch.equal(got, got2) and len(buf._blob_of) == 11


def test_device_evaluation_agrees_with_the_episode_loop():
    """DeepCFR(device_eval=True).evaluate_vs_random: all episodes in two launches of the policy-evaluation kernel against
    the reference-shaped episode loop (same average policy, independent random streams).  Rewards lie in [-4.5, 4.5]
    with a per-game standard deviation of about 1.5-2.5: 20 000 device episodes vs 1500 looped ones differ by < 0.35 (> 5 sigma)."""
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401
    torch.manual_seed(2)
    np.random.seed(2)
    game = pyspiel.load_game("mini_scopa")
    d = DeepCFR(game, device="cuda", traversals_per_iteration=512, seed=5, optimizer="fused", device_eval=True)
    r0, sc0 = d.evaluate_vs_random(1000)                       # no strategies yet: uniform vs uniform
    assert abs(r0) < 0.25 and len(sc0) == 2
    d.train(iterations=4, advantage_epochs=8, eval_freq=2, eval_episodes=200)
    assert len(d.training_history["eval_rewards"]) == 1 + 2    # the call above + iterations 0 and 2
    r_dev, sc_dev = d.evaluate_vs_random(20000)
    r_loop, sc_loop = d.evaluate_vs_random(1500, on_device=False)
    assert abs(r_dev - r_loop) < 0.35, (r_dev, r_loop)
    assert abs(sc_dev[0] - sc_loop[0]) < 0.15 and abs(sc_dev[1] - sc_loop[1]) < 0.15, (sc_dev, sc_loop)
    r1, _ = d.evaluate_vs_random(1)                            # odd / tiny counts: second half empty
    assert -4.5 <= r1 <= 4.5


# The cluster form of the optimiser, the Philox row sampler and the 2-D average-policy grid first ran on a device in round 2
# (profiles/prof_r02b.sh: bit parity with their emulations, then timing); SCOPA_B200_SKIP_CLUSTER=1 leaves them out.
unverified = pytest.mark.skipif(os.environ.get("SCOPA_B200_SKIP_CLUSTER") == "1", reason="SCOPA_B200_SKIP_CLUSTER=1")


@unverified
@pytest.mark.parametrize("batch,epochs,n_rows", [(128, 6, 4096), (32, 5, 100), (17, 3, 40), (1, 2, 1), (100, 4, 300)])
def test_cluster_kernel_equals_emulation_bit_for_bit(batch, epochs, n_rows):
    from scopa_b200 import _lib
    _lib.load()
    exe = emu_build.build_check()
    res = subprocess.run([exe, "cluster", str(batch), str(epochs), str(n_rows), "0"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "differing words net 0 m 0 v 0 loss 0" in res.stdout, res.stdout


@unverified
def test_deepcfr_trains_with_the_cluster_optimiser():
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401
    torch.manual_seed(0)
    game = pyspiel.load_game("mini_scopa")
    d = DeepCFR(game, device="cuda", traversals_per_iteration=256, seed=3, optimizer="fused-cluster")
    w0 = [a.blob().clone() for a in d.advantage_nets]
    d.train(iterations=3, advantage_epochs=6, eval_freq=10, eval_episodes=0)
    for p in (0, 1):
        assert all(np.isfinite(d.training_history["losses"][p])) and d.advantage_nets[p]._fused.steps_done == 18
        assert not torch.equal(d.advantage_nets[p].blob(), w0[p])


@unverified
@pytest.mark.parametrize("batch,epochs,n_rows", [(128, 10, 100000), (128, 3, 130), (32, 4, 41), (32, 3, 32), (1, 3, 1)])
def test_sampler_kernel_equals_emulation(batch, epochs, n_rows):
    """ms_sdcfr_sample_rows (written after the GPU budget was spent, like the cluster kernel): device rows == emulated rows
    (the emulation equals the draw-for-draw restatement in tests/test_sd_train_emu.py)."""
    from scopa_b200 import _lib
    _lib.load()
    exe = emu_build.build_check()
    res = subprocess.run([exe, "sample", str(batch), str(epochs), str(n_rows), "0"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "differing words 0 of" in res.stdout, res.stdout


@unverified
def test_deepcfr_with_the_philox_sampler_is_repeatable():
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401
    game = pyspiel.load_game("mini_scopa")
    blobs = []
    for _ in range(2):
        torch.manual_seed(0)
        d = DeepCFR(game, device="cuda", traversals_per_iteration=64, seed=3, optimizer="fused", sampler_seed=9)
        d.train(iterations=3, advantage_epochs=4, eval_freq=10, eval_episodes=0)
        blobs.append([a.blob().clone() for a in d.advantage_nets])
    assert all(torch.equal(x, y) for x, y in zip(*blobs))
