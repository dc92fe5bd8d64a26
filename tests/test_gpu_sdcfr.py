"""SDCFR kernels against the oracle / reference fixtures: batched advantage-net inference (fp32 CUDA-core
parity path and bf16 tensor-core path) and the level-batched external-sampling traversal."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import ms_oracle as ora
from scopa_b200 import sdcfr
from scopa_b200.solver import deal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nets():
    g = np.load(os.path.join(GOLDEN, "sdcfr_seed0.npz"))
    arrs = [[g[f"net{p}.backbone.0.fc.weight"], g[f"net{p}.backbone.0.fc.bias"], g[f"net{p}.backbone.1.fc.weight"],
             g[f"net{p}.backbone.1.fc.bias"], g[f"net{p}.head.weight"], g[f"net{p}.head.bias"]] for p in range(2)]
    return g, [sdcfr.blob_from_arrays(*a) for a in arrs], [ora.Mlp(*a) for a in arrs]


def test_mlp_forward_fp32_matches_reference_and_oracle(nets):
    g, blobs, omlps = nets
    hist = g["node_hist"]
    cp = np.array([(h >= 0).sum() & 1 for h in hist])
    for p in (0, 1):
        rows = np.nonzero(cp == p)[0]
        feat = torch.from_numpy(g["node_feat"][rows]).cuda()
        mask = torch.from_numpy(g["node_mask"][rows]).cuda()
        adv, pol = sdcfr.mlp_forward(blobs[p], feat, mask, sdcfr.FP32)
        adv, pol = adv.cpu().numpy(), pol.cpu().numpy()
        # reference (torch CPU sgemm, different summation order): fp32 tolerance
        np.testing.assert_allclose(adv, g["node_adv"][rows], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(pol, g["node_pol"][rows], rtol=1e-3, atol=2e-5)
        # oracle: same summation order, separate mul/add -> identical bits
        for i in rows[::25]:
            oa, op = ora.advantages_policy(omlps[p], g["node_feat"][i], g["node_mask"][i])
            k = int(np.nonzero(rows == i)[0][0])
            assert np.array_equal(adv[k], oa) and np.array_equal(pol[k], op)


def test_mlp_forward_tensor_core_path(nets):
    """tcgen05 path: bf16 operands, fp32 accumulate.  Stated tolerance: |adv_tc - adv_fp32| <= 0.03 on
    advantages of magnitude <= 1 (bf16 has 8 mantissa bits; three chained layers)."""
    g, blobs, _ = nets
    rng = np.random.default_rng(0)
    for n in (1, 127, 128, 129, 5000):
        idx = rng.integers(0, len(g["node_feat"]), n)
        feat = torch.from_numpy(g["node_feat"][idx]).cuda()
        mask = torch.from_numpy(g["node_mask"][idx]).cuda()
        a32, _ = sdcfr.mlp_forward(blobs[0], feat, mask, sdcfr.FP32)
        atc, ptc = sdcfr.mlp_forward(blobs[0], feat, mask, sdcfr.TENSOR_CORE)
        legal = mask > 0
        err = (a32 - atc).abs()[legal].max().item()
        assert err < 0.03, (n, err)
        assert torch.all(atc[~legal] == -1e6)
        s = ptc.sum(1)
        assert torch.all((s < 1e-6) | ((s - 1).abs() < 1e-4))
    # general (non 0/1) features also go through the bf16 path
    feat = torch.randn(300, 34, device="cuda") * 0.5
    mask = torch.ones(300, 16, device="cuda")
    a32, _ = sdcfr.mlp_forward(blobs[1], feat, mask, sdcfr.FP32)
    atc, _ = sdcfr.mlp_forward(blobs[1], feat, mask, sdcfr.TENSOR_CORE)
    assert (a32 - atc).abs().max().item() < 0.06


def _sorted_rows(*arrs):
    m = np.concatenate([np.asarray(a, dtype=np.float64) for a in arrs], axis=1)
    return m[np.lexsort(m.T[::-1])]


@pytest.mark.parametrize("player", [0, 1])
def test_traversal_fp32_matches_oracle_on_the_same_stream(nets, player):
    g, blobs, omlps = nets
    root, ho = deal(42)
    tr = sdcfr.Traverser(root, ho)
    n = 300
    feat, target, mask, value = tr.run(player, blobs, n, philox_seed=77, first_trav=5, precision=sdcfr.FP32)
    feat, target, mask, value = (t.cpu().numpy() for t in (feat, target, mask, value))
    assert feat.shape == (n * 41, 34)
    rng = ora.Rng(1, 77)
    for t in range(0, n, 7):
        v, of, ot, om = ora.sdcfr_traverse(omlps, player, rng, trav_id=5 + t)
        sl = slice(t * 41, (t + 1) * 41)
        assert len(of) == 41
        got = _sorted_rows(feat[sl], mask[sl], target[sl])
        want = _sorted_rows(of, om, ot)
        assert np.array_equal(got[:, :50], want[:, :50]), t          # same nodes visited (same sampled actions)
        np.testing.assert_allclose(got[:, 50:], want[:, 50:], rtol=1e-6, atol=1e-7)
        assert abs(value[t] - v) < 1e-6


def test_traversal_tensor_core_path_is_consistent(nets):
    g, blobs, _ = nets
    root, ho = deal(42)
    tr = sdcfr.Traverser(root, ho)
    n = 2048
    f32 = tr.run(0, blobs, n, philox_seed=3, precision=sdcfr.FP32)
    f32 = [t.clone() for t in f32]
    ftc = tr.run(0, blobs, n, philox_seed=3, precision=sdcfr.TENSOR_CORE)
    feat, target, mask, value = ftc
    assert torch.isfinite(target).all() and target.abs().max().item() <= 1.0 + 1e-6
    assert torch.all(feat[:, 32] == 1) and torch.all(feat[:, 33] == 0)
    assert torch.equal(mask, feat[:, :16])                      # legal mask == the mover's hand one-hot
    # the root sample (slot 0 of every traversal) sees the same state in both paths
    assert torch.equal(feat[0::41], f32[0][0::41])
    # bf16 rounding can flip a near-zero advantage, which changes the sampled path; most traversals agree
    same = (value - f32[3]).abs() < 0.05
    assert same.float().mean().item() > 0.8


@pytest.mark.parametrize("n", [1, 100, 128, 129, 4000])
def test_level_inference_kernel_equals_mlp_forward(nets, n):
    """sd_level_mlp_kernel (the inference half of a traversal level: packed states in, 16 raw outputs out) against
    ms_mlp_forward on the features of the same states, bit for bit on both paths: fp32 (same summation order) and tcgen05
    (the operand built from the state's bit masks equals the one converted from 0 / 1 floats; same MMAs)."""
    from scopa_b200.batch import BatchedMiniScopa
    g, blobs, _ = nets
    b = BatchedMiniScopa("cuda").reset(np.arange(1, n + 1, dtype=np.int64))
    rng = np.random.default_rng(n)
    plies = int(rng.integers(0, 4)) * 2                       # an even number of random plies: player 0 to move again
    for _ in range(plies):
        _, ordered, count = b.legal_actions()
        legal, cnt = ordered.cpu().numpy(), count.cpu().numpy()
        pick = np.array([legal[i, rng.integers(0, max(1, cnt[i]))] for i in range(n)], dtype=np.uint8)
        b.step(torch.from_numpy(pick).cuda())
    states = b.states.clone()
    w = states.cpu().numpy().view(np.uint32).reshape(n, 4)
    hand0 = w[:, 0] & 0xFFFF
    tlen = w[:, 3] & 0xF
    feat = np.zeros((n, 34), dtype=np.float32)
    for i in range(n):
        for c in range(16):
            feat[i, c] = (hand0[i] >> c) & 1
        for j in range(int(tlen[i])):
            feat[i, 16 + ((int(w[i, 1]) >> (4 * j)) & 0xF)] = 1.0
    feat[:, 32] = 1.0
    f_t = torch.from_numpy(feat).cuda()
    ones = torch.ones((n, 16), device="cuda")
    a32, _ = sdcfr.mlp_forward(blobs[0], f_t, ones, sdcfr.FP32)      # mask of ones: advantages == raw outputs
    atc, _ = sdcfr.mlp_forward(blobs[0], f_t, ones, sdcfr.TENSOR_CORE)
    r32 = sdcfr.infer_states(blobs[0], states, 0, sdcfr.FP32)
    assert torch.equal(r32, a32)
    rtc = sdcfr.infer_states(blobs[0], states, 0, sdcfr.TENSOR_CORE)
    assert torch.equal(rtc, atc)
    assert (rtc - r32).abs().max().item() < 0.03
