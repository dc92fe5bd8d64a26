"""The PRODUCT's SDCFR kernels (fp32 path) on the CPU: sd_mlp_kernel<0> and the level-batched external-sampling
traversal of scopa_b200/csrc/ms_sdcfr.cu run by the CTA emulator (tests/emu/ms_sdcfr_host.cpp), against the fixture
recorded from the unmodified reference (tests/golden/sdcfr_seed0.npz: the trained nets of DeepCFR seed 0, features /
advantages / policies of every node of its traversals) and the oracle on the same Philox stream.  Same assertions as
tests/test_gpu_sdcfr.py, which runs the device build (and the tcgen05 path, which has no host form)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ms_oracle as ora
from scopa_b200 import codec

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_build  # noqa: E402

vp = C.c_void_p


@pytest.fixture(scope="module")
def sd():
    L = C.CDLL(emu_build.build_sdcfr_host())
    L.host_sd_mlp_forward.argtypes = [vp, vp, vp, vp, vp, C.c_longlong]
    L.host_sd_traverse.argtypes = [vp, C.c_uint32, C.c_int, vp, vp, C.c_longlong, C.c_ulonglong, C.c_ulonglong, vp, vp, vp, vp]
    return L


@pytest.fixture(scope="module")
def nets():
    g = np.load(os.path.join(GOLDEN, "sdcfr_seed0.npz"))
    arrs = [[g[f"net{p}.backbone.0.fc.weight"], g[f"net{p}.backbone.0.fc.bias"], g[f"net{p}.backbone.1.fc.weight"],
             g[f"net{p}.backbone.1.fc.bias"], g[f"net{p}.head.weight"], g[f"net{p}.head.bias"]] for p in range(2)]
    blobs = [np.concatenate([np.asarray(a, dtype=np.float32).reshape(-1) for a in ar]) for ar in arrs]   # nn.Linear order
    return g, blobs, [ora.Mlp(*a) for a in arrs]


def mlp_forward(sd, blob, feat, mask):
    feat, mask = np.ascontiguousarray(feat, np.float32), np.ascontiguousarray(mask, np.float32)
    n = len(feat)
    adv, pol = np.zeros((n, 16), np.float32), np.zeros((n, 16), np.float32)
    assert sd.host_sd_mlp_forward(blob.ctypes.data, feat.ctypes.data, mask.ctypes.data, adv.ctypes.data, pol.ctypes.data, n) == 0
    return adv, pol


def test_mlp_forward_fp32_matches_reference_and_oracle(sd, nets):
    g, blobs, omlps = nets
    hist = g["node_hist"]
    cp = np.array([(h >= 0).sum() & 1 for h in hist])
    for p in (0, 1):
        rows = np.nonzero(cp == p)[0]
        adv, pol = mlp_forward(sd, blobs[p], g["node_feat"][rows], g["node_mask"][rows])
        # the reference (torch CPU sgemm, another summation order): fp32 tolerance
        np.testing.assert_allclose(adv, g["node_adv"][rows], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(pol, g["node_pol"][rows], rtol=1e-3, atol=2e-5)
        # the oracle: same summation order, separate mul / add -> identical bits
        for i in rows[::25]:
            oa, op = ora.advantages_policy(omlps[p], g["node_feat"][i], g["node_mask"][i])
            k = int(np.nonzero(rows == i)[0][0])
            assert np.array_equal(adv[k], oa) and np.array_equal(pol[k], op)
    for n in (1, 127, 129):                                       # ragged tiles
        adv, _ = mlp_forward(sd, blobs[0], g["node_feat"][:n], g["node_mask"][:n])
        full, _ = mlp_forward(sd, blobs[0], g["node_feat"][:300], g["node_mask"][:300])
        assert np.array_equal(adv, full[:n])


def _sorted_rows(*arrs):
    m = np.concatenate([np.asarray(a, dtype=np.float64) for a in arrs], axis=1)
    return m[np.lexsort(m.T[::-1])]


@pytest.mark.parametrize("player", [0, 1])
def test_traversal_fp32_matches_oracle_on_the_same_stream(sd, nets, player):
    """DeepCFR._external_sampling_cfr (deep_cfr.py:284-365) for 300 traversals at once: the same nodes visited (same
    sampled actions), the same samples (features, masks, normalised regrets) and root values as the oracle's recursion."""
    g, blobs, omlps = nets
    cards = ora.deck(42)
    root = np.array(codec.pack_state([codec.mask_of(cards[:4]), codec.mask_of(cards[4:8])], [], [0, 0], [0, 0], 0, 0, False, 8),
                    dtype=np.uint32)
    ho = codec.pack_nibbles(cards[:8])
    n, per = 300, 41
    feat, target, mask = np.zeros((n * per, 34), np.float32), np.zeros((n * per, 16), np.float32), np.zeros((n * per, 16), np.float32)
    value = np.zeros(n, np.float32)
    assert sd.host_sd_samples_per_traversal(player) == per
    assert sd.host_sd_traverse(root.ctypes.data, ho, player, blobs[0].ctypes.data, blobs[1].ctypes.data, n, 77, 5,
                               feat.ctypes.data, target.ctypes.data, mask.ctypes.data, value.ctypes.data) == 0
    rng = ora.Rng(1, 77)
    for t in range(0, n, 7):
        v, of, ot, om = ora.sdcfr_traverse(omlps, player, rng, trav_id=5 + t)
        sl = slice(t * per, (t + 1) * per)
        assert len(of) == per
        got, want = _sorted_rows(feat[sl], mask[sl], target[sl]), _sorted_rows(of, om, ot)
        assert np.array_equal(got[:, :50], want[:, :50]), t          # same nodes visited (same sampled actions)
        np.testing.assert_allclose(got[:, 50:], want[:, 50:], rtol=1e-6, atol=1e-7)
        assert abs(value[t] - v) < 1e-6
    assert np.all(feat[:, 32] == 1) and np.all(feat[:, 33] == 0) and np.array_equal(mask, feat[:, :16])
    assert np.isfinite(target).all() and np.abs(target).max() <= 1.0 + 1e-6
