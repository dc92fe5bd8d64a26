"""sd_train_kernel (scopa_b200/csrc/ms_sd_train.cuh), executed on the HOST through tests/emu/cta_emu.h, against the
reference's own arithmetic: torch (CPU) running AdvantageNetwork.train's step -- MSELoss(pred * mask, target * mask),
clip_grad_norm_(1.0), Adam(5e-4) (/root/reference/src/algorithms/deep_cfr/deep_cfr.py:77-110).  torch is the oracle here
(the reference *is* torch code); its sgemm summation order is unspecified, hence tolerances: 1e-6 relative on the loss,
1e-6 absolute on parameters that move by ~5e-4 per step.  The GPU build of the same source is compared with this
emulation bit for bit in tests/test_gpu_sd_train.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_build  # noqa: E402

NF = 13776
fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(emu_build.build_emu())
    lib.emu_sd_train.argtypes = [fp, fp, fp, C.c_longlong, fp, fp, fp, C.c_longlong, ip, C.c_int, C.c_int] + [C.c_double] * 5 + [fp, fp]
    lib.emu_sd_train.restype = C.c_int
    return lib


def P(a):
    return a.ctypes.data_as(fp)


def make_net(seed):
    torch.manual_seed(seed)
    net = nn.Sequential(nn.Linear(34, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 16))
    for l in net:
        if isinstance(l, nn.Linear):               # AdvantageNetwork's init (deep_cfr.py:40-44)
            nn.init.xavier_uniform_(l.weight)
            nn.init.constant_(l.bias, 0.1)
    return net


def blob_of(net):
    return np.concatenate([p.detach().numpy().reshape(-1) for p in net.parameters()]).astype(np.float32).copy()


def make_problem(rng, n_rows, batch, epochs, scale=1.0):
    feat = (rng.random((n_rows, 34)) < 0.25).astype(np.float32)
    mask = np.zeros((n_rows, 16), np.float32)
    for r in range(n_rows):
        mask[r, rng.choice(16, rng.integers(1, 5), replace=False)] = 1
    target = (rng.uniform(-1, 1, (n_rows, 16)) * mask * scale).astype(np.float32)
    idx = np.stack([rng.choice(n_rows, batch, replace=False) for _ in range(epochs)]).astype(np.int32)
    return feat, target, mask, idx


def torch_steps(net, feat, target, mask, idx):
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    crit = nn.MSELoss()
    losses, norms = [], []
    for row in idx:
        s, t, m = (torch.from_numpy(x[row]) for x in (feat, target, mask))
        opt.zero_grad()
        loss = crit(net(s) * m, t * m)
        loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)))
        opt.step()
        losses.append(loss.item())
    return np.array(losses), norms, opt


def emu_steps(emu, blob, m, v, steps_done, feat, target, mask, idx):
    idx = np.ascontiguousarray(idx)
    loss = np.zeros(idx.shape[0], np.float32)
    grad = np.zeros(NF, np.float32)
    rc = emu.emu_sd_train(P(blob), P(m), P(v), steps_done, P(feat), P(target), P(mask), feat.shape[0],
                          idx.ctypes.data_as(ip), idx.shape[1], idx.shape[0], 5e-4, 0.9, 0.999, 1e-8, 1.0, P(loss), P(grad))
    assert rc == 0
    return loss


@pytest.mark.parametrize("batch,epochs,n_rows,scale", [(128, 6, 2000, 1.0), (32, 4, 100, 1.0), (7, 3, 7, 1.0), (1, 3, 1, 1.0),
                                                      (128, 5, 500, 40.0), (5, 4, 9, 40.0)])
def test_emulated_kernel_matches_torch(emu, batch, epochs, n_rows, scale):
    rng = np.random.default_rng(batch * 1000 + epochs)
    net = make_net(batch)
    blob, m, v = blob_of(net), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
    start = blob.copy()
    feat, target, mask, idx = make_problem(rng, n_rows, batch, epochs, scale)
    t_loss, norms, opt = torch_steps(net, feat, target, mask, idx)
    e_loss = emu_steps(emu, blob, m, v, 0, feat, target, mask, idx)
    if scale > 1:
        assert max(norms) > 1.0                      # the clip must have engaged in these cases
    np.testing.assert_allclose(e_loss, t_loss, rtol=2e-6)
    ref = blob_of(net)
    assert np.abs(ref - start).max() > 1e-4          # the parameters did move
    np.testing.assert_allclose(blob, ref, rtol=0, atol=1e-6)
    t_m = np.concatenate([opt.state[p]["exp_avg"].numpy().reshape(-1) for p in net.parameters()])
    t_v = np.concatenate([opt.state[p]["exp_avg_sq"].numpy().reshape(-1) for p in net.parameters()])
    np.testing.assert_allclose(m, t_m, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(v, t_v, rtol=1e-4, atol=1e-10)


def test_split_calls_equal_one_call(emu):
    """steps_done hands the bias correction over: 3 + 4 steps in two launches == 7 steps in one, bit for bit."""
    rng = np.random.default_rng(5)
    feat, target, mask, idx = make_problem(rng, 300, 64, 7)
    a, b = blob_of(make_net(1)), blob_of(make_net(1))
    ma, va, mb, vb = (np.zeros(NF, np.float32) for _ in range(4))
    la = emu_steps(emu, a, ma, va, 0, feat, target, mask, idx)
    lb = np.concatenate([emu_steps(emu, b, mb, vb, 0, feat, target, mask, idx[:3]),
                         emu_steps(emu, b, mb, vb, 3, feat, target, mask, idx[3:])])
    assert np.array_equal(a, b) and np.array_equal(ma, mb) and np.array_equal(va, vb) and np.array_equal(la, lb)


def test_repeatable(emu):
    """The emulation runs 512 truly asynchronous threads: a missing barrier in the kernel shows up as run-to-run noise."""
    rng = np.random.default_rng(9)
    feat, target, mask, idx = make_problem(rng, 400, 128, 3)
    outs = []
    for _ in range(4):
        blob, m, v = blob_of(make_net(2)), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
        loss = emu_steps(emu, blob, m, v, 0, feat, target, mask, idx)
        outs.append((blob, m, v, loss))
    for o in outs[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(o, outs[0]))


def test_out_of_range_rows_skip_the_step(emu):
    rng = np.random.default_rng(11)
    feat, target, mask, idx = make_problem(rng, 50, 16, 3)
    idx[1, 4] = 50                                     # one past the end, middle epoch only
    blob, m, v = blob_of(make_net(3)), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
    loss = emu_steps(emu, blob, m, v, 0, feat, target, mask, idx)
    assert np.isnan(loss[1]) and np.isfinite(loss[0]) and np.isfinite(loss[2])
    good = np.ascontiguousarray(idx[[0, 2]])
    blob2, m2, v2 = blob_of(make_net(3)), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
    loss2 = emu_steps(emu, blob2, m2, v2, 0, feat, target, mask, good)
    assert np.array_equal(blob, blob2) and np.array_equal(loss[[0, 2]], loss2)


def test_python_binding_drives_the_emulated_kernel(emu):
    """scopa_b200.sdcfr.FusedAdam / AdvantageNetwork(optimizer='fused') end to end on the CPU: the emulator exports an
    entry with ms_sdcfr_train's exact signature, bound with the argtypes of scopa_b200._lib, so argument order, the
    parameter re-homing (flatten_parameters_) and the minibatch sampler are what the GPU path uses."""
    from scopa_b200 import _lib, sdcfr
    from scopa_b200.algorithms.deep_cfr.deep_cfr import AdvantageNetwork

    entry = emu.emu_ms_sdcfr_train
    entry.argtypes, entry.restype = _lib._SIGS["ms_sdcfr_train"]
    torch.manual_seed(4)
    fused, plain = AdvantageNetwork(34, 16, device="cpu"), AdvantageNetwork(34, 16, device="cpu")
    plain.net.load_state_dict(fused.net.state_dict())
    fused._fused = sdcfr.FusedAdam(sdcfr.flatten_parameters_(fused.net), lr=5e-4, _entry=entry)
    assert fused.blob() is fused._fused.blob
    rng = np.random.default_rng(3)
    feat, target, mask, _ = make_problem(rng, 700, 1, 1)
    for adv in (fused, plain):
        adv.buffer.add_batch(*(torch.from_numpy(x) for x in (feat, target, mask)))
    rows = fused._sample_rows(128, 5)
    assert rows.shape == (5, 128) and rows.dtype == torch.int32 and int(rows.max()) < 700 and int(rows.min()) >= 0
    assert all(len(set(r.tolist())) == 128 for r in rows)                   # without replacement
    # the same minibatches through both optimisers
    fused._sample_rows = lambda batch_size, epochs: rows
    loss_fused = fused.train(batch_size=128, epochs=5)
    it = iter(rows.long())

    def same_rows(batch_size):
        r = next(it)
        return plain.buffer.feat[r], plain.buffer.target[r], plain.buffer.mask[r]

    plain.buffer.sample = same_rows
    loss_plain = plain.train(batch_size=128, epochs=5)
    assert abs(loss_fused - loss_plain) < 2e-6 * abs(loss_plain)
    for (n1, p1), (n2, p2) in zip(fused.net.named_parameters(), plain.net.named_parameters()):
        assert n1 == n2 and torch.allclose(p1, p2, rtol=0, atol=1e-6), n1
    # the module's parameters ARE the blob the kernel updated (views), and snapshots copy them
    assert np.allclose(sdcfr.flatten_net(fused.net).numpy(), fused._fused.blob.numpy(), rtol=0, atol=0)
    assert fused._fused.steps_done == 5
    # small buffer: the reference's fallback batch size min(len, 32)
    small = AdvantageNetwork(34, 16, device="cpu")
    small._fused = sdcfr.FusedAdam(sdcfr.flatten_parameters_(small.net), _entry=entry)
    small.buffer.add_batch(*(torch.from_numpy(x[:20]) for x in (feat, target, mask)))
    assert np.isfinite(small.train(batch_size=128, epochs=2)) and small._fused.steps_done == 2
    assert AdvantageNetwork(34, 16, device="cpu").train() == 0.0             # empty buffer


def test_average_policy_kernels_match_the_reference_loop(emu):
    """sd_avgpol_kernel + sd_avgpol_reduce_kernel (emulated) through StrategyBuffer against the reference's loop of
    batch-1 forwards (deep_cfr.py:136-160, here: torch on the CPU): single decisions, a batch that is not a multiple of
    the 64-row chunk, the max_size eviction and the cache invalidation when the lists change."""
    from scopa_b200 import _lib
    from scopa_b200.algorithms.deep_cfr.deep_cfr import HIDDEN, StrategyBuffer
    from scopa_b200.algorithms.deep_cfr.nets import FlexibleNet

    entry = emu.emu_ms_sdcfr_average_policy
    entry.argtypes, entry.restype = _lib._SIGS["ms_sdcfr_average_policy"]
    torch.manual_seed(8)
    rng = np.random.default_rng(8)

    def snapshot():
        net = FlexibleNet(mode="mlp", input_shape=(34,), output_dim=16, mlp_hidden=HIDDEN, mlp_act="relu", mlp_norm="none")
        for l in net.modules():
            if isinstance(l, nn.Linear):
                nn.init.xavier_uniform_(l.weight)
                nn.init.normal_(l.bias, 0.0, 0.3)
        return net

    fast, ref = StrategyBuffer(max_size=4), StrategyBuffer(max_size=4)
    fast._entry = entry
    feat, _, mask, _ = make_problem(rng, 130, 1, 1)
    assert np.array_equal(fast.get_average_policy(feat[0], mask[0]), mask[0] / mask[0].sum())     # empty buffer: uniform
    for it in range(1, 7):                       # six snapshots into a buffer of four: the two oldest are evicted
        net = snapshot()
        fast.add_strategy(net, it)
        ref.add_strategy(net, it)
        for r in (0, 5):
            a, b = fast.get_average_policy(feat[r], mask[r]), ref.get_average_policy(feat[r], mask[r])
            assert a.dtype == np.float32 and a.shape == (16,)
            np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-6)   # pos / z amplifies the sgemm-order rounding of the advantages
            assert np.all(a[mask[r] == 0] == 0)
    assert fast.weights == [4, 5, 6, 7] and len(fast._blob_of) == 4
    got = fast.average_policy_batch(torch.from_numpy(feat), torch.from_numpy(mask)).numpy()
    want = np.stack([ref.get_average_policy(feat[r], mask[r]) for r in range(130)])
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
    assert float(got.sum(1).max()) < 1 + 1e-5


def test_net_snapshot_behaves_like_the_stored_module(emu):
    """NetSnapshot (the fused configuration's stored strategy net: one blob copy) against a FlexibleNet copy made the
    reference's way (construct + load_state_dict, deep_cfr.py:460-472)."""
    from scopa_b200 import _lib, sdcfr
    from scopa_b200.algorithms.deep_cfr.deep_cfr import HIDDEN, NetSnapshot, StrategyBuffer
    from scopa_b200.algorithms.deep_cfr.nets import FlexibleNet

    torch.manual_seed(12)
    live = FlexibleNet(mode="mlp", input_shape=(34,), output_dim=16, mlp_hidden=HIDDEN, mlp_act="relu", mlp_norm="none")
    blob = sdcfr.flatten_parameters_(live)
    snap = NetSnapshot(blob)
    copy_ = FlexibleNet(mode="mlp", input_shape=(34,), output_dim=16, mlp_hidden=HIDDEN, mlp_act="relu", mlp_norm="none")
    copy_.load_state_dict(live.state_dict())
    x = torch.rand(9, 34)
    with torch.no_grad():
        want = copy_(x)
        assert torch.equal(snap(x), want)
        blob += 1.0                                         # the live net trains on: the snapshot must not follow
        assert torch.equal(snap(x), want) and not torch.equal(live(x), want)
        assert snap._module is None                         # nothing above needed the nn.Module
        # module-shaped access materialises a FlexibleNet over the same blob
        assert [tuple(p.shape) for p in snap.parameters()] == [tuple(p.shape) for p in copy_.parameters()]
        assert set(snap.state_dict()) == set(copy_.state_dict())
        assert all(torch.equal(snap.state_dict()[k], copy_.state_dict()[k]) for k in copy_.state_dict())
        assert torch.equal(snap.backbone[0].fc.weight, copy_.backbone[0].fc.weight) and torch.equal(snap.module()(x), want)
        assert snap.backbone[0].fc.weight.data_ptr() == snap._blob.data_ptr()
    with pytest.raises(AttributeError):
        snap.no_such_attribute
    # inside a StrategyBuffer: same average policy as module snapshots, through the emulated kernels
    entry = emu.emu_ms_sdcfr_average_policy
    entry.argtypes, entry.restype = _lib._SIGS["ms_sdcfr_average_policy"]
    a, b = StrategyBuffer(), StrategyBuffer()
    a._entry = b._entry = entry
    rng = np.random.default_rng(1)
    feat, _, mask, _ = make_problem(rng, 3, 1, 1)
    for it in (1, 2, 3):
        net = FlexibleNet(mode="mlp", input_shape=(34,), output_dim=16, mlp_hidden=HIDDEN, mlp_act="relu", mlp_norm="none")
        a.add_strategy(net, it)
        b.add_strategy(NetSnapshot(sdcfr.flatten_net(net)), it)
    for r in range(3):
        assert np.array_equal(a.get_average_policy(feat[r], mask[r]), b.get_average_policy(feat[r], mask[r]))


# ---------------------------------------------------------------------------------------------------------------------
# sd_train_cluster_kernel (csrc/ms_sd_train_cluster.cuh): 8 CTAs emulated concurrently (8 x 256 host threads), cluster
# barrier = a barrier over all of them, distributed shared memory = the other blocks' buffers.

def cluster_steps(emu, blob, m, v, steps_done, feat, target, mask, idx):
    from scopa_b200 import _lib
    f = emu.emu_ms_sdcfr_train_cluster
    f.argtypes, f.restype = _lib._SIGS["ms_sdcfr_train_cluster"]
    idx = np.ascontiguousarray(idx)
    loss, ws = np.zeros(idx.shape[0], np.float32), np.zeros(NF, np.float32)
    rc = f(blob.ctypes.data, m.ctypes.data, v.ctypes.data, steps_done, feat.ctypes.data, target.ctypes.data, mask.ctypes.data,
           feat.shape[0], idx.ctypes.data, idx.shape[1], idx.shape[0], 5e-4, 0.9, 0.999, 1e-8, 1.0, loss.ctypes.data,
           ws.ctypes.data, ws.nbytes, None)
    assert rc == 0
    return loss


@pytest.mark.parametrize("batch,epochs,n_rows,scale", [(128, 4, 1000, 1.0), (32, 3, 100, 1.0), (17, 3, 40, 40.0), (1, 2, 1, 1.0),
                                                      (100, 3, 300, 40.0)])
def test_emulated_cluster_kernel_matches_torch(emu, batch, epochs, n_rows, scale):
    """Rows are split 16 per CTA (batch 17: one CTA with 16 rows, one with 1, six with none; batch 100: the 7th CTA has 4)."""
    rng = np.random.default_rng(batch * 1000 + epochs)
    net = make_net(batch)
    blob, m, v = blob_of(net), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
    feat, target, mask, idx = make_problem(rng, n_rows, batch, epochs, scale)
    t_loss, norms, opt = torch_steps(net, feat, target, mask, idx)
    c_loss = cluster_steps(emu, blob, m, v, 0, feat, target, mask, idx)
    if scale > 1:
        assert max(norms) > 1.0
    np.testing.assert_allclose(c_loss, t_loss, rtol=2e-6)
    np.testing.assert_allclose(blob, blob_of(net), rtol=0, atol=1e-6)
    t_m = np.concatenate([opt.state[p]["exp_avg"].numpy().reshape(-1) for p in net.parameters()])
    t_v = np.concatenate([opt.state[p]["exp_avg_sq"].numpy().reshape(-1) for p in net.parameters()])
    np.testing.assert_allclose(m, t_m, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(v, t_v, rtol=1e-4, atol=1e-10)


def test_cluster_kernel_repeatable_split_and_bad_rows(emu):
    rng = np.random.default_rng(21)
    feat, target, mask, idx = make_problem(rng, 300, 50, 5)
    runs = []
    for _ in range(3):                                   # 2048 asynchronous host threads: a missing barrier shows as noise
        blob, m, v = blob_of(make_net(6)), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
        loss = cluster_steps(emu, blob, m, v, 0, feat, target, mask, idx)
        runs.append((blob, m, v, loss))
    for r in runs[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(r, runs[0]))
    # 2 + 3 steps in two launches == 5 in one (moments and bias correction are handed over through global memory)
    blob, m, v = blob_of(make_net(6)), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
    l2 = np.concatenate([cluster_steps(emu, blob, m, v, 0, feat, target, mask, idx[:2]),
                         cluster_steps(emu, blob, m, v, 2, feat, target, mask, idx[2:])])
    assert all(np.array_equal(x, y) for x, y in zip((blob, m, v, l2), runs[0]))
    # the one-CTA kernel sums the gradient in another order: same result to fp32 rounding
    blob1, m1, v1 = blob_of(make_net(6)), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
    l1 = emu_steps(emu, blob1, m1, v1, 0, feat, target, mask, idx)
    np.testing.assert_allclose(l1, runs[0][3], rtol=2e-6)
    np.testing.assert_allclose(blob1, runs[0][0], rtol=0, atol=1e-6)
    # a bad row in ANY CTA's share skips the step in all of them
    bad = idx.copy()
    bad[1, 40] = 300                                     # row 40 belongs to CTA 2
    blob, m, v = blob_of(make_net(6)), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
    lb = cluster_steps(emu, blob, m, v, 0, feat, target, mask, bad)
    assert np.isnan(lb[1]) and np.isfinite(lb[[0, 2, 3, 4]]).all()
    blob_g, m_g, v_g = blob_of(make_net(6)), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
    lg = cluster_steps(emu, blob_g, m_g, v_g, 0, feat, target, mask, np.ascontiguousarray(idx[[0, 2, 3, 4]]))
    assert np.array_equal(blob, blob_g) and np.array_equal(lb[[0, 2, 3, 4]], lg)


# ---------------------------------------------------------------------------------------------------------------------
# sd_sample_rows_kernel (csrc/ms_sd_sample.cuh): the minibatch rows of all epochs in one launch.

def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al. 2011), plain Python integers."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = (p1 >> 32) ^ c1 ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c3 ^ k1, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def sample_rows_restated(batch, epochs, n_rows, seed, first_epoch):
    """The kernel's procedure as its header states it, sequentially."""
    out = np.zeros((epochs, batch), np.int32)
    key = (seed & 0xFFFFFFFF, seed >> 32)
    for ep in range(epochs):
        E = first_epoch + ep
        attempt, row, dirty = [0] * batch, [0] * batch, [True] * batch
        while any(dirty):
            for m in range(batch):
                if dirty[m]:
                    x = philox4x32_10((E & 0xFFFFFFFF, E >> 32, m | (attempt[m] << 8), 0x52544453), key)[0]
                    row[m] = (x * n_rows) >> 32
                    attempt[m] += 1
            dirty = [any(row[j] == row[m] for j in range(m)) for m in range(batch)]
        out[ep] = row
    return out


def emu_sample(emu, batch, epochs, n_rows, seed, first_epoch=0):
    from scopa_b200 import _lib
    f = emu.emu_ms_sdcfr_sample_rows
    f.argtypes, f.restype = _lib._SIGS["ms_sdcfr_sample_rows"]
    idx = np.full((epochs, batch), -7, np.int32)
    assert f(idx.ctypes.data, batch, epochs, n_rows, seed, first_epoch, None) == 0
    return idx


def test_philox_known_answer():
    # Random123's published vectors for philox4x32-10
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


@pytest.mark.parametrize("batch,epochs,n_rows,seed,first", [(128, 5, 100000, 1234, 0), (128, 3, 130, 7, 40), (32, 4, 41, 99, 3),
                                                           (32, 3, 32, 5, 0), (1, 3, 1, 1, 0), (7, 2, 9, 2 ** 40 + 3, 2 ** 33)])
def test_emulated_sampler_equals_the_restatement(emu, batch, epochs, n_rows, seed, first):
    got = emu_sample(emu, batch, epochs, n_rows, seed, first)
    assert np.array_equal(got, sample_rows_restated(batch, epochs, n_rows, seed, first))
    assert got.min() >= 0 and got.max() < n_rows
    assert all(len(set(r.tolist())) == batch for r in got)          # random.sample: distinct rows


def test_sampler_stream_properties(emu):
    # consecutive calls continue the stream: epochs [0, 5) == [0, 2) + [2, 5)
    a = emu_sample(emu, 64, 5, 5000, 11, 0)
    assert np.array_equal(a, np.concatenate([emu_sample(emu, 64, 2, 5000, 11, 0), emu_sample(emu, 64, 3, 5000, 11, 2)]))
    assert not np.array_equal(a, emu_sample(emu, 64, 5, 5000, 12, 0))                   # another seed, another stream
    # every row equally likely: 400 epochs x 32 of 64 rows -> each row is drawn about 200 times (sd = 10)
    counts = np.bincount(emu_sample(emu, 32, 400, 64, 3, 0).reshape(-1), minlength=64)
    assert counts.sum() == 400 * 32 and counts.min() > 150 and counts.max() < 250


def test_python_binding_of_the_sampler(emu):
    """AdvantageNetwork(sampler_seed=...) -> FusedAdam.sample_rows -> (emulated) ms_sdcfr_sample_rows: the rows follow the
    optimiser's step count, feed the (emulated) optimiser kernel, and a run is repeatable from its seed."""
    from scopa_b200 import _lib, sdcfr
    from scopa_b200.algorithms.deep_cfr.deep_cfr import AdvantageNetwork

    train, sample = emu.emu_ms_sdcfr_train, emu.emu_ms_sdcfr_sample_rows
    train.argtypes, train.restype = _lib._SIGS["ms_sdcfr_train"]
    sample.argtypes, sample.restype = _lib._SIGS["ms_sdcfr_sample_rows"]
    rng = np.random.default_rng(13)
    feat, target, mask, _ = make_problem(rng, 500, 1, 1)

    def run():
        torch.manual_seed(21)
        adv = AdvantageNetwork(34, 16, device="cpu", sampler_seed=77)
        adv._fused = sdcfr.FusedAdam(sdcfr.flatten_parameters_(adv.net), _entry=train)
        adv._sample_entry = sample
        adv.buffer.add_batch(*(torch.from_numpy(x) for x in (feat, target, mask)))
        rows0 = adv._sample_rows(128, 3)
        losses = [adv.train(batch_size=128, epochs=3), adv.train(batch_size=128, epochs=2)]
        return rows0.numpy(), losses, adv._fused.blob.clone(), adv._fused.steps_done

    r1, l1, b1, n1 = run()
    r2, l2, b2, n2 = run()
    assert np.array_equal(r1, sample_rows_restated(128, 3, 500, 77, 0)) and n1 == 5
    assert np.array_equal(r1, r2) and l1 == l2 and torch.equal(b1, b2)           # repeatable from the seed alone
    assert all(np.isfinite(l1))


def test_device_replay_buffer_is_a_deque_of_maxlen():
    """DeviceReplayBuffer (ring of three tensors, slice copies) against collections.deque(maxlen=...) -- the reference's
    buffer (deep_cfr.py:36) -- through wrap-arounds, oversize batches and single appends."""
    from collections import deque
    from scopa_b200.algorithms.deep_cfr.deep_cfr import DeviceReplayBuffer
    rng = np.random.default_rng(4)
    buf, ref = DeviceReplayBuffer(50, "cpu"), deque(maxlen=50)
    tag = 0
    for n in (7, 30, 20, 50, 3, 120, 1, 49, 2):
        f = np.zeros((n, 34), np.float32)
        f[:, 0] = np.arange(tag, tag + n)
        t, m = rng.random((n, 16)).astype(np.float32), rng.random((n, 16)).astype(np.float32)
        tag += n
        if n == 1:
            buf.append((f[0], t[0], m[0]))
        else:
            buf.add_batch(torch.from_numpy(f), torch.from_numpy(t), torch.from_numpy(m))
        for i in range(n):
            ref.append((f[i], t[i], m[i]))
        assert len(buf) == len(ref)
        have = sorted(((r[0][0], r[1].tobytes(), r[2].tobytes()) for r in buf))
        want = sorted(((r[0][0], r[1].tobytes(), r[2].tobytes()) for r in ref))
        assert have == want                                  # same multiset of rows (a minibatch is a random subset)


def test_numpy_oracle_pinned_to_torch_and_kernels_to_the_oracle(emu):
    """oracle/sd_train.py (explicit matrix algebra, no autograd) == torch (the reference's arithmetic) == emulated kernels."""
    from oracle import sd_train as ora
    rng = np.random.default_rng(31)
    for batch, epochs, n_rows, scale in ((128, 5, 800, 1.0), (32, 3, 60, 40.0)):
        net = make_net(batch + 1)
        start = blob_of(net)
        feat, target, mask, idx = make_problem(rng, n_rows, batch, epochs, scale)
        t_loss, _, opt = torch_steps(net, feat, target, mask, idx)
        o_blob, o_m, o_v = start.copy(), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
        o_loss = ora.train_steps(o_blob, o_m, o_v, 0, feat, target, mask, idx)
        np.testing.assert_allclose(o_loss, t_loss, rtol=2e-6)
        np.testing.assert_allclose(o_blob, blob_of(net), rtol=0, atol=1e-6)
        e_blob, e_m, e_v = start.copy(), np.zeros(NF, np.float32), np.zeros(NF, np.float32)
        e_loss = emu_steps(emu, e_blob, e_m, e_v, 0, feat, target, mask, idx)
        np.testing.assert_allclose(e_loss, o_loss, rtol=2e-6)
        np.testing.assert_allclose(e_blob, o_blob, rtol=0, atol=1e-6)
        np.testing.assert_allclose(e_m, o_m, rtol=1e-4, atol=1e-8)
    # average policy: oracle vs the emulated two-kernel path
    from scopa_b200 import _lib
    f = emu.emu_ms_sdcfr_average_policy
    f.argtypes, f.restype = _lib._SIGS["ms_sdcfr_average_policy"]
    blobs = np.stack([blob_of(make_net(s)) for s in (3, 4, 5)])
    weights = [2, 3, 4]
    w32 = np.array([np.float32(w / sum(weights)) for w in weights], np.float32)
    feat, _, mask, _ = make_problem(rng, 90, 1, 1)
    pol, ws = np.zeros((90, 16), np.float32), np.zeros(3 * 90 * 16, np.float32)
    assert f(blobs.ctypes.data, w32.ctypes.data, 3, feat.ctypes.data, mask.ctypes.data, 90, pol.ctypes.data, ws.ctypes.data,
             ws.nbytes, None) == 0
    np.testing.assert_allclose(pol, ora.average_policy(blobs, weights, feat, mask), rtol=2e-5, atol=2e-6)
