"""Low-level SDCFR entry points over the C ABI: batched advantage-net inference and the level-batched
external-sampling traversal (csrc/ms_sdcfr.cu)."""
import ctypes as C

import numpy as np
import torch

from . import _lib

NET_FLOATS = 13776
FP32, TENSOR_CORE = 0, 1


def flatten_net(net):
    """FlexibleNet (mlp 34->128->64->16) -> one fp32 CUDA blob in nn.Linear order (w1 b1 w2 b2 w3 b3)."""
    layers = [net.backbone[0].fc, net.backbone[1].fc, net.head]
    parts = []
    for l in layers:
        parts += [l.weight.detach().reshape(-1), l.bias.detach().reshape(-1)]
    blob = torch.cat(parts).to(dtype=torch.float32).contiguous()
    assert blob.numel() == NET_FLOATS, blob.numel()
    return blob


def blob_from_arrays(w1, b1, w2, b2, w3, b3, device="cuda"):
    parts = [np.asarray(a, dtype=np.float32).reshape(-1) for a in (w1, b1, w2, b2, w3, b3)]
    blob = torch.from_numpy(np.concatenate(parts)).to(device)
    assert blob.numel() == NET_FLOATS
    return blob


def mlp_forward(blob, feat, mask, precision=FP32):
    """-> (advantages [n,16] masked like get_advantages, policy [n,16] = positive_regret_policy)."""
    lib = _lib.load()
    feat = feat.to(dtype=torch.float32).contiguous()
    mask = mask.to(dtype=torch.float32).contiguous()
    n = feat.shape[0]
    adv = torch.empty((n, 16), dtype=torch.float32, device=feat.device)
    pol = torch.empty((n, 16), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.check(lib.ms_mlp_forward(blob.data_ptr(), int(precision), feat.data_ptr(), mask.data_ptr(), adv.data_ptr(),
                                      pol.data_ptr(), n, _lib.stream_ptr()))
    return adv, pol


def infer_states(blob, states, player_to_move, precision=FP32):
    """The inference half of one traversal level on its own (sd_level_mlp_kernel): packed states [n,4] (the uint32 words
    as an int32 torch tensor on the device, all with `player_to_move` to move) -> the advantage net's raw outputs [n,16]
    (before masking), features of the mover's view."""
    lib = _lib.load()
    states = states.contiguous()
    n = states.shape[0]
    raw = torch.empty((n, 16), dtype=torch.float32, device=states.device)
    with torch.cuda.device(states.device):
        _lib.check(lib.ms_sdcfr_infer_states(states.data_ptr(), n, int(player_to_move), blob.data_ptr(), int(precision),
                                             raw.data_ptr(), _lib.stream_ptr()))
    return raw


class Traverser:
    """Reusable workspace for batches of external-sampling traversals from one root."""

    def __init__(self, root_words, hand_order, device="cuda"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.root = (C.c_uint32 * 4)(*[int(w) & 0xFFFFFFFF for w in root_words])
        self.hand_order = int(hand_order) & 0xFFFFFFFF
        self._ws = None
        self.samples_per_traversal = [self.lib.ms_sdcfr_samples_per_traversal(p) for p in (0, 1)]

    def run(self, player, blobs, n_trav, philox_seed=0, first_trav=0, precision=FP32):
        """-> feat [n*41, 34], target [n*41, 16], mask [n*41, 16], root values [n] (CUDA float32 tensors)."""
        need = self.lib.ms_sdcfr_workspace_bytes(int(n_trav))
        with torch.cuda.device(self.device):
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            m = n_trav * self.samples_per_traversal[player]
            feat = torch.empty((m, 34), dtype=torch.float32, device=self.device)
            target = torch.empty((m, 16), dtype=torch.float32, device=self.device)
            mask = torch.empty((m, 16), dtype=torch.float32, device=self.device)
            value = torch.empty((n_trav,), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.ms_sdcfr_traverse(self.root, self.hand_order, int(player), blobs[0].data_ptr(),
                                                  blobs[1].data_ptr(), int(precision), int(n_trav), int(philox_seed),
                                                  int(first_trav), self._ws.data_ptr(), self._ws.numel(),
                                                  feat.data_ptr(), target.data_ptr(), mask.data_ptr(), value.data_ptr(),
                                                  _lib.stream_ptr()))
        return feat, target, mask, value


class FusedAdam:
    """State of the fused optimiser (`ms_sdcfr_train`): Adam moments and step count of ONE advantage net whose
    parameters live in a flat fp32 blob.  `step()` runs `epochs` optimiser steps of AdvantageNetwork.train
    (/root/reference/src/algorithms/deep_cfr/deep_cfr.py:77-110) in one kernel launch."""

    def __init__(self, blob, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, kernel="cta", _entry=None):
        """`kernel`: "cta" = sd_train_kernel (one CTA), "cluster" = sd_train_cluster_kernel (8 CTAs exchanging gradients
        and weights through distributed shared memory; `ms_sdcfr_train_cluster`).  `_entry` (tests only): a host function with ms_sdcfr_train's signature, e.g. the emulated kernel of
        tests/emu -- the tensors then live on the CPU.  The product path always launches the CUDA kernel."""
        assert blob.dtype == torch.float32 and blob.numel() == NET_FLOATS and blob.is_contiguous()
        self._emulated = _entry is not None
        if self._emulated:
            assert not blob.is_cuda
            self._entry, ws_bytes = _entry, NET_FLOATS * 4
        else:
            if not blob.is_cuda:
                raise _lib.MsError("FusedAdam needs a CUDA blob: scopa_b200 has no CPU path")
            if kernel not in ("cta", "cluster"):
                raise ValueError(f"kernel must be 'cta' or 'cluster', not {kernel!r}")
            lib = _lib.load()
            self._entry = lib.ms_sdcfr_train if kernel == "cta" else lib.ms_sdcfr_train_cluster
            ws_bytes = lib.ms_sdcfr_train_workspace_bytes()
        self.blob = blob
        self.m = torch.zeros_like(blob)
        self.v = torch.zeros_like(blob)
        self.steps_done = 0
        self.lr, self.betas, self.eps, self.max_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(max_norm)
        self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=blob.device)

    def sample_rows(self, batch, epochs, n_rows, seed, _entry=None):
        """[epochs, batch] int32 minibatch rows for the next `epochs` optimiser steps (`ms_sdcfr_sample_rows`: distinct
        rows per epoch from the Philox stream (seed, global step); one launch).  `_entry`: emulated entry point (tests)."""
        idx = torch.empty((epochs, batch), dtype=torch.int32, device=self.blob.device)
        args = (idx.data_ptr(), int(batch), int(epochs), int(n_rows), int(seed) & (2 ** 64 - 1), int(self.steps_done))
        if self._emulated:
            rc = _entry(*args, None)
            if rc != 0:
                raise _lib.MsError(f"emulated ms_sdcfr_sample_rows returned {rc}")
        else:
            with torch.cuda.device(self.blob.device):
                _lib.check(_lib.load().ms_sdcfr_sample_rows(*args, _lib.stream_ptr()))
        return idx

    def step(self, feat, target, mask, idx):
        """idx [epochs, batch] int32 rows of (feat [n,34], target [n,16], mask [n,16]) -> losses [epochs]; all tensors on
        the blob's device."""
        assert idx.dtype == torch.int32 and idx.dim() == 2 and idx.is_contiguous() and idx.device == self.blob.device
        for t, w in ((feat, 34), (target, 16), (mask, 16)):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.shape[1] == w and t.device == self.blob.device
        n_rows = min(feat.shape[0], target.shape[0], mask.shape[0])
        epochs, batch = idx.shape
        loss = torch.empty((epochs,), dtype=torch.float32, device=self.blob.device)
        args = (self.blob.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.steps_done, feat.data_ptr(),
                target.data_ptr(), mask.data_ptr(), n_rows, idx.data_ptr(), batch, epochs, self.lr, self.betas[0],
                self.betas[1], self.eps, self.max_norm, loss.data_ptr(), self._ws.data_ptr(), self._ws.numel())
        if self._emulated:
            rc = self._entry(*args, None)
            if rc != 0:
                raise _lib.MsError(f"emulated ms_sdcfr_train returned {rc}")
        else:
            with torch.cuda.device(self.blob.device):
                _lib.check(self._entry(*args, _lib.stream_ptr()))
        # (an epoch with out-of-range rows is skipped by the kernel and reported as NaN; it still counts here -- such rows
        # are a caller bug, not a state to resume from)
        self.steps_done += epochs
        return loss


def flatten_parameters_(net):
    """Re-homes the parameters of a FlexibleNet (mlp 34->128->64->16) as views into one flat fp32 blob (nn.Linear
    order) and returns the blob: the fused optimiser then updates the module in place and `flatten_net` costs nothing."""
    blob = flatten_net(net).clone()
    layers = [net.backbone[0].fc, net.backbone[1].fc, net.head]
    off = 0
    for l in layers:
        for p in (l.weight, l.bias):
            n = p.numel()
            p.data = blob[off:off + n].view(p.shape)
            off += n
    assert off == NET_FLOATS
    return blob


def average_policy(nets, weights, feat, mask, _entry=None):
    """StrategyBuffer.get_average_policy for a batch (`ms_sdcfr_average_policy`): nets [K, 13776] fp32 blobs, weights
    [K] fp32 (= weight_k / total_weight), feat [n, 34], mask [n, 16] -> policy [n, 16]; every tensor on the same CUDA
    device.  `_entry` (tests only): a host function with the entry point's signature (the emulated kernels of
    tests/emu), tensors on the CPU."""
    dev = nets.device
    assert nets.dim() == 2 and nets.shape[1] == NET_FLOATS and weights.shape == (nets.shape[0],)
    assert feat.dim() == 2 and feat.shape[1] == 34 and mask.shape == (feat.shape[0], 16)
    for t in (nets, weights, feat, mask):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.device == dev
    k, n = nets.shape[0], feat.shape[0]
    policy = torch.empty((n, 16), dtype=torch.float32, device=dev)
    if n == 0:
        return policy
    ws = torch.empty(max(1, k * n * 16), dtype=torch.float32, device=dev)
    args = (nets.data_ptr(), weights.data_ptr(), k, feat.data_ptr(), mask.data_ptr(), n, policy.data_ptr(), ws.data_ptr(),
            ws.numel() * 4)
    if _entry is not None:
        assert not nets.is_cuda
        rc = _entry(*args, None)
        if rc != 0:
            raise _lib.MsError(f"emulated ms_sdcfr_average_policy returned {rc}")
    else:
        if not nets.is_cuda:
            raise _lib.MsError("average_policy needs CUDA tensors: scopa_b200 has no CPU path")
        with torch.cuda.device(dev):
            _lib.check(_lib.load().ms_sdcfr_average_policy(*args, _lib.stream_ptr()))
    return policy
