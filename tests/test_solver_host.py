"""The PRODUCT's tree enumeration and vanilla-CFR kernels on the CPU: tree_expand_kernel and cfr_kernel of
scopa_b200/csrc/ms_solver.cu run by the CTA emulator (tests/emu/ms_solver_host.cpp: 256 / 512 pthreads, barriers for
__syncthreads, IEEE double without contraction), against

  * the tables recorded from the UNMODIFIED reference (tests/golden/cfr_seed42.npz: CFRTrainer.train on the seed-42
    deal after 1 / 2 / 5 / 20 iterations) -- the same float64 bits, as in tests/test_gpu_solver.py on the device;
  * the oracle on other deals, and CFRTrainer._cfr_recursive called directly (one traversal, explicit reaches).
"""
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


class HostSolver:
    def __init__(self, lib, seed):
        cards = ora.deck(seed)
        words = codec.pack_state([codec.mask_of(cards[:4]), codec.mask_of(cards[4:8])], [], [0, 0], [0, 0], 0, 0, False, 8)
        root = np.array(words, dtype=np.uint32)
        self.hand_order = codec.pack_nibbles(cards[:8])
        n, s, l = C.c_int(), C.c_int(), C.c_int()
        assert lib.host_solver_build(root.ctypes.data, self.hand_order, C.byref(n), C.byref(s), C.byref(l)) == 0
        self.lib, self.n_nodes, self.n_slots, self.n_levels = lib, n.value, s.value, l.value

    def table(self):
        S = self.n_slots
        keys, nl, legal = np.zeros(S, np.uint64), np.zeros(S, np.uint8), np.zeros((S, 4), np.uint8)
        reg, strat = np.zeros((S, 4)), np.zeros((S, 4))
        self.lib.host_solver_export(keys.ctypes.data, nl.ctypes.data, legal.ctypes.data, reg.ctypes.data, strat.ctypes.data)
        return {"strings": [codec.key_to_string(k, self.hand_order) for k in keys], "nlegal": nl, "legal": legal,
                "regret": reg, "strategy": strat}

    def tree(self):
        N = self.n_nodes
        st, par, cb = np.zeros((N, 4), np.uint32), np.zeros(N, np.int32), np.zeros(N, np.int32)
        nc, slot, lb = np.zeros(N, np.uint8), np.zeros(N, np.int32), np.zeros(self.n_levels + 1, np.int32)
        self.lib.host_solver_tree(st.ctypes.data, par.ctypes.data, cb.ctypes.data, nc.ctypes.data, slot.ctypes.data, lb.ctypes.data)
        return {"state": st, "parent": par, "child_begin": cb, "nchild": nc, "slot": slot, "level_begin": lb}

    def cfr(self, iters, only_player=-1, r0=1.0, r1=1.0):
        out = C.c_double(0.0)
        assert self.lib.host_cfr(iters, only_player, r0, r1, C.byref(out)) == 0
        return out.value


@pytest.fixture(scope="module")
def lib():
    L = C.CDLL(emu_build.build_solver_host())
    L.host_solver_build.argtypes = [vp, C.c_uint32, vp, vp, vp]
    L.host_solver_tree.argtypes = [vp] * 6
    L.host_solver_export.argtypes = [vp] * 5
    L.host_cfr.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, vp]
    return L


def _perm(strings, keys):
    idx = {k: i for i, k in enumerate(strings)}
    return np.array([idx[k] for k in keys])


def test_tree_enumeration_seed42(lib):
    sv = HostSolver(lib, 42)
    assert (sv.n_nodes, sv.n_slots, sv.n_levels) == (2229, 738, 9)
    g = np.load(os.path.join(GOLDEN, "cfr_seed42.npz"))
    tab, t = sv.table(), sv.tree()
    # depth-first first-visit order of the slots = the reference's dict insertion order
    order, seen, stack = [], set(), [0]
    while stack:
        v = stack.pop()
        s = int(t["slot"][v])
        if s >= 0 and s not in seen:
            seen.add(s)
            order.append(s)
        c0, n = int(t["child_begin"][v]), int(t["nchild"][v])
        stack.extend(range(c0 + n - 1, c0 - 1, -1))
    assert [tab["strings"][i] for i in order] == list(g["keys"])
    perm = _perm(tab["strings"], list(g["keys"]))
    assert np.array_equal(tab["nlegal"][perm], g["nlegal"].astype(np.uint8))
    legal = tab["legal"][perm].astype(np.int16)
    legal[legal == 255] = -1
    assert np.array_equal(legal, g["legal"].astype(np.int16))
    assert np.diff(t["level_begin"]).tolist() == [1, 4, 16, 48, 144, 288, 576, 576, 576]
    assert int((t["nchild"] == 0).sum()) == 576


def test_vanilla_cfr_bit_identical_to_reference_tables(lib):
    g = np.load(os.path.join(GOLDEN, "cfr_seed42.npz"))
    sv = HostSolver(lib, 42)
    perm = _perm(sv.table()["strings"], list(g["keys"]))
    done = 0
    for it in (1, 2, 5, 20):
        sv.cfr(it - done)
        done = it
        tab = sv.table()
        assert np.array_equal(tab["regret"][perm], g[f"reg_{it}"]), it          # the same float64 bits as numpy
        assert np.array_equal(tab["strategy"][perm], g[f"strat_{it}"]), it
    root = tab["strings"].index("P0:H[9f-6p-5f-7f]_T[]")
    np.testing.assert_allclose(tab["regret"][root], [-15.55635194, 1.44809481, -19.8187442, -15.12762408], rtol=1e-8)


@pytest.mark.parametrize("seed", [1, 43, 2**33 + 7])
def test_vanilla_cfr_other_deals_vs_oracle(lib, seed):
    sv = HostSolver(lib, seed)
    sv.cfr(5)
    tab = sv.table()
    t = ora.Table()
    t.cfr_train(5, seed=seed)
    keys, oreg, ostrat, _, _ = t.arrays()
    assert len(keys) == sv.n_slots
    perm = _perm(tab["strings"], [k.split("|", 1)[1] if "|" in k else k for k in keys])
    assert np.array_equal(tab["regret"][perm], oreg) and np.array_equal(tab["strategy"][perm], ostrat)


def test_cfr_traverse_is_one_recursive_call(lib):
    """CFRTrainer._cfr_recursive(state, player, 1.0, 1.0), called directly by the reference's
    run_vanilla_cfr_experiment.py:91: traverser 0 then traverser 1 == one train() step."""
    sv = HostSolver(lib, 42)
    sv.cfr(2)
    a = sv.table()
    sv = HostSolver(lib, 42)
    vals = []
    for _ in range(2):
        vals.append((sv.cfr(1, only_player=0), sv.cfr(1, only_player=1)))
    b = sv.table()
    assert np.array_equal(a["regret"], b["regret"]) and np.array_equal(a["strategy"], b["strategy"])
    assert all(np.isfinite(v) and abs(v) <= 4.5 for pair in vals for v in pair)


def test_mccfr_inplace_matches_oracle_stream(lib):
    """mccfr_inplace_kernel (the reference's sampled estimator with reference semantics: every update visible to the next
    node visit; mc_cfr.py:37-86) on the oracle's Philox stream: the same float64 bits, the same first-touch set as the
    reference's dict, and SURVEY 3.2's per-iteration counts (703 calls, 172 updates)."""
    lib.host_mccfr_inplace.argtypes = [C.c_longlong, C.c_ulonglong, C.c_ulonglong]
    lib.host_solver_counters.argtypes = [vp, vp]
    sv = HostSolver(lib, 42)
    strings = sv.table()["strings"]
    t = ora.Table()
    rng = ora.Rng(1, 777)
    done = 0
    for it in (1, 3, 25):
        assert lib.host_mccfr_inplace(it - done, 777, done) == 0
        t.mccfr_iterate(it - done, rng, first_iter=done)
        done = it
        tab = sv.table()
        cnt, touched = np.zeros(3, np.uint64), np.zeros(sv.n_slots, np.uint8)
        lib.host_solver_counters(cnt.ctypes.data, touched.ctypes.data)
        keys, oreg, ostrat, _, _ = t.arrays()
        perm = _perm(strings, [k.split("|", 1)[1] for k in keys])
        assert int(touched.sum()) == len(keys) and touched[perm].all()
        assert np.array_equal(tab["regret"][perm], oreg) and np.array_equal(tab["strategy"][perm], ostrat)
    assert int(cnt[1]) == 703 * 25 and int(cnt[0]) == 172 * 25
