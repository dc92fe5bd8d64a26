"""The PRODUCT's multi-deal MCCFR kernels on the CPU: md_mccfr_kernel<768>, md_apply_kernel,
md_export_kernel and md_lookup_kernel of scopa_b200/csrc/ms_multideal.cu run by the CTA emulator
(tests/emu/ms_multideal_host.cpp; the open-addressing infoset table lives in host memory, atomicCAS insertion and the
delta atomics are real atomics).  Same anchors as tests/test_gpu_multideal.py on the device:
  * one deal: the hash-table path reproduces the one-deal batch solver (mccfr_tree_kernel on the same emulator, itself
    checked against the reference-pinned oracle in test_solver_host.py) on the same Philox streams, 1e-9;
  * several deals: the oracle's restatement -- same infoset set, tables to 1e-9, same update / visit counts."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from oracle import ms_oracle as ora
from scopa_b200 import codec

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_build  # noqa: E402
from test_solver_host import HostSolver, _counters  # noqa: E402
from test_solver_host import lib as solver_lib  # noqa: E402,F401  (fixture)

vp = C.c_void_p


@pytest.fixture(scope="module")
def md():
    L = C.CDLL(emu_build.build_multideal_host())
    L.host_md_create.argtypes = [vp, vp, C.c_longlong, C.c_int]
    L.host_md_batch.argtypes = [C.c_int, C.c_longlong, C.c_ulonglong, C.c_ulonglong]
    L.host_md_counters.argtypes = [vp, C.c_int]
    L.host_md_export.argtypes = [vp, vp, vp, C.c_longlong]
    L.host_md_export.restype = C.c_longlong
    L.host_md_lookup.argtypes = [vp, C.c_longlong, vp, vp, vp]
    L.host_md_blocked.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_int, C.c_ulonglong]
    return L


def create(md, seeds, log2_capacity):
    roots, ho = [], []
    for s in seeds:
        cards = ora.deck(s)
        roots.append(codec.pack_state([codec.mask_of(cards[:4]), codec.mask_of(cards[4:8])], [], [0, 0], [0, 0], 0, 0, False, 8))
        ho.append(codec.pack_nibbles(cards[:8]))
    roots, ho = np.array(roots, dtype=np.uint32), np.array(ho, dtype=np.uint32)
    assert md.host_md_create(roots.ctypes.data, ho.ctypes.data, len(seeds), log2_capacity) == 0


def counters(md, reset=False):
    out = np.zeros(5, np.uint64)
    md.host_md_counters(out.ctypes.data, int(reset))
    assert int(out[4]) == 0, "table overflow / invariant flag"
    return {"updates": int(out[0]), "visits": int(out[1]), "env_steps": int(out[2]), "infosets": int(out[3])}


def export(md):
    n = counters(md)["infosets"]
    keys, reg, strat = np.zeros(max(n, 1), np.uint64), np.zeros((max(n, 1), 4)), np.zeros((max(n, 1), 4))
    got = md.host_md_export(keys.ctypes.data, reg.ctypes.data, strat.ctypes.data, n)
    assert got == n
    order = np.argsort(keys[:n], kind="stable")
    return keys[:n][order], reg[:n][order], strat[:n][order]


def _assert_tables_close(reg, oreg, strat, ostrat):
    """1e-9, except rows touched by an ill-conditioned importance weight (a regret that cancels to a rounding residue
    becomes a sampling probability of 1e-17 and a delta of 1e16; the residue depends on the order of the additions)."""
    wild = (np.abs(reg) > 1e9).any(1) | (np.abs(oreg) > 1e9).any(1)
    assert wild.mean() < 0.01
    np.testing.assert_allclose(reg[~wild], oreg[~wild], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(strat[~wild], ostrat[~wild], rtol=1e-9, atol=1e-9)


def test_one_deal_reproduces_batch_solver(md, solver_lib):  # noqa: F811
    sv = HostSolver(solver_lib, 42)
    create(md, [42], 12)
    n = 2048
    for b in range(3):
        for p in (0, 1):
            assert solver_lib.host_mccfr_batch(4, p, n, 5, b * n) == 0 and solver_lib.host_mccfr_apply() == 0
            assert md.host_md_batch(p, n, 5, b * n) == 0 and md.host_md_apply() == 0
    tab = sv.table()
    c1, touched = _counters(sv)
    S = sv.n_slots
    skeys = np.zeros(S, np.uint64)
    solver_lib.host_solver_export(skeys.ctypes.data, None, None, None, None)
    keys, mreg, mstrat = export(md)
    stored = touched.astype(bool) & (tab["nlegal"] > 1)          # one-card infosets are not materialised (sigma = [1])
    c2 = counters(md)
    assert c2["infosets"] == len(keys) == int(stored.sum())
    assert (c1["updates"], c1["visits"]) == (c2["updates"], c2["visits"])
    pos = {int(k): i for i, k in enumerate(keys)}
    worst = 0.0
    for s in range(S):
        if not stored[s]:
            assert int(skeys[s]) not in pos
            continue
        i = pos[int(skeys[s])]
        nl = int(tab["nlegal"][s])
        hand = sorted(int(c) for c in tab["legal"][s][:nl])
        for a in range(nl):                                      # table columns are in ascending card id
            col = hand.index(int(tab["legal"][s][a]))
            worst = max(worst, abs(mreg[i, col] - tab["regret"][s, a]) / max(1.0, abs(tab["regret"][s, a])),
                        abs(mstrat[i, col] - tab["strategy"][s, a]) / max(1.0, abs(tab["strategy"][s, a])))
    assert worst < 1e-9, worst


@pytest.mark.parametrize("player", [2, 0])
def test_many_deals_match_oracle(md, player):
    seeds = [42, 1, 43, 7, 2 ** 33 + 7, 12345, 99, 1000]
    create(md, seeds, 15)
    om = ora.MultiDealTable(seeds)
    n, nu, nv = 1500, 0, 0
    for b in range(3):
        assert md.host_md_batch(player, n, 9, b * n) == 0 and md.host_md_apply() == 0
        u, v = om.batch(player, 9, b * n, n)
        nu, nv = nu + u, nv + v
        om.apply()
    c = counters(md)
    assert (c["updates"], c["visits"]) == (nu, nv)
    keys, reg, strat = export(md)
    _, okeys, oreg, ostrat, _ = om.arrays()
    order = np.argsort(okeys, kind="stable")
    assert np.array_equal(keys, okeys[order]), "infoset sets differ"
    assert len(keys) > 2000 and c["infosets"] == len(keys)
    _assert_tables_close(reg, oreg[order], strat, ostrat[order])
    # lookups: present keys return the same rows, an absent key is reported
    q = np.concatenate([keys[:100], np.array([(1 << 52) | (0xF << 36)], dtype=np.uint64)])
    lreg, lstrat, found = np.zeros((101, 4)), np.zeros((101, 4)), np.zeros(101, np.uint8)
    assert md.host_md_lookup(q.ctypes.data, 101, lreg.ctypes.data, lstrat.ctypes.data, found.ctypes.data) == 0
    assert found.tolist() == [1] * 100 + [0]
    assert np.array_equal(lreg[:100], reg[:100]) and np.array_equal(lstrat[:100], strat[:100])


def test_blocked_one_deal_reproduces_batch_solver(md, solver_lib):  # noqa: F811
    """Deal-blocked form (md_build_kernel + md_blocked_kernel), one deal: visit b with n traversal pairs ==
    mccfr_static_kernel (the one-deal solver's headline kernel: same walk, same sequential Philox stream) on ids
    [n b, n (b + 1))."""
    sv = HostSolver(solver_lib, 42)
    create(md, [42], 12)
    n = 2048
    for b in range(3):
        for p in (0, 1):
            assert solver_lib.host_mccfr_batch(0, p, n, 5, b * n) == 0 and solver_lib.host_mccfr_apply() == 0
            assert md.host_md_blocked(p, b, 1, n, 5) == 0 and md.host_md_apply() == 0
    tab = sv.table()
    c1, _ = _counters(sv)
    skeys = np.zeros(sv.n_slots, np.uint64)
    solver_lib.host_solver_export(skeys.ctypes.data, None, None, None, None)
    keys, mreg, mstrat = export(md)
    multi = tab["nlegal"] > 1
    c2 = counters(md)
    assert c2["infosets"] == len(keys) == int(multi.sum())        # every multi-action infoset of the deal exists
    assert (c1["updates"], c1["visits"]) == (c2["updates"], c2["visits"])
    pos = {int(k): i for i, k in enumerate(keys)}
    worst = 0.0
    for s in np.nonzero(multi)[0]:
        i = pos[int(skeys[s])]
        nl = int(tab["nlegal"][s])
        hand = sorted(int(c) for c in tab["legal"][s][:nl])
        for a in range(nl):
            col = hand.index(int(tab["legal"][s][a]))
            worst = max(worst, abs(mreg[i, col] - tab["regret"][s, a]) / max(1.0, abs(tab["regret"][s, a])),
                        abs(mstrat[i, col] - tab["strategy"][s, a]) / max(1.0, abs(tab["strategy"][s, a])))
    assert worst < 1e-9, worst


@pytest.mark.parametrize("player", [2, 0])
def test_blocked_many_deals_match_oracle(md, player):
    seeds = [42, 1, 43, 7, 2 ** 33 + 7, 12345, 99, 1000, 5, 6, 8, 9]
    create(md, seeds, 15)
    om = ora.MultiDealTable(seeds)
    om.populate()
    nu, nv = 0, 0
    for b in range(3):
        assert md.host_md_blocked(player, 7 * b, 7, 300, 21) == 0 and md.host_md_apply() == 0
        u, v = om.batch_blocked(player, 21, 7 * b, 7, 300)
        nu, nv = nu + u, nv + v
        om.apply()
    c = counters(md)
    assert (c["updates"], c["visits"]) == (nu, nv)
    keys, reg, strat = export(md)
    _, okeys, oreg, ostrat, _ = om.arrays()
    order = np.argsort(okeys, kind="stable")
    assert np.array_equal(keys, okeys[order]), "infoset sets differ"
    _assert_tables_close(reg, oreg[order], strat, ostrat[order])
    assert np.abs(reg).sum() > 0 and c["infosets"] == len(keys)
    # the in-place-table kernel keeps working on the same table afterwards (slots, dirty bits and deltas are shared)
    assert md.host_md_batch(player, 500, 2, 10 ** 6) == 0 and md.host_md_apply() == 0
    om.batch(player, 2, 10 ** 6, 500)
    om.apply()
    keys2, reg2, strat2 = export(md)
    _, okeys2, oreg2, ostrat2, _ = om.arrays()
    order2 = np.argsort(okeys2, kind="stable")
    assert np.array_equal(keys2, okeys2[order2])
    _assert_tables_close(reg2, oreg2[order2], strat2, ostrat2[order2])


# ---- the table sharded over several ranks ("shard by hash(key) % G"), ranks emulated in one process ----

def create_world(md, seeds, log2_capacity, world):
    md.host_md_world_create.argtypes = [vp, vp, C.c_longlong, C.c_int, C.c_int]
    roots, ho = [], []
    for s in seeds:
        cards = ora.deck(s)
        roots.append(codec.pack_state([codec.mask_of(cards[:4]), codec.mask_of(cards[4:8])], [], [0, 0], [0, 0], 0, 0, False, 8))
        ho.append(codec.pack_nibbles(cards[:8]))
    roots, ho = np.array(roots, dtype=np.uint32), np.array(ho, dtype=np.uint32)
    assert md.host_md_world_create(roots.ctypes.data, ho.ctypes.data, len(seeds), log2_capacity, world) == 0


def export_shard(md):
    """the selected rank's shard (ms_md_export with max_n = 0 counts)"""
    dummy = np.zeros(4, np.uint64)
    n = md.host_md_export(dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, 0)
    keys, reg, strat = np.zeros(max(n, 1), np.uint64), np.zeros((max(n, 1), 4)), np.zeros((max(n, 1), 4))
    assert md.host_md_export(keys.ctypes.data, reg.ctypes.data, strat.ctypes.data, n) == n
    return keys[:n], reg[:n], strat[:n]


def export_world(md, world):
    parts = []
    for r in range(world):
        assert md.host_md_select(r) == 0
        parts.append(export_shard(md))
    keys = np.concatenate([p[0] for p in parts])
    assert len(np.unique(keys)) == len(keys), "an infoset lives in two shards"
    order = np.argsort(keys, kind="stable")
    return (keys[order], np.concatenate([p[1] for p in parts])[order], np.concatenate([p[2] for p in parts])[order],
            [len(p[0]) for p in parts])


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_table_matches_oracle(md, world):
    """`world` ranks, each with its shard of the table and its share of the visits of every iteration: blocked traversal
    (gather / REDs through the peers' shards) ; barrier ; apply on the own shard ; barrier.  The union of the shards
    equals the oracle's single table (same infoset set, 1e-9) -- i.e. the result does not depend on the number of ranks."""
    from scopa_b200.sharding import shard_bounds
    seeds = [42, 1, 43, 7, 2 ** 33 + 7, 12345, 99, 1000, 5, 6, 8, 9]
    create_world(md, seeds, 14, world)
    om = ora.MultiDealTable(seeds)
    om.populate()
    nu = nv = 0
    for b in range(3):
        for r in range(world):
            lo, n = shard_bounds(7, r, world)
            assert md.host_md_select(r) == 0
            assert md.host_md_blocked(2, 7 * b + lo, n, 300, 21) == 0
        assert md.host_md_barrier_all(-1) == 0
        for r in range(world):
            assert md.host_md_select(r) == 0 and md.host_md_apply() == 0
        assert md.host_md_barrier_all(-1) == 0
        u, v = om.batch_blocked(2, 21, 7 * b, 7, 300)
        nu, nv = nu + u, nv + v
        om.apply()
    tot = {"updates": 0, "visits": 0, "infosets": 0}
    for r in range(world):
        md.host_md_select(r)
        c = counters(md)
        assert md.host_md_peer_error(r) == 0
        for k in tot:
            tot[k] += c[k]
    assert (tot["updates"], tot["visits"]) == (nu, nv)
    keys, reg, strat, sizes = export_world(md, world)
    assert min(sizes) > 0.6 * len(keys) / world, sizes            # the owner hash spreads the infosets
    _, okeys, oreg, ostrat, _ = om.arrays()
    order = np.argsort(okeys, kind="stable")
    assert np.array_equal(keys, okeys[order]), "infoset sets differ"
    assert tot["infosets"] == len(keys)                            # every infoset was created exactly once, by some rank
    _assert_tables_close(reg, oreg[order], strat, ostrat[order])
    assert np.abs(reg).sum() > 0
    # any rank can look any infoset up; the per-traversal kernel works on the sharded table too
    md.host_md_select(world - 1)
    q = np.ascontiguousarray(keys[::37])
    lreg, lstrat, found = np.zeros((len(q), 4)), np.zeros((len(q), 4)), np.zeros(len(q), np.uint8)
    assert md.host_md_lookup(q.ctypes.data, len(q), lreg.ctypes.data, lstrat.ctypes.data, found.ctypes.data) == 0
    assert found.all() and np.array_equal(lreg, reg[::37]) and np.array_equal(lstrat, strat[::37])
    for r in range(world):
        lo, n = shard_bounds(600, r, world)
        md.host_md_select(r)
        assert md.host_md_batch(2, n, 2, 10 ** 6 + lo) == 0
    for r in range(world):
        md.host_md_select(r)
        assert md.host_md_apply() == 0
    om.batch(2, 2, 10 ** 6, 600)
    om.apply()
    keys2, reg2, strat2, _ = export_world(md, world)
    _, okeys2, oreg2, ostrat2, _ = om.arrays()
    order2 = np.argsort(okeys2, kind="stable")
    assert np.array_equal(keys2, okeys2[order2])
    _assert_tables_close(reg2, oreg2[order2], strat2, ostrat2[order2])


def test_sharded_barrier_reports_an_absent_rank(md):
    """A rank that never reaches the barrier: the others give up after the time limit and record which one was missing
    (instead of spinning until a watchdog kills the GPU); the error is sticky."""
    create_world(md, [42, 1], 12, 3)
    assert md.host_md_barrier_all(-1) == 0
    assert [md.host_md_peer_error(r) for r in range(3)] == [0, 0, 0]
    assert md.host_md_barrier_all(1) == 0
    assert [md.host_md_peer_error(r) for r in range(3)] == [2, 0, 2]
    assert md.host_md_barrier_all(-1) == 0                         # ranks in error return at once
    assert md.host_md_peer_error(0) == 2
