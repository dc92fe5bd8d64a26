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

from conftest import GOLDEN, load_golden_json
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
    L.host_mccfr_inplace.argtypes = [C.c_longlong, C.c_ulonglong, C.c_ulonglong]
    L.host_mccfr_inplace_tree.argtypes = [C.c_longlong, C.c_ulonglong, C.c_ulonglong]
    L.host_mccfr_batch.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_ulonglong, C.c_ulonglong]
    L.host_solver_counters.argtypes = [vp, vp, C.c_int]
    L.host_solver_delta_abs_sum.restype = C.c_double
    L.host_cfr_many.argtypes = [C.c_int, C.c_int]
    L.host_best_response.argtypes = [C.c_int, vp]
    L.host_policy.argtypes = [C.c_int, vp]
    L.host_eval.argtypes = [vp, vp, C.c_longlong, C.c_ulonglong, C.c_ulonglong, vp, vp]
    L.host_mccfr_inplace_many.argtypes = [C.c_int, C.c_longlong, C.c_ulonglong, C.c_ulonglong, vp, vp, vp]
    L.host_apply_peers.argtypes = [vp, vp, vp, vp, C.c_int]
    L.host_batch_peers.argtypes = [vp, vp, vp, vp, vp, C.c_longlong, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_int]
    L.host_solver_delta.argtypes = [vp]
    L.host_solver_set_delta.argtypes = [vp]
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


def test_cfr_many_kernel_is_cfr_run_per_job(lib):
    """cfr_many_kernel (one CTA per job, ms_cfr_iterate_many): here three CTAs, run one after another on the same
    table, two iterations each == six iterations of cfr_kernel."""
    a = HostSolver(lib, 43)
    a.cfr(6)
    want = a.table()
    b = HostSolver(lib, 43)
    assert lib.host_cfr_many(3, 2) == 0
    got = b.table()
    assert np.array_equal(got["regret"], want["regret"]) and np.array_equal(got["strategy"], want["strategy"])


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


def _counters(sv, reset=False):
    cnt, touched = np.zeros(3, np.uint64), np.zeros(sv.n_slots, np.uint8)
    sv.lib.host_solver_counters(cnt.ctypes.data, touched.ctypes.data, int(reset))
    return {"updates": int(cnt[0]), "visits": int(cnt[1]), "env_steps": int(cnt[2])}, touched


@pytest.mark.parametrize("kernel", ["mccfr_inplace_tree_kernel", "mccfr_inplace_kernel"])
def test_mccfr_inplace_matches_oracle_stream(lib, kernel):
    """The reference's sampled estimator with reference semantics (every update visible to the next node visit;
    mc_cfr.py:37-86) -- the tree-walking form MCCFRTrainer.iteration() runs by default and the form re-stepping the env --
    on the oracle's Philox stream: the same float64 bits, the same first-touch set as the reference's dict, and SURVEY
    3.2's per-iteration counts (703 calls, 172 updates)."""
    run = lib.host_mccfr_inplace_tree if kernel == "mccfr_inplace_tree_kernel" else lib.host_mccfr_inplace
    sv = HostSolver(lib, 42)
    strings = sv.table()["strings"]
    t = ora.Table()
    rng = ora.Rng(1, 777)
    done = 0
    for it in (1, 3, 25):
        assert run(it - done, 777, done) == 0
        t.mccfr_iterate(it - done, rng, first_iter=done)
        done = it
        tab = sv.table()
        cnt, touched = _counters(sv)
        keys, oreg, ostrat, _, _ = t.arrays()
        perm = _perm(strings, [k.split("|", 1)[1] for k in keys])
        assert int(touched.sum()) == len(keys) and touched[perm].all()
        assert np.array_equal(tab["regret"][perm], oreg) and np.array_equal(tab["strategy"][perm], ostrat)
    assert cnt["visits"] == 703 * 25 and cnt["updates"] == 172 * 25


@pytest.mark.parametrize("mode,player,ntrav", [(0, 0, 1), (0, 1, 1), (0, 0, 700), (0, 1, 700), (0, 2, 1500), (4, 0, 1), (4, 1, 700),
                                               (4, 2, 1200), (3, 0, 700), (3, 2, 900)])
def test_mccfr_batch_matches_oracle_frozen_sigma(lib, mode, player, ntrav):
    """mccfr_static_kernel (mode 0: the headline kernel of bench.py, sequential Philox stream), mccfr_tree_kernel<1024>
    (mode 4) and mccfr_batch_kernel (mode 3) (both: call-indexed stream), then mccfr_apply_kernel, against the oracle's
    frozen-sigma batch on the same stream and traversal ids: tables to 1e-9 (fp64 sums in another order; the static
    kernel multiplies by 1/sigma where the reference divides), update / visit / step counts exactly."""
    sv = HostSolver(lib, 42)
    strings = sv.table()["strings"]
    t = ora.Table()
    t.mccfr_populate()
    keys0, _, _, _, _ = t.arrays()
    perm = _perm(strings, [k.split("|", 1)[1] for k in keys0])
    assert lib.host_mccfr_inplace_tree(6, 9, 0) == 0           # a non-trivial common starting table
    t.mccfr_iterate(6, ora.Rng(1, 9))
    _counters(sv, reset=True)
    assert lib.host_mccfr_batch(mode, player, ntrav, 31337, 1000) == 0
    tab0 = sv.table()
    _, oreg0, ostrat0, _, _ = t.arrays()
    assert np.array_equal(tab0["regret"][perm], oreg0)          # deltas are not applied yet
    assert lib.host_mccfr_apply() == 0
    oreg, ostrat, nu, nv = oreg0.copy(), ostrat0.copy(), 0, 0
    for p in ((0, 1) if player == 2 else (player,)):            # both players against the SAME frozen table
        t.set_arrays(oreg0, ostrat0)
        u, v = (t.mccfr_batch_seq if mode == 0 else t.mccfr_batch)(p, 31337, 1000, ntrav)
        _, r1, s1, _, _ = t.arrays()
        oreg += r1 - oreg0
        ostrat += s1 - ostrat0
        nu, nv = nu + u, nv + v
    tab = sv.table()
    np.testing.assert_allclose(tab["regret"][perm], oreg, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(tab["strategy"][perm], ostrat, rtol=1e-9, atol=1e-9)
    cnt, _ = _counters(sv)
    assert (cnt["updates"], cnt["visits"]) == (nu, nv)
    assert cnt["env_steps"] == {0: 290, 1: 231, 2: 521}[player] * ntrav
    assert lib.host_solver_delta_abs_sum() == 0.0               # apply cleared the delta buffer


@pytest.mark.parametrize("mode,player,ntrav", [(1, 0, 600), (1, 2, 900), (2, 1, 2000), (2, 2, 2500)])
def test_textbook_estimators_match_oracle(lib, mode, player, ntrav):
    """External sampling on the tree (mode 1, mccfr_es_tree_kernel) and outcome sampling (mode 2, mccfr_os_kernel): the
    estimators the north star names and the reference lacks, against the C oracle on the same Philox stream."""
    sv = HostSolver(lib, 42)
    strings = sv.table()["strings"]
    t = ora.Table()
    t.mccfr_populate()
    keys0, _, _, _, _ = t.arrays()
    perm = _perm(strings, [k.split("|", 1)[1] for k in keys0])
    assert lib.host_mccfr_inplace_tree(5, 3, 0) == 0
    t.mccfr_iterate(5, ora.Rng(1, 3))
    _, r0, s0, _, _ = t.arrays()
    _counters(sv, reset=True)
    assert lib.host_mccfr_batch(mode, player, ntrav, 99, 40) == 0
    assert lib.host_mccfr_apply() == 0
    oreg, ostr, nu, nv = r0.copy(), s0.copy(), 0, 0
    for p in ((0, 1) if player == 2 else (player,)):
        t.set_arrays(r0, s0)
        u, v = t.mccfr_batch_mode(mode, p, 99, 40, ntrav)
        _, r1, s1, _, _ = t.arrays()
        oreg += r1 - r0
        ostr += s1 - s0
        nu, nv = nu + u, nv + v
    tab = sv.table()
    np.testing.assert_allclose(tab["regret"][perm], oreg, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(tab["strategy"][perm], ostr, rtol=1e-9, atol=1e-9)
    cnt, _ = _counters(sv)
    assert (cnt["updates"], cnt["visits"]) == (nu, nv)


def _exploitability(lib, kind):
    out = np.zeros(2)
    assert lib.host_best_response(kind, out.ctypes.data) == 0
    return (out[0] + out[1]) / 2.0


def test_best_response_vs_restated_openspiel(lib):
    """best_response_kernel against the documented restatement of open_spiel's exploitability (oracle/ms_exploit.py,
    values in tests/golden/policies_eval.json; third-party, parity unpinned) after CFR, and against the C oracle on an
    MCCFR table (policy kind 1: touched infosets with a 1e-12 threshold, uniform elsewhere)."""
    g = load_golden_json("policies_eval.json")
    sv = HostSolver(lib, 42)
    assert abs(_exploitability(lib, 2) - g["uniform"]) < 1e-12
    done = 0
    for it in (1, 2, 5, 10, 20, 50):
        sv.cfr(it - done)
        done = it
        assert abs(_exploitability(lib, 0) - g["cfr"][str(it)]) < 1e-9, it
    sv = HostSolver(lib, 42)
    t = ora.Table()
    assert lib.host_mccfr_inplace_tree(40, 4, 0) == 0
    t.mccfr_iterate(40, ora.Rng(1, 4))
    e_ora, _ = t.exploitability(1)
    assert abs(_exploitability(lib, 1) - e_ora) < 1e-9


def test_batched_policy_evaluation_matches_exact_expectation(lib):
    """policy_kernel + eval_kernel (evaluate_agent's episodes, vanilla_cfr.py:157-216, one thread per episode): the
    Monte-Carlo mean against the exact tree expectation of the same two policies."""
    sv = HostSolver(lib, 42)
    sv.cfr(30)
    S = sv.n_slots
    trained, uni = np.zeros((S, 4)), np.zeros((S, 4))
    assert lib.host_policy(0, trained.ctypes.data) == 0 and lib.host_policy(2, uni.ctypes.data) == 0
    assert np.allclose(trained.sum(1), 1) and np.allclose(uni.sum(1), 1)
    t = sv.tree()

    def exact(pol0, pol1):
        val = np.zeros(sv.n_nodes)
        for v in range(sv.n_nodes - 1, -1, -1):
            n, c0, w = int(t["nchild"][v]), int(t["child_begin"][v]), t["state"][v]
            if n == 0:
                s0 = bin(int(w[2]) & 0xFFFF).count("1") + 2 * ((int(w[3]) >> 4) & 0xF)
                s1 = bin(int(w[2]) >> 16).count("1") + 2 * ((int(w[3]) >> 8) & 0xF)
                val[v] = 0.5 * (s0 - s1)
            else:
                pr = (pol0 if (int(w[3]) >> 17) & 1 == 0 else pol1)[int(t["slot"][v]), :n]
                val[v] = float(np.dot(pr, val[c0:c0 + n]))
        return val[0]

    n = 60_000
    for a, b in ((trained, uni), (uni, trained), (uni, uni)):
        rew, sc = np.zeros(n, np.float32), np.zeros((n, 2), np.uint8)
        assert lib.host_eval(a.ctypes.data, b.ctypes.data, n, 12, 0, rew.ctypes.data, sc.ctypes.data) == 0
        r = rew.astype(np.float64)
        se = r.std() / np.sqrt(n)
        assert abs(r.mean() - exact(a, b)) < 5 * se + 1e-9, (exact(a, b), r.mean(), se)
        assert int(sc.max()) <= 4
    assert abs(exact(uni, uni) - (-0.9201)) < 1e-3             # on-policy value of the uniform profile


def test_mccfr_batch_shards_sum_to_whole(lib):
    """What the multi-GPU path relies on, on the headline kernel itself: traversals [0, n) split across
    ranks by traversal id give delta buffers whose SUM equals the single-rank buffer (same frozen table, same Philox
    ids) -- regret deltas to 1e-9 (fp64 sums in another order), update counts exactly."""
    lib.host_solver_delta.argtypes = [vp]

    def deltas(first, n):
        sv = HostSolver(lib, 42)
        assert lib.host_mccfr_inplace_tree(4, 2, 0) == 0          # the same starting table on every "rank"
        assert lib.host_mccfr_batch(0, 2, n, 5, first) == 0
        out = np.zeros(6 * sv.n_slots)
        lib.host_solver_delta(out.ctypes.data)
        return out, sv.n_slots

    whole, S = deltas(0, 3000)
    a, _ = deltas(0, 1500)
    b, _ = deltas(1500, 1500)
    np.testing.assert_allclose((a + b)[:5 * S], whole[:5 * S], rtol=1e-9, atol=1e-9)
    assert np.array_equal((a + b)[4 * S:5 * S], whole[4 * S:5 * S]) and whole[4 * S:5 * S].sum() > 0
    assert np.array_equal((a + b)[5 * S:] != 0, whole[5 * S:] != 0)       # first-touch marks: the same SET of infosets


def test_mccfr_inplace_many_runs_equal_solo_runs(lib):
    """mccfr_inplace_many_kernel (the reference's 10-independent-runs protocol, run_mccfr_experiment.py:195-202, as one
    launch, one warp per run): run r of the launch holds the same float64 bits and first-touch set as a solo
    mccfr_inplace_tree_kernel run on philox seed seed0 + r -- also when the launch is continued, and with more runs than
    one CTA holds (4 per CTA for this deal)."""
    sv = HostSolver(lib, 42)
    S, runs, seed0 = sv.n_slots, 6, 4000
    reg, strat, tch = np.zeros((runs, S, 4)), np.zeros((runs, S, 4)), np.zeros((runs, S), np.uint8)
    assert lib.host_mccfr_inplace_many(runs, 3, seed0, 0, reg.ctypes.data, strat.ctypes.data, tch.ctypes.data) == 0
    assert lib.host_mccfr_inplace_many(runs, 4, seed0, 3, reg.ctypes.data, strat.ctypes.data, tch.ctypes.data) == 0
    cnt, _ = _counters(sv)
    assert cnt["updates"] == 172 * 7 * runs and cnt["visits"] == 703 * 7 * runs
    assert np.array_equal(sv.table()["regret"], np.zeros((S, 4)))          # the solver's own table is not touched
    for r in range(runs):
        solo = HostSolver(lib, 42)
        assert lib.host_mccfr_inplace_tree(7, seed0 + r, 0) == 0
        tab = solo.table()
        _, touched = _counters(solo)
        assert np.array_equal(reg[r], tab["regret"]) and np.array_equal(strat[r], tab["strategy"])
        assert np.array_equal(tch[r], touched)
    assert not np.array_equal(reg[0], reg[1])                              # the runs are independent streams


def _shares(lib, S, first, n0, n1):
    """one batch split over two ranks on the emulated solver's CURRENT table: rank 1's share [first + n0, first + n0 + n1)
    is run first and moved out of the solver's delta buffer, then rank 0's share [first, first + n0) is run into it
    -> (delta0 copy, delta1)"""
    delta0, delta1 = np.zeros(6 * S), np.zeros(6 * S)
    assert lib.host_mccfr_batch(0, 2, n1, 99, first + n0) == 0
    lib.host_solver_delta(delta1.ctypes.data)
    lib.host_solver_set_delta(np.zeros(6 * S).ctypes.data)
    assert lib.host_mccfr_batch(0, 2, n0, 99, first) == 0
    lib.host_solver_delta(delta0.ctypes.data)
    return delta0, delta1


def test_apply_peers_two_emulated_ranks_equal_sum_and_apply(lib):
    """mccfr_apply_peers_kernel with two ranks running CONCURRENTLY in this process (flag barrier with release / acquire
    semantics, both ranks reading both delta buffers, rank-ordered sum, table update, first-touch marks): after every
    exchange the two replicas hold the same bits, and they equal {delta0 + delta1, mccfr_apply_kernel} -- two exchanges in
    a row (the epoch advances, the buffers are cleared)."""
    lib.host_peers_reset()
    sv = HostSolver(lib, 42)
    S = sv.n_slots
    assert lib.host_mccfr_inplace_tree(1, 5, 0) == 0           # one iteration: most infosets are still untouched
    start = sv.table()
    _, tch_start = _counters(sv)
    reg1, strat1, tch1 = start["regret"].copy(), start["strategy"].copy(), tch_start.copy()
    sums = []
    for first, n0, n1 in ((0, 700, 500), (1200, 300, 900)):
        delta0, delta1 = _shares(lib, S, first, n0, n1)
        sums.append(delta0 + delta1)
        assert lib.host_apply_peers(reg1.ctypes.data, strat1.ctypes.data, tch1.ctypes.data, delta1.ctypes.data, 0) == 0
        tab0 = sv.table()
        _, tch0 = _counters(sv)
        assert np.array_equal(tab0["regret"], reg1) and np.array_equal(tab0["strategy"], strat1)
        assert np.array_equal(tch0, tch1) and tch0.sum() > tch_start.sum()
        assert lib.host_solver_delta_abs_sum() == 0.0 and not delta1.any()
    ref = HostSolver(lib, 42)                                   # (the emulator holds ONE solver: this replaces `sv`)
    assert lib.host_mccfr_inplace_tree(1, 5, 0) == 0
    for d in sums:
        lib.host_solver_set_delta(d.ctypes.data)
        assert lib.host_mccfr_apply() == 0
    want = ref.table()
    _, tch_want = _counters(ref)
    assert np.array_equal(tab0["regret"], want["regret"]) and np.array_equal(tab0["strategy"], want["strategy"])
    assert np.array_equal(tch0, tch_want)


def test_fused_batch_and_exchange_kernel_two_emulated_ranks(lib):
    """mccfr_static_peers_kernel (ms_mccfr_batch_peers: the traversals and, in the last CTA to finish, the cross-rank
    exchange + table update -- ONE launch per iteration per GPU) with two emulated ranks of 3 CTAs each: both replicas end
    up with the same bits, equal to {each rank's batch, delta0 + delta1, mccfr_apply_kernel}; counters and first-touch
    flags included; two iterations in a row."""
    lib.host_peers_reset()
    sv = HostSolver(lib, 42)
    S = sv.n_slots
    assert lib.host_mccfr_inplace_tree(1, 5, 0) == 0
    start = sv.table()
    _, tch_start = _counters(sv, reset=True)
    reg1, strat1, tch1 = start["regret"].copy(), start["strategy"].copy(), tch_start.copy()
    delta1, cnt1 = np.zeros(6 * S), np.zeros(4, np.uint64)
    n = 1500
    for it in range(2):
        rc = lib.host_batch_peers(reg1.ctypes.data, strat1.ctypes.data, tch1.ctypes.data, delta1.ctypes.data, cnt1.ctypes.data,
                                  n, 99, (2 * it) * n, (2 * it + 1) * n, 3)
        assert rc == 0
        tab0 = sv.table()
        c0, tch0 = _counters(sv)
        assert np.array_equal(tab0["regret"], reg1) and np.array_equal(tab0["strategy"], strat1) and np.array_equal(tch0, tch1)
        assert lib.host_solver_delta_abs_sum() == 0.0 and not delta1.any()
    assert c0["updates"] == int(cnt1[0]) == 172 * n * 2 and c0["visits"] == int(cnt1[1]) == 703 * n * 2
    # the same two iterations as {batch of rank 0's ids, batch of rank 1's ids, apply} on one solver
    ref = HostSolver(lib, 42)
    assert lib.host_mccfr_inplace_tree(1, 5, 0) == 0
    for it in range(2):
        assert lib.host_mccfr_batch(0, 2, n, 99, (2 * it) * n) == 0
        assert lib.host_mccfr_batch(0, 2, n, 99, (2 * it + 1) * n) == 0
        assert lib.host_mccfr_apply() == 0
    want = ref.table()
    _, tch_want = _counters(ref)
    np.testing.assert_allclose(tab0["regret"], want["regret"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(tab0["strategy"], want["strategy"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(tch0, tch_want)


def test_apply_peers_gives_up_on_an_absent_rank(lib):
    """The barrier is bounded: when rank 1 never arrives, rank 0 stops waiting after MS_PEER_TIMEOUT_NS, reports
    1 + the missing rank in its error word and leaves its table untouched; later exchanges return at once."""
    import time
    lib.host_peers_reset()
    sv = HostSolver(lib, 42)
    S = sv.n_slots
    assert lib.host_mccfr_inplace_tree(2, 5, 0) == 0
    before = sv.table()
    _shares(lib, S, 0, 200, 100)
    junk = [np.zeros((S, 4)), np.zeros((S, 4)), np.zeros(S, np.uint8), np.zeros(6 * S)]
    t0 = time.perf_counter()
    assert lib.host_apply_peers(*[j.ctypes.data for j in junk], 1) == 2          # 1 + rank 1
    waited = time.perf_counter() - t0
    assert 1.5 < waited < 30.0
    after = sv.table()
    assert np.array_equal(before["regret"], after["regret"]) and np.array_equal(before["strategy"], after["strategy"])
    t0 = time.perf_counter()
    assert lib.host_apply_peers(*[j.ctypes.data for j in junk], 1) == 2          # sticky, and immediate
    assert time.perf_counter() - t0 < 1.0
    lib.host_peers_reset()
