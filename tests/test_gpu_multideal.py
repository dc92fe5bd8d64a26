"""GPU tests of the multi-deal MCCFR table, through the C ABI.

No reference solver plays more than one deal, so the anchors are:
  * one deal (seed 42): the HBM hash-table path must reproduce the one-deal batch solver (itself checked against the
    reference-pinned oracle in test_gpu_solver.py) on the same Philox streams, 1e-9;
  * several deals: the CPU oracle's restatement of the same specification (oracle/ms_oracle.c "multi-deal MCCFR"),
    same infoset set, tables to 1e-9, same update / visit counts;
  * learning: exploitability of the average policy in the chance-root game (restated best response, parity unpinned)
    falls with training;
  * a full table is reported as MS_ERR_CAPACITY, never a hang.
"""
import numpy as np
import pytest
import torch

from oracle import ms_exploit
from oracle import ms_oracle as ora
from scopa_b200 import _lib, multideal
from scopa_b200.solver import Solver

pytestmark = pytest.mark.gpu


def test_one_deal_reproduces_batch_solver():
    sv = Solver(seed=42)
    md = multideal.MultiDealSolver([42], log2_capacity=12)
    n = 4096
    for b in range(3):
        for p in (0, 1):
            sv.mccfr_batch(p, n, philox_seed=5, first_trav=b * n, mode=4)     # md_mccfr_kernel shares the call-indexed stream of mode 4
            sv.mccfr_apply()
            md.mccfr_batch(n, philox_seed=5, first_trav=b * n, player=p)
            md.apply()
    reg, strat, touched = sv.export()
    st = sv.static_table()
    keys, mreg, mstrat = md.export()
    stored = touched.astype(bool) & (st["nlegal"] > 1)          # one-card infosets are not materialised (sigma = [1])
    assert md.counters()["infosets"] == len(keys) == int(stored.sum())
    pos = {int(k): i for i, k in enumerate(keys)}
    c1, c2 = sv.counters(), md.counters()
    assert (c1["updates"], c1["visits"]) == (c2["updates"], c2["visits"])
    worst = 0.0
    for s in range(sv.n_slots):
        if not stored[s]:
            assert int(st["keys"][s]) not in pos
            continue
        i = pos[int(st["keys"][s])]
        hand = sorted(int(c) for c in st["legal"][s][:st["nlegal"][s]])
        for a in range(int(st["nlegal"][s])):
            col = hand.index(int(st["legal"][s][a]))
            worst = max(worst, abs(mreg[i, col] - reg[s, a]) / max(1.0, abs(reg[s, a])),
                        abs(mstrat[i, col] - strat[s, a]) / max(1.0, abs(strat[s, a])))
    assert worst < 1e-9, worst


@pytest.mark.parametrize("player", [2, 0])
def test_many_deals_match_oracle(player):
    seeds = [42, 1, 43, 7, 2 ** 33 + 7, 12345, 99, 1000]
    md = multideal.MultiDealSolver(seeds, log2_capacity=15)
    om = ora.MultiDealTable(seeds)
    n, nu, nv = 1500, 0, 0
    for b in range(3):
        md.mccfr_batch(n, philox_seed=9, first_trav=b * n, player=player)
        md.apply()
        u, v = om.batch(player, 9, b * n, n)
        nu, nv = nu + u, nv + v
        om.apply()
    c = md.counters()
    assert (c["updates"], c["visits"]) == (nu, nv)
    keys, reg, strat = md.export()
    _, okeys, oreg, ostrat, _ = om.arrays()
    order = np.argsort(okeys, kind="stable")
    assert np.array_equal(keys, okeys[order]), "infoset sets differ"
    assert len(keys) > 2000 and c["infosets"] == len(keys)
    _assert_tables_close(reg, oreg[order], strat, ostrat[order])
    # lookups: present keys return the same rows, an absent key is reported
    q = np.concatenate([keys[:100], np.array([(1 << 52) | (0xF << 36)], dtype=np.uint64)])
    lreg, lstrat, found = md.lookup(q)
    assert found.cpu().numpy().tolist() == [1] * 100 + [0]
    assert np.array_equal(lreg.cpu().numpy()[:100], reg[:100]) and np.array_equal(lstrat.cpu().numpy()[:100], strat[:100])


def test_blocked_one_deal_reproduces_batch_solver():
    """Deal-blocked form, one deal: visit b with 4096 traversal pairs == ms_mccfr_batch on ids [4096 b, 4096 (b + 1))."""
    sv = Solver(seed=42)
    md = multideal.MultiDealSolver([42], log2_capacity=12)
    n = 4096
    for b in range(3):
        for p in (0, 1):
            sv.mccfr_batch(p, n, philox_seed=5, first_trav=b * n)     # mode 0: the same walk and sequential stream as md_blocked_kernel
            sv.mccfr_apply()
            md.mccfr_blocked(1, pairs_per_visit=n, philox_seed=5, first_visit=b, player=p)
            md.apply()
    reg, strat, touched = sv.export()
    st = sv.static_table()
    keys, mreg, mstrat = md.export()
    multi = st["nlegal"] > 1
    assert md.counters()["infosets"] == len(keys) == int(multi.sum())     # every multi-action infoset of the deal exists
    c1, c2 = sv.counters(), md.counters()
    assert (c1["updates"], c1["visits"]) == (c2["updates"], c2["visits"])
    pos = {int(k): i for i, k in enumerate(keys)}
    worst = 0.0
    for s in np.nonzero(multi)[0]:
        i = pos[int(st["keys"][s])]
        hand = sorted(int(c) for c in st["legal"][s][:st["nlegal"][s]])
        for a in range(int(st["nlegal"][s])):
            col = hand.index(int(st["legal"][s][a]))
            worst = max(worst, abs(mreg[i, col] - reg[s, a]) / max(1.0, abs(reg[s, a])),
                        abs(mstrat[i, col] - strat[s, a]) / max(1.0, abs(strat[s, a])))
    assert worst < 1e-9, worst


def _assert_tables_close(reg, oreg, strat, ostrat):
    """1e-9 agreement, except where the estimator itself is ill-conditioned: its importance weight is
    opp_reach / own_sampling_prob with no floor, so a regret that cancels to a rounding residue (1e-14 instead of 0)
    becomes a sampling probability of 1e-17 and a delta of 1e16 -- and the residue depends on the order of the fp64
    additions, which differs between the CPU and the atomics of the GPU.  Rows touched by such a weight are skipped
    (they must be rare)."""
    wild = (np.abs(reg) > 1e9).any(1) | (np.abs(oreg) > 1e9).any(1)
    assert wild.mean() < 0.01
    np.testing.assert_allclose(reg[~wild], oreg[~wild], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(strat[~wild], ostrat[~wild], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("player", [2, 0])
def test_blocked_many_deals_match_oracle(player):
    seeds = [42, 1, 43, 7, 2 ** 33 + 7, 12345, 99, 1000, 5, 6, 8, 9]
    md = multideal.MultiDealSolver(seeds, log2_capacity=15)
    om = ora.MultiDealTable(seeds)
    om.populate()
    nu, nv = 0, 0
    for b in range(3):
        md.mccfr_blocked(7, pairs_per_visit=300, philox_seed=21, first_visit=7 * b, player=player)
        md.apply()
        u, v = om.batch_blocked(player, 21, 7 * b, 7, 300)
        nu, nv = nu + u, nv + v
        om.apply()
    c = md.counters()
    assert (c["updates"], c["visits"]) == (nu, nv)
    keys, reg, strat = md.export()
    _, okeys, oreg, ostrat, _ = om.arrays()
    order = np.argsort(okeys, kind="stable")
    assert np.array_equal(keys, okeys[order]), "infoset sets differ"
    _assert_tables_close(reg, oreg[order], strat, ostrat[order])
    assert np.abs(reg).sum() > 0 and c["infosets"] == len(keys)
    # the in-place-table kernel keeps working on the same table afterwards (slots, dirty bits and deltas are shared)
    md.mccfr_batch(500, philox_seed=2, first_trav=10 ** 6, player=player)
    md.apply()
    om.batch(player, 2, 10 ** 6, 500)
    om.apply()
    keys2, reg2, strat2 = md.export()
    _, okeys2, oreg2, ostrat2, _ = om.arrays()
    order2 = np.argsort(okeys2, kind="stable")
    assert np.array_equal(keys2, okeys2[order2])
    _assert_tables_close(reg2, oreg2[order2], strat2, ostrat2[order2])


class _ChanceRootState:
    """pyspiel.State protocol over the chance-root game (for the restated best response): the root draws a deal,
    below it the oracle's one-deal states; infoset strings list the hand by ascending card id."""

    def __init__(self, seeds, inner=None, deal=None):
        self.seeds, self.inner, self.deal = seeds, inner, deal

    def is_chance_node(self):
        return self.inner is None

    def chance_outcomes(self):
        return [(i, 1.0 / len(self.seeds)) for i in range(len(self.seeds))]

    def is_terminal(self):
        return self.inner is not None and self.inner.is_terminal()

    def current_player(self):
        return -1 if self.inner is None else self.inner.current_player()

    def legal_actions(self, player=None):
        return list(range(len(self.seeds))) if self.inner is None else self.inner.legal_actions()

    def child(self, a):
        if self.inner is None:
            return _ChanceRootState(self.seeds, ora.State(int(self.seeds[a])), a)
        return _ChanceRootState(self.seeds, self.inner.child(a), self.deal)

    def returns(self):
        return self.inner.rewards()

    def key(self, player):
        e = self.inner.s.env
        hand = sum(1 << e.hand[player][i] for i in range(e.nhand[player]))
        table = [e.table[i] for i in range(e.ntable)]
        return (player << 52) | (hand << 36) | (len(table) << 32) | sum(c << (4 * i) for i, c in enumerate(table))

    def information_state_string(self, player=None):
        player = self.inner.current_player() if player is None else player
        return multideal.key_string(self.key(player))

    def history_str(self):
        return f"{self.deal}:{self.inner.history_str()}" if self.inner is not None else "root"


class _Game:
    def __init__(self, seeds):
        self.seeds = seeds

    def new_initial_state(self):
        return _ChanceRootState(self.seeds)

    def num_players(self):
        return 2


class _TablePolicy:
    def __init__(self, md):
        self.md, self.cache = md, {}

    def action_probabilities(self, state):
        p = state.current_player()
        key = state.key(p)
        if key not in self.cache:
            self.cache[key] = self.md.average_policy([key])[0] if self.md is not None else None
        row = self.cache[key]
        hand = sorted(state.legal_actions())
        if row is None:
            return {a: 1.0 / len(hand) for a in hand}
        return {a: float(row[hand.index(a)]) for a in state.legal_actions()}


def test_exploitability_falls_in_the_chance_root_game():
    seeds = [42, 1, 43]
    game = _Game(seeds)
    uniform = ms_exploit.exploitability(game, _TablePolicy(None))
    md = multideal.MultiDealSolver(seeds)
    n, done, curve = 2048, 0, []
    for target in (1, 4, 16):
        while done < target:
            md.mccfr_batch(n, philox_seed=3, first_trav=done * n)
            md.apply()
            done += 1
        curve.append(ms_exploit.exploitability(game, _TablePolicy(md)))
    # the first batch plays the uniform strategy, so its average policy is uniform; then it falls
    assert abs(curve[0] - uniform) < 1e-12 and curve[2] < curve[1] < 0.5 * uniform, (uniform, curve)
    # the CPU oracle's tables for the same batches (ora.MultiDealTable, same Philox seed) give these values
    np.testing.assert_allclose([uniform, curve[1], curve[2]], [1.8680555555555554, 0.7795025146405892, 0.6164106587760572],
                               atol=1e-6)


def test_full_table_is_an_error_not_a_hang():
    md = multideal.MultiDealSolver(list(range(1, 9)), log2_capacity=10)     # 8 deals need ~2500 slots, 1024 given
    md.mccfr_batch(4096, philox_seed=1)
    with pytest.raises(_lib.MsError, match="full"):
        md.counters()
    torch.cuda.synchronize()


def test_argument_validation():
    with pytest.raises(ValueError):
        multideal.MultiDealSolver([])
    with pytest.raises(_lib.MsError):
        multideal.MultiDealSolver([1], log2_capacity=40)
    md = multideal.MultiDealSolver([1], log2_capacity=12)
    with pytest.raises(_lib.MsError):
        md.mccfr_batch(10, player=3)
    md.mccfr_batch(0)
    assert md.counters()["updates"] == 0 and md.table_bytes == 4096 * 128


def test_sharding_entry_points_in_a_world_of_one():
    """The sharded-table entry points (ms_md_ipc_export / _attach / _peer_barrier / _peer_error) on ONE GPU: a world of
    one rank attaches to itself (no IPC handle is opened), runs the same kernels with its barrier between them and must
    reproduce the unattached solver bit for bit; argument and state errors are reported, never a hang.  (Two ranks on two
    GPUs: tests/test_gpu_multigpu.py; two and three emulated ranks: tests/test_multideal_host.py.)"""
    import ctypes as C
    lib = _lib.load()
    seeds = [42, 1, 43, 7, 99]
    a = multideal.MultiDealSolver(seeds, log2_capacity=14)
    b = multideal.MultiDealSolver(seeds, log2_capacity=14)
    handles = (C.c_ubyte * 128)()
    assert lib.ms_md_ipc_export(a.h, handles) == 0
    assert lib.ms_md_ipc_attach(a.h, 0, 9, bytes(handles) * 9) == -2           # more ranks than MD_MAX_PEERS
    assert lib.ms_md_ipc_attach(a.h, 1, 1, bytes(handles)) == -2               # rank outside the world
    assert lib.ms_md_ipc_attach(a.h, 0, 1, bytes(handles)) == 0
    assert lib.ms_md_ipc_attach(a.h, 0, 1, bytes(handles)) == -4               # already attached
    with pytest.raises(_lib.MsError):
        a.reset()                                                              # peers would still be reading the shard
    a.rank, a.world, a.group = 0, 1, None
    for it in range(3):
        a.iterate_blocked(6, 256, philox_seed=4, first_visit=6 * it)           # blocked ; barrier ; apply ; barrier
        b.mccfr_blocked(6, 256, philox_seed=4, first_visit=6 * it)
        b.apply()
    assert a.peer_error() == 0
    ka, ra, sa = a.export()
    kb, rb, sb = b.export()
    assert np.array_equal(ka, kb)
    np.testing.assert_allclose(ra, rb, rtol=1e-12, atol=1e-12)                 # (CTAs flush their REDs in any order)
    np.testing.assert_allclose(sa, sb, rtol=1e-12, atol=1e-12)
    ks, _, _ = a.export_shard()
    assert np.array_equal(np.sort(ks), ka)
    # attaching after the first traversal is refused: the deal descriptions already name table slots
    assert lib.ms_md_ipc_attach(b.h, 0, 1, bytes(handles)) == -4
