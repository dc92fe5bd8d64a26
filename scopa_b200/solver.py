"""Device-resident one-deal solver: Python handle over the ms_solver C ABI (include/scopa_b200.h).

Owns, in HBM: the enumerated game tree of one deal, the slot-aligned infoset table (regret_sum /
strategy_sum, float64) and the delta buffer that multi-GPU runs all-reduce.  The drop-in trainer
classes (scopa_b200.algorithms) are thin views over this object.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, codec
from .batch import BatchedMiniScopa


class _DevArray:
    """CUDA array interface over a raw device pointer (no copy), for torch.as_tensor."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def deal(seed, device="cuda"):
    """-> (packed root state words [4] as python ints, hand_order int) for MiniScopaEnv.reset(seed)."""
    b = BatchedMiniScopa(device).reset([seed])
    st = b.states.cpu().numpy().view(np.uint32)[0]
    ho = int(b.hand_order.cpu().numpy().view(np.uint32)[0])
    return tuple(int(x) for x in st), ho


class Solver:
    def __init__(self, root_words=None, hand_order=None, seed=42, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MsError("scopa_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        if root_words is None:
            root_words, hand_order = deal(seed, self.device)
        self.root_words = tuple(int(x) & 0xFFFFFFFF for x in root_words)
        self.hand_order = int(hand_order) & 0xFFFFFFFF
        root = (C.c_uint32 * 4)(*self.root_words)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_solver_create(root, self.hand_order, C.byref(h)))
        self.h = h
        n, s, l = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(self.lib.ms_solver_counts(self.h, C.byref(n), C.byref(s), C.byref(l)))
        self.n_nodes, self.n_slots, self.n_levels = n.value, s.value, l.value
        self._static = None
        self._tree = None
        self._delta_t = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ms_solver_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _stream(self):
        return _lib.stream_ptr(torch.cuda.current_stream(self.device))

    # ------------------------------------------------------------------------------- table / tree
    def reset(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_solver_reset(self.h, self._stream()))

    def tree(self):
        if self._tree is None:
            N = self.n_nodes
            st = np.zeros((N, 4), dtype=np.uint32)
            parent = np.zeros(N, dtype=np.int32)
            level = np.zeros(N, dtype=np.uint8)
            slot = np.zeros(N, dtype=np.int32)
            cb = np.zeros(N, dtype=np.int32)
            nc = np.zeros(N, dtype=np.uint8)
            _lib.check(self.lib.ms_solver_export_tree(self.h, st.ctypes.data, parent.ctypes.data, level.ctypes.data,
                                                      slot.ctypes.data, cb.ctypes.data, nc.ctypes.data))
            self._tree = {"state": st, "parent": parent, "level": level, "slot": slot, "child_begin": cb, "nchild": nc}
        return self._tree

    def static_table(self):
        """keys (uint64), info strings, n_legal, legal ids [S,4], player, and the DFS first-visit order of
        the slots (the reference's dict insertion order)."""
        if self._static is None:
            S = self.n_slots
            keys = np.zeros(S, dtype=np.uint64)
            nl = np.zeros(S, dtype=np.uint8)
            legal = np.zeros((S, 4), dtype=np.uint8)
            with torch.cuda.device(self.device):
                _lib.check(self.lib.ms_solver_export_table(self.h, keys.ctypes.data, nl.ctypes.data, legal.ctypes.data,
                                                           None, None, None, self._stream()))
            strings = [codec.key_to_string(k, self.hand_order) for k in keys]
            player = ((keys >> np.uint64(52)) & np.uint64(1)).astype(np.int64)
            t = self.tree()
            order, seen = [], set()
            stack = [0]
            while stack:                       # depth-first, children in legal order
                v = stack.pop()
                s = int(t["slot"][v])
                if s >= 0 and s not in seen:
                    seen.add(s)
                    order.append(s)
                c0, n = int(t["child_begin"][v]), int(t["nchild"][v])
                stack.extend(range(c0 + n - 1, c0 - 1, -1))
            self._static = {"keys": keys, "strings": strings, "nlegal": nl, "legal": legal, "player": player,
                            "dfs_order": np.array(order, dtype=np.int64)}
        return self._static

    def export(self):
        """-> regret [S,4], strategy [S,4] (float64), touched [S] (uint8), host copies."""
        S = self.n_slots
        reg = np.zeros((S, 4))
        strat = np.zeros((S, 4))
        touched = np.zeros(S, dtype=np.uint8)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_solver_export_table(self.h, None, None, None, reg.ctypes.data, strat.ctypes.data,
                                                       touched.ctypes.data, self._stream()))
        return reg, strat, touched

    def import_table(self, regret=None, strategy=None):
        r = None if regret is None else np.ascontiguousarray(regret, dtype=np.float64)
        s = None if strategy is None else np.ascontiguousarray(strategy, dtype=np.float64)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_solver_import_table(self.h, None if r is None else r.ctypes.data,
                                                       None if s is None else s.ctypes.data, self._stream()))

    def delta_tensor(self):
        """torch float64 view [6*S] of the device delta buffer (regret deltas, update counts, first-touch marks):
        the thing a multi-GPU run all-reduces once per iteration."""
        if self._delta_t is None:
            pr, ps, pd = C.c_void_p(), C.c_void_p(), C.c_void_p()
            nt, nd = C.c_size_t(), C.c_size_t()
            _lib.check(self.lib.ms_solver_device_ptrs(self.h, C.byref(pr), C.byref(ps), C.byref(pd), C.byref(nt),
                                                      C.byref(nd)))
            self._delta_t = torch.as_tensor(_DevArray(pd.value, nd.value), device=self.device)
            self._regret_t = torch.as_tensor(_DevArray(pr.value, nt.value), device=self.device)
            self._strategy_t = torch.as_tensor(_DevArray(ps.value, nt.value), device=self.device)
        return self._delta_t

    def table_tensors(self):
        self.delta_tensor()
        return self._regret_t, self._strategy_t

    # ------------------------------------------------------------------------------- solvers
    def cfr_iterate(self, iters):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_cfr_iterate(self.h, int(iters), self._stream()))

    def cfr_traverse(self, player, reach_p0=1.0, reach_p1=1.0):
        v = C.c_double()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_cfr_traverse(self.h, int(player), float(reach_p0), float(reach_p1), C.byref(v),
                                                self._stream()))
        return v.value

    def mccfr_inplace(self, iters, philox_seed=0, first_iter=0):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_mccfr_inplace(self.h, int(iters), int(philox_seed), int(first_iter), self._stream()))

    def mccfr_batch(self, player, n_trav, philox_seed=0, first_trav=0, mode=0):
        """mode 0 = the reference's estimator (static-shape kernel on a fresh deal's tree), 1 = external sampling,
        2 = outcome sampling (textbook, opt-in), 3 = the reference's estimator re-stepping the env at every node,
        4 = the reference's estimator on the generic tree-walking kernel (3 and 4 produce the same tables)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_mccfr_batch_mode(self.h, int(mode), int(player), int(n_trav), int(philox_seed),
                                                    int(first_trav), self._stream()))

    def mccfr_apply(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_mccfr_apply(self.h, self._stream()))

    # ------------------------------------------------------------------------------- peer-memory exchange
    def attach_peers(self, group=None):
        """Map every rank's inboxes into this process (CUDA IPC over NVLink / NVSwitch).  Collective: all ranks of
        `group` must call it.  Afterwards use apply_peers() / mccfr_batch_peers() instead of {all_reduce, mccfr_apply}.
        A failure on ANY rank (no peer access, IPC refused) raises MsError on EVERY rank: the outcome is agreed on with
        an all-reduce, so no rank is left waiting in a collective the failing rank never joins."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        handle = (C.c_ubyte * 64)()
        offs = (C.c_uint64 * 3)()
        err = None
        try:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.ms_solver_ipc_export(self.h, handle, offs))
        except _lib.MsError as e:
            err = e
        everyone = [None] * world
        dist.all_gather_object(everyone, None if err else (bytes(handle), [int(o) for o in offs]), group=group)
        if err is None and all(x is not None for x in everyone):
            try:
                handles = b"".join(h for h, _ in everyone)
                flat = (C.c_uint64 * (3 * world))(*[o for _, oo in everyone for o in oo])
                with torch.cuda.device(self.device):
                    _lib.check(self.lib.ms_solver_ipc_attach(self.h, rank, world, handles, flat))
            except _lib.MsError as e:
                err = e
        ok = torch.tensor([0.0 if (err is not None or any(x is None for x in everyone)) else 1.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() < 1.0:
            raise _lib.MsError(f"attach_peers failed on at least one rank (this rank: {err or 'ok'})")
        self._peers = True

    def apply_peers(self):
        """barrier + sum of all ranks' deltas (rank order) + table update, one kernel per rank."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_mccfr_apply_peers(self.h, self._stream()))

    def mccfr_batch_peers(self, player, n_trav, philox_seed=0, first_trav=0):
        """mccfr_batch + apply_peers as ONE launch (the last CTA to finish its traversals does the exchange)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_mccfr_batch_peers(self.h, int(player), int(n_trav), int(philox_seed), int(first_trav),
                                                     self._stream()))

    def peer_error(self):
        """0, or 1 + the rank that did not arrive at a peer exchange within its time limit (synchronises)."""
        err = C.c_uint32(0)
        with torch.cuda.device(self.device):
            self.lib.ms_solver_peer_error(self.h, C.byref(err), self._stream())
        return int(err.value)

    def counters(self, reset=False):
        out = (C.c_uint64 * 3)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_solver_counters(self.h, out, 1 if reset else 0, self._stream()))
        return {"updates": int(out[0]), "visits": int(out[1]), "env_steps": int(out[2])}

    def best_response_values(self, policy_kind):
        out = (C.c_double * 2)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_best_response(self.h, int(policy_kind), out, self._stream()))
        return [out[0], out[1]]

    def average_policy(self, policy_kind):
        """-> torch float64 [S, 4]: the table's average policy per slot (device tensor)."""
        out = torch.empty((self.n_slots, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_solver_policy(self.h, int(policy_kind), out.data_ptr(), self._stream()))
        return out

    def uniform_policy(self):
        return self.average_policy(2)

    def policy_table_from_dict(self, mapping, key_fn, probs_fn):
        """[S, 4] policy tensor from a host dict of InfoNode-like objects; slots missing from the dict are
        uniform (LearnedCFRPolicy / ScopaLearnedPolicy fallbacks)."""
        st = self.static_table()
        tab = np.zeros((self.n_slots, 4))
        for s in range(self.n_slots):
            n = int(st["nlegal"][s])
            node = mapping.get(key_fn(int(st["player"][s]), st["strings"][s]))
            tab[s, :n] = probs_fn(node) if node is not None else 1.0 / n
        return torch.from_numpy(tab).to(self.device)

    def evaluate(self, policy_seat0, policy_seat1, n_games, philox_seed=0, first_game=0):
        """Play n_games episodes on the device -> (reward of player 0 [n] f32, scopas [n, 2] u8) CUDA tensors."""
        p0 = policy_seat0.to(dtype=torch.float64).contiguous()
        p1 = policy_seat1.to(dtype=torch.float64).contiguous()
        rew = torch.empty((n_games,), dtype=torch.float32, device=self.device)
        sc = torch.empty((n_games, 2), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_eval_policies(self.h, p0.data_ptr(), p1.data_ptr(), int(n_games), int(philox_seed),
                                                 int(first_game), rew.data_ptr(), sc.data_ptr(), self._stream()))
        return rew, sc

    def exploitability(self, policy_kind):
        """(sum_b BR_b(root) - utility_sum) / num_players with utility_sum = 0 (zero-sum game)."""
        v = self.best_response_values(policy_kind)
        return (v[0] + v[1]) / 2.0


def cfr_iterate_many(solvers, iters):
    """`iters` vanilla-CFR iterations on every solver (independent deals) in one launch, one CTA per deal."""
    if not solvers:
        return
    arr = (C.c_void_p * len(solvers))(*[s.h for s in solvers])
    with torch.cuda.device(solvers[0].device):
        _lib.check(solvers[0].lib.ms_cfr_iterate_many(arr, len(solvers), int(iters), solvers[0]._stream()))


class ManyRuns:
    """Tables of n independent reference-semantics MCCFR runs on one deal (device tensors [n, S, 4] / [n, S])."""

    def __init__(self, solver, n_runs):
        self.solver, self.n = solver, int(n_runs)
        S = solver.n_slots
        self.regret = torch.zeros((self.n, S, 4), dtype=torch.float64, device=solver.device)
        self.strategy = torch.zeros((self.n, S, 4), dtype=torch.float64, device=solver.device)
        self.touched = torch.zeros((self.n, S), dtype=torch.uint8, device=solver.device)
        self.iterations = 0


def mccfr_inplace_many(solver, runs, iters, philox_seed0=0):
    """MCCFRTrainer.iteration() x iters for every one of `runs` independent runs (an int = that many fresh tables, or a
    ManyRuns to continue), one launch, one warp per run; run r uses philox seed philox_seed0 + r.  -> ManyRuns"""
    m = runs if isinstance(runs, ManyRuns) else ManyRuns(solver, runs)
    with torch.cuda.device(solver.device):
        _lib.check(solver.lib.ms_mccfr_inplace_many(solver.h, m.n, int(iters), int(philox_seed0), m.iterations,
                                                    m.regret.data_ptr(), m.strategy.data_ptr(), m.touched.data_ptr(),
                                                    solver._stream()))
    m.iterations += int(iters)
    return m


def smoke_check(ora):
    """Used by __graft_entry__.smoke(): a few CFR iterations and an MCCFR batch against the oracle."""
    sv = Solver(seed=42, device="cuda:0")
    sv.cfr_iterate(3)
    reg, strat, _ = sv.export()
    t = ora.Table()
    t.cfr_train(3)
    keys, oreg, ostrat, _, _ = t.arrays()
    st = sv.static_table()
    idx = {k: i for i, k in enumerate(st["strings"])}
    perm = np.array([idx[k] for k in keys])
    assert np.array_equal(reg[perm], oreg) and np.array_equal(strat[perm], ostrat), "CFR tables differ from the oracle"
    sv.reset()
    sv.mccfr_batch(0, 256, philox_seed=5, first_trav=0)
    sv.mccfr_apply()
    reg, strat, _ = sv.export()
    t = ora.Table()
    t.mccfr_populate()
    t.mccfr_batch_seq(0, 5, 0, 256)      # the static-shape kernel consumes the sequential Philox stream
    keys, oreg, ostrat, _, _ = t.arrays()
    perm = np.array([idx[k.split("|", 1)[1]] for k in keys])
    assert np.allclose(reg[perm], oreg, rtol=1e-9, atol=1e-9), "MCCFR regret deltas differ from the oracle"
    assert np.allclose(strat[perm], ostrat, rtol=1e-9, atol=1e-9), "MCCFR strategy sums differ from the oracle"
    print("smoke: CFR (bit-exact) and MCCFR batch (1e-9) match the oracle")
