"""Multi-deal MCCFR: Python handle over the ms_md_* C ABI (include/scopa_b200.h).

The reference's MCCFRTrainer solves the one deal its game object fixes (seed 42,
src/algorithms/mc_cfr.py:88-92).  This object runs the same `_sample` estimator on a game whose root is a
uniform chance node over a list of deals, with ONE infoset table for all deals held in HBM (open
addressing on the 64-bit infoset key, one 128-byte line per infoset, nodes created on first touch).
Infosets are merged by information content; table columns follow the reference's per-card arrays compacted
to the hand: column k = k-th smallest card id of the hand.  SURVEY.md section 8(f) row 3.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, codec


def key_fields(key):
    """packed key -> (player, hand card ids ascending, table card ids in order)"""
    key = int(key)
    player = (key >> 52) & 1
    hand = [c for c in range(16) if (key >> (36 + c)) & 1]
    n = (key >> 32) & 0xF
    return player, hand, codec.nibbles(key & 0xFFFFFFFF, n)


def key_string(key):
    """The canonical text form of an infoset ("player|hand ids ascending|table ids in order")."""
    p, hand, table = key_fields(key)
    return f"{p}|{','.join(map(str, hand))}|{','.join(map(str, table))}"


def info_string(key):
    """The reference's information_state_string (openspiel_mini_scopa.py:86-95) with the hand listed by ascending
    card id (the reference lists it in deal order, which is not part of the information)."""
    p, hand, table = key_fields(key)
    return f"P{p}:H[{'-'.join(codec.CARD_STR[c] for c in hand)}]_T[{'-'.join(codec.CARD_STR[c] for c in table)}]"


class MultiDealSolver:
    def __init__(self, seeds, log2_capacity=None, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MsError("scopa_b200 runs on CUDA devices only (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        seeds = np.ascontiguousarray(seeds, dtype=np.int64).reshape(-1)
        if seeds.size == 0:
            raise ValueError("at least one deal is needed")
        self.seeds = seeds
        if log2_capacity is None:        # a deal has 738 infosets; keep the load factor under one half
            log2_capacity = max(12, int(np.ceil(np.log2(seeds.size * 738 * 2))))
        self.log2_capacity = int(log2_capacity)
        d_seeds = torch.from_numpy(seeds).to(self.device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_create(d_seeds.data_ptr(), seeds.size, self.log2_capacity, self._stream(), C.byref(h)))
            torch.cuda.current_stream(self.device).synchronize()    # d_seeds may be released after this
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ms_md_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _stream(self):
        return _lib.stream_ptr(torch.cuda.current_stream(self.device))

    @property
    def capacity(self):
        return 1 << self.log2_capacity

    @property
    def table_bytes(self):
        b = C.c_int64()
        _lib.check(self.lib.ms_md_info(self.h, None, None, C.byref(b)))
        return b.value

    def reset(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_reset(self.h, self._stream()))

    # ---- the table sharded over the GPUs of one box (one process per GPU; SURVEY.md 8(e), last sentence) ----

    def attach_peers(self, group=None):
        """Shard the infoset table over the ranks of `group`: infoset `key` then lives on rank owner(key) (a hash of
        the key), in that rank's table, and every rank maps every rank's table (CUDA IPC over NVLink / NVSwitch).
        Collective; every rank must have created its solver with the same seeds and capacity (= the capacity of ONE
        shard), and must call this before its first traversal.  A failure on any rank raises MsError on every rank."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        handles = (C.c_ubyte * 128)()
        err = None
        try:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.ms_md_ipc_export(self.h, handles))
        except _lib.MsError as e:
            err = e
        everyone = [None] * world
        dist.all_gather_object(everyone, None if err else bytes(handles), group=group)
        if err is None and all(x is not None for x in everyone):
            try:
                with torch.cuda.device(self.device):
                    _lib.check(self.lib.ms_md_ipc_attach(self.h, rank, world, b"".join(everyone)))
            except _lib.MsError as e:
                err = e
        ok = torch.tensor([0.0 if (err is not None or any(x is None for x in everyone)) else 1.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() < 1.0:
            raise _lib.MsError(f"MultiDealSolver.attach_peers failed on at least one rank (this rank: {err or 'ok'})")
        self.group, self.rank, self.world = group, rank, world

    def barrier(self):
        """Stream-ordered barrier across the attached ranks (bounded: see peer_error); a no-op on one GPU."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_peer_barrier(self.h, self._stream()))

    def peer_error(self):
        """0, or 1 + the rank that did not arrive at a barrier within its time limit (synchronises)."""
        err = C.c_uint32(0)
        with torch.cuda.device(self.device):
            self.lib.ms_md_peer_error(self.h, C.byref(err), self._stream())
        return int(err.value)

    def iterate_blocked(self, n_visits, pairs_per_visit=3072, philox_seed=0, first_visit=0, player=2):
        """One iteration of the deal-blocked solver over ALL ranks: this rank runs its share of the visits
        [first_visit, first_visit + n_visits) -- regrets gathered from, deltas sent to the owners' shards inside the
        kernel --, then every rank folds the deltas of its own shard in.  Same table as one GPU running all the visits."""
        from .sharding import shard_bounds
        lo, n = shard_bounds(int(n_visits), getattr(self, "rank", 0), getattr(self, "world", 1))
        self.mccfr_blocked(n, pairs_per_visit, philox_seed, first_visit + lo, player)
        self.barrier()
        self.apply()
        self.barrier()

    def export_shard(self):
        """-> keys, regret, strategy of the infosets in THIS rank's shard (unsorted numpy arrays)."""
        got = C.c_int64()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_export(self.h, None, None, None, 0, C.byref(got), self._stream()))
            n = got.value
            keys = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
            reg = torch.empty((max(n, 1), 4), dtype=torch.float64, device=self.device)
            strat = torch.empty((max(n, 1), 4), dtype=torch.float64, device=self.device)
            _lib.check(self.lib.ms_md_export(self.h, keys.data_ptr(), reg.data_ptr(), strat.data_ptr(), n, C.byref(got),
                                             self._stream()))
        return keys[:n].cpu().numpy().view(np.uint64), reg[:n].cpu().numpy(), strat[:n].cpu().numpy()

    def mccfr_batch(self, n_trav, philox_seed=0, first_trav=0, player=2):
        """Launch n_trav traversals (each: one sampled deal, traverser = player, or both when player == 2) against
        the strategies frozen at launch; deltas stay in the table until apply()."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_mccfr_batch(self.h, player, n_trav, philox_seed, first_trav, self._stream()))

    def mccfr_blocked(self, n_visits, pairs_per_visit=3072, philox_seed=0, first_visit=0, player=2):
        """Deal-blocked form: each visit draws one deal and runs `pairs_per_visit` traversals on it entirely on chip
        (the deal's tree and the strategies of its infosets staged in shared memory, deltas written back once).
        Traversal ids are visit * pairs_per_visit + i.  The first call describes every deal's tree and creates all of
        its multi-action infosets in the table."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_mccfr_blocked(self.h, player, first_visit, n_visits, pairs_per_visit, philox_seed,
                                                    self._stream()))

    def apply(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_apply(self.h, self._stream()))

    def counters(self, reset=False):
        """-> dict(updates, visits, env_steps, infosets); raises when the table overflowed (synchronises)."""
        out = (C.c_uint64 * 5)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_counters(self.h, out, int(reset), self._stream()))
        return {"updates": out[0], "visits": out[1], "env_steps": out[2], "infosets": out[3]}

    def export(self):
        """-> keys [n] uint64 (numpy, sorted), regret [n,4], strategy [n,4] float64 of every infoset in the table
        (on a sharded table: of all shards, gathered from every rank -- a collective)"""
        if getattr(self, "world", 1) > 1:
            import torch.distributed as dist
            parts = [None] * self.world
            dist.all_gather_object(parts, self.export_shard(), group=self.group)
            k = np.concatenate([p[0] for p in parts])
            order = np.argsort(k, kind="stable")
            return k[order], np.concatenate([p[1] for p in parts])[order], np.concatenate([p[2] for p in parts])[order]
        n = int(self.counters()["infosets"])
        keys = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        reg = torch.empty((max(n, 1), 4), dtype=torch.float64, device=self.device)
        strat = torch.empty((max(n, 1), 4), dtype=torch.float64, device=self.device)
        got = C.c_int64()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_export(self.h, keys.data_ptr(), reg.data_ptr(), strat.data_ptr(), n, C.byref(got),
                                             self._stream()))
        k = keys[:got.value].cpu().numpy().view(np.uint64)
        order = np.argsort(k, kind="stable")
        return k[order], reg[:got.value].cpu().numpy()[order], strat[:got.value].cpu().numpy()[order]

    def lookup(self, keys):
        """keys (uint64 array-like) -> regret [n,4], strategy [n,4], found [n] (torch tensors on the device)"""
        k = np.ascontiguousarray(keys, dtype=np.uint64).view(np.int64)
        d_k = torch.from_numpy(k).to(self.device)
        n = d_k.numel()
        reg = torch.empty((n, 4), dtype=torch.float64, device=self.device)
        strat = torch.empty((n, 4), dtype=torch.float64, device=self.device)
        found = torch.empty(n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_md_lookup(self.h, d_k.data_ptr(), n, reg.data_ptr(), strat.data_ptr(), found.data_ptr(),
                                             self._stream()))
        return reg, strat, found

    def average_policy(self, keys):
        """ScopaLearnedPolicy's rule (mc_cfr.py:118-130) per queried infoset: strategy_sum normalised when its total
        exceeds 1e-12, uniform over the hand otherwise.  -> [n,4] numpy, columns = hand cards by ascending id."""
        _, strat, _ = self.lookup(keys)
        s = strat.cpu().numpy()
        out = np.zeros_like(s)
        for i, key in enumerate(np.asarray(keys, dtype=np.uint64)):
            n = bin((int(key) >> 36) & 0xFFFF).count("1")
            tot = s[i, :n].sum()
            out[i, :n] = s[i, :n] / tot if tot > 1e-12 else 1.0 / max(n, 1)
        return out
