"""Drop-in for the reference's src/algorithms/mc_cfr.py; sampling and updates run in the CUDA solver.

MCCFRTrainer keeps the reference's surface (/root/reference/src/algorithms/mc_cfr.py:27-102):
`iteration()`, `train(iterations)` (returns []), `info_sets` (dict (player, info-string) -> InfoNode),
`tabular_policy()`.  Two execution modes:
  * default ("in-place"): the reference's semantics -- one traversal per player per iteration, every
    regret update visible to the next node visit -- run by one device thread per launch;
  * `traversals_per_iteration=B` (batched): B traversals per player against a frozen table, one
    all-reduce of the slot-aligned delta buffer across `process_group` (if given), then table += delta.
The reference samples from numpy's global Mersenne Twister; here the stream is counter-based Philox
(seed `seed`), so runs agree with the reference statistically, not draw for draw.
"""
from dataclasses import dataclass

import numpy as np

from ..sharding import allreduce_delta, shard_bounds
from ..solver import Solver
from ._policy_base import Policy, root_of
from ._evaluate import evaluate_agent  # noqa: F401  (one implementation for both trainers)


@dataclass
class InfoNode:
    legal_actions: np.ndarray
    regret_sum: np.ndarray = None
    strategy_sum: np.ndarray = None

    def __post_init__(self):
        n = self.legal_actions.size
        if self.regret_sum is None:
            self.regret_sum = np.zeros(n)
        if self.strategy_sum is None:
            self.strategy_sum = np.zeros(n)

    def current_strategy(self):
        pos = np.maximum(self.regret_sum, 0)
        if pos.sum() == 0:
            return np.ones_like(pos) / len(pos)
        return pos / pos.sum()


class MCCFRTrainer:
    ESTIMATORS = {"reference": 0, "external": 1, "outcome": 2}

    def __init__(self, game, seed=0, traversals_per_iteration=None, process_group=None, device="cuda",
                 peer_memory=False, estimator="reference"):
        """estimator: "reference" = the estimator of the reference's _sample (default); "external" / "outcome" =
        textbook external / outcome sampling MCCFR (batched mode only: set traversals_per_iteration)."""
        self.game = game
        self.mode = self.ESTIMATORS[estimator]
        if self.mode and traversals_per_iteration is None:
            raise ValueError("the textbook estimators run in batched mode: pass traversals_per_iteration")
        words, order = root_of(game)
        self.solver = Solver(words, order, device=device)
        self.seed = int(seed)
        self.batch = traversals_per_iteration
        self.process_group = process_group
        self.peer_memory = bool(peer_memory) and traversals_per_iteration is not None
        if self.peer_memory:          # exchange deltas through NVLink peer memory instead of an NCCL all-reduce
            self.solver.attach_peers(process_group)
        self._iter = 0
        self._map = {}
        self._dirty = False

    @property
    def info_sets(self):
        if self._dirty:
            self._refresh()
        return self._map

    def _refresh(self):
        if self.peer_memory and self.solver.peer_error():
            raise RuntimeError("peer-memory exchange failed: a rank did not arrive (see ms_solver_peer_error)")
        st = self.solver.static_table()
        reg, strat, touched = self.solver.export()
        m = {}
        for s in st["dfs_order"]:
            if not touched[s]:
                continue                      # the reference creates a node on first touch only (:52)
            n = int(st["nlegal"][s])
            m[(int(st["player"][s]), st["strings"][s])] = InfoNode(
                np.array([int(a) for a in st["legal"][s, :n]]), reg[s, :n].copy(), strat[s, :n].copy())
        self._map = m
        self._dirty = False

    def _world(self):
        if self.process_group is None:
            return 0, 1
        import torch.distributed as dist
        return dist.get_rank(self.process_group), dist.get_world_size(self.process_group)

    def iterate(self, iterations):
        """`iterations` iterations in as few launches as possible."""
        if iterations <= 0:
            return
        if self.batch is None:
            self.solver.mccfr_inplace(iterations, philox_seed=self.seed, first_iter=self._iter)
            self._iter += iterations
        else:
            rank, world = self._world()
            B = int(self.batch)
            lo, n = shard_bounds(B, rank, world)
            for _ in range(iterations):
                if self.peer_memory and self.mode == 0:     # traversals + exchange + apply: one launch per iteration
                    self.solver.mccfr_batch_peers(2, n, philox_seed=self.seed, first_trav=self._iter * B + lo)
                    self._iter += 1
                    continue
                self.solver.mccfr_batch(2, n, philox_seed=self.seed, first_trav=self._iter * B + lo, mode=self.mode)
                if self.peer_memory:
                    self.solver.apply_peers()
                else:
                    if world > 1:
                        allreduce_delta(self.solver.delta_tensor(), self.process_group)
                    self.solver.mccfr_apply()
                self._iter += 1
        self._dirty = True

    def iteration(self):
        """Run a single iteration of MCCFR (one pass for each player)."""
        self.iterate(1)

    def train(self, iterations=10000):
        self.iterate(iterations)
        return []

    def exploitability(self):
        return self.solver.exploitability(1)

    def tabular_policy(self):
        pol = ScopaLearnedPolicy(self.game, self.info_sets)
        pol._solver = self.solver          # lets evaluate_agent play the episodes on the GPU
        return pol


def _mccfr_probs(node):
    total = node.strategy_sum.sum()
    if total > 1e-12:
        return node.strategy_sum / total
    return np.ones(len(node.legal_actions)) / len(node.legal_actions)


class ScopaLearnedPolicy(Policy):
    def __init__(self, game, info_sets):
        super().__init__(game, list(range(game.num_players())))
        self.info_sets = info_sets
        self._solver = None

    def _device_table(self, solver):
        return solver.policy_table_from_dict(self.info_sets, lambda p, s: (p, s), _mccfr_probs)

    def action_probabilities(self, state):
        if state.is_terminal():
            return {}
        player = state.current_player()
        key = (player, state.information_state_string(player))
        if key in self.info_sets:
            node = self.info_sets[key]
            total = node.strategy_sum.sum()
            if total > 1e-12:
                probs = node.strategy_sum / total
            else:
                probs = np.ones(len(node.legal_actions)) / len(node.legal_actions)
            return {action: probs[i] for i, action in enumerate(node.legal_actions)}
        legal = state.legal_actions(player)
        prob = 1.0 / len(legal)
        return {action: prob for action in legal}


class RandomPolicy(Policy):
    """Policy that chooses actions uniformly at random."""

    def __init__(self, game):
        super().__init__(game, list(range(game.num_players())))

    def _device_table(self, solver):
        return solver.uniform_policy()

    def action_probabilities(self, state):
        if state.is_terminal():
            return {}
        player = state.current_player()
        legal_actions = state.legal_actions(player)
        prob = 1.0 / len(legal_actions)
        return {action: prob for action in legal_actions}
