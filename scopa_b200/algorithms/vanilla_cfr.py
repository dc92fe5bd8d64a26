"""Vanilla CFR behind the reference's names (drop-in for /root/reference/src/algorithms/vanilla_cfr.py):
`InfoNode`, `CFRTrainer`, `LearnedCFRPolicy`, `RandomPolicy`, `evaluate_agent`.

The traversal itself (`CFRTrainer._cfr_recursive`, reference :56-99) runs in the CUDA solver
(csrc/ms_solver.cu, cfr_kernel: float64, bit-identical tables); this module is the host view:
`info_set_map` is re-formatted lazily from the device table into the reference's dict of `InfoNode`s
(info string -> regret_sum / strategy_sum / local_strategy, in first-visit order).
"""
from dataclasses import dataclass, field

import numpy as np

from ..solver import Solver
from ._evaluate import evaluate_agent  # noqa: F401  (re-exported, as in the reference module)
from ._policy_base import Policy, root_id, root_of


def _uniform(n):
    return np.ones(n) / n


def _regret_matching(regrets):
    pos = np.maximum(regrets, 0)
    total = np.sum(pos)
    return pos / total if total > 0 else _uniform(regrets.size)


@dataclass
class InfoNode:
    """Per-infoset accumulators, arrays indexed by position in `legal_actions` (hand order)."""
    legal_actions: np.ndarray
    regret_sum: np.ndarray = None
    strategy_sum: np.ndarray = None
    local_strategy: np.ndarray = None

    def __post_init__(self):
        n = self.legal_actions.size
        self.regret_sum = np.zeros(n) if self.regret_sum is None else self.regret_sum
        self.strategy_sum = np.zeros(n) if self.strategy_sum is None else self.strategy_sum
        self.local_strategy = _uniform(n) if self.local_strategy is None else self.local_strategy

    def get_strategy(self):
        return _regret_matching(self.regret_sum)

    @property
    def policy(self) -> np.ndarray:
        total = np.sum(self.strategy_sum)
        return self.strategy_sum / total if total > 0 else _uniform(self.legal_actions.size)


class CFRTrainer:
    """`train(steps, eval_interval, compute_exploitability)`, `_cfr_recursive(state, player, r0, r1)`,
    `info_set_map`, `get_openspiel_policy()` -- the reference's surface (:41-120) over a device-resident table."""

    def __init__(self, game, device="cuda"):
        self.game = game
        words, order = root_of(game)
        self._root = root_id(words, order)
        self.solver = Solver(words, order, device=device)
        self._view, self._stale = {}, False

    # ---- host view ----------------------------------------------------------------------------------------
    @property
    def info_set_map(self):
        if self._stale:
            st = self.solver.static_table()
            reg, strat, _ = self.solver.export()
            view = {}
            for s in st["dfs_order"]:                       # the reference dict's insertion order
                n = int(st["nlegal"][s])
                node = InfoNode(st["legal"][s, :n].astype(np.int64), reg[s, :n].copy(), strat[s, :n].copy())
                node.local_strategy = node.get_strategy()   # holds after every visit in the reference (:97)
                view[st["strings"][s]] = node
            self._view, self._stale = view, False
        return self._view

    def _get_or_create_node(self, info_set_key, legal_actions) -> InfoNode:
        return self.info_set_map.setdefault(info_set_key, InfoNode(np.array(legal_actions)))

    # ---- solver calls -------------------------------------------------------------------------------------
    def _cfr_recursive(self, state, traversing_player, reach_p0, reach_p1):
        """One traversal from `state` for `traversing_player`.  The reference's callers only ever pass a fresh
        root (`train`, and run_vanilla_cfr_experiment.py:87-91); other states are refused."""
        if state.is_terminal():
            return state.rewards()[traversing_player]
        words, order = state.env.packed()
        if root_id(words, order) != self._root:
            raise NotImplementedError("_cfr_recursive on a non-root state: build a CFRTrainer for a game rooted there")
        self._stale = True
        return self.solver.cfr_traverse(traversing_player, reach_p0, reach_p1)

    def train(self, steps: int, eval_interval: int = 1000, compute_exploitability: bool = False):
        history, done = [], 0
        while done < steps:
            n = steps - done
            if compute_exploitability:                      # stop at every multiple of eval_interval
                n = min(n, eval_interval - done % eval_interval)
            self.solver.cfr_iterate(n)
            done += n
            if compute_exploitability and done % eval_interval == 0:
                history.append((done, self.solver.exploitability(0)))
        self._stale = True
        return history

    def exploitability(self):
        return self.solver.exploitability(0)

    def get_openspiel_policy(self):
        pol = LearnedCFRPolicy(self.game, self.info_set_map)
        pol._solver = self.solver                           # lets evaluate_agent play its episodes on the GPU
        return pol


class LearnedCFRPolicy(Policy):
    """Average policy of the visited infosets; uniform elsewhere (reference :122-144)."""

    def __init__(self, game, info_set_map):
        super().__init__(game, list(range(game.num_players())))
        self.info_set_map = info_set_map
        self._solver = None

    def action_probabilities(self, state):
        if state.is_terminal():
            return {}
        legal = state.legal_actions()
        node = self.info_set_map.get(state.information_state_string(state.current_player()))
        probs = node.policy if node is not None else _uniform(len(legal))
        return {a: probs[i] for i, a in enumerate(legal)}

    def _device_table(self, solver):
        return solver.policy_table_from_dict(self.info_set_map, lambda p, s: s, lambda node: node.policy)


class RandomPolicy(Policy):
    """Uniform over the legal actions."""

    def __init__(self, game):
        super().__init__(game, list(range(game.num_players())))

    def action_probabilities(self, state):
        legal = state.legal_actions()
        return {a: 1.0 / len(legal) for a in legal}

    def _device_table(self, solver):
        return solver.uniform_policy()
