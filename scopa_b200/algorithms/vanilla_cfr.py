"""Drop-in for the reference's src/algorithms/vanilla_cfr.py; the traversal runs in the CUDA solver.

CFRTrainer keeps the reference's surface (/root/reference/src/algorithms/vanilla_cfr.py:41-120):
`train(steps, eval_interval, compute_exploitability)`, `_cfr_recursive(state, player, r0, r1)`,
`info_set_map` (dict info-string -> InfoNode with regret_sum / strategy_sum / local_strategy),
`get_openspiel_policy()`.  The table lives in HBM; `info_set_map` is a host view refreshed lazily.
"""
from dataclasses import dataclass

import numpy as np

from ..solver import Solver
from ._policy_base import Policy, root_of


@dataclass
class InfoNode:
    legal_actions: np.ndarray
    regret_sum: np.ndarray = None
    strategy_sum: np.ndarray = None
    local_strategy: np.ndarray = None

    def __post_init__(self):
        if self.regret_sum is None:
            self.regret_sum = np.zeros(self.legal_actions.size)
        if self.strategy_sum is None:
            self.strategy_sum = np.zeros(self.legal_actions.size)
        if self.local_strategy is None:
            self.local_strategy = np.ones(self.legal_actions.size) / self.legal_actions.size

    def get_strategy(self):
        positive_regrets = np.maximum(self.regret_sum, 0)
        norm_sum = np.sum(positive_regrets)
        if norm_sum > 0:
            return positive_regrets / norm_sum
        return np.ones(self.legal_actions.size) / self.legal_actions.size

    @property
    def policy(self) -> np.ndarray:
        norm_sum = np.sum(self.strategy_sum)
        if norm_sum > 0:
            return self.strategy_sum / norm_sum
        return np.ones(self.legal_actions.size) / self.legal_actions.size


class CFRTrainer:
    def __init__(self, game, device="cuda"):
        self.game = game
        words, order = root_of(game)
        self._root = (tuple(int(w) for w in words), int(order))
        self.solver = Solver(words, order, device=device)
        self._map = {}
        self._dirty = False
        self._visited = False

    # -- host view of the device table ----------------------------------------------------------------
    @property
    def info_set_map(self):
        if self._dirty:
            self._refresh()
        return self._map

    def _refresh(self):
        st = self.solver.static_table()
        reg, strat, _ = self.solver.export()
        m = {}
        for s in st["dfs_order"]:                      # the reference's dict insertion order
            n = int(st["nlegal"][s])
            node = InfoNode(np.array([int(a) for a in st["legal"][s, :n]]), reg[s, :n].copy(), strat[s, :n].copy())
            node.local_strategy = node.get_strategy()  # invariant after every visit (vanilla_cfr.py:97)
            m[st["strings"][s]] = node
        self._map = m
        self._dirty = False

    def _get_or_create_node(self, info_set_key, legal_actions) -> InfoNode:
        m = self.info_set_map
        if info_set_key not in m:
            m[info_set_key] = InfoNode(np.array(legal_actions))
        return m[info_set_key]

    def _cfr_recursive(self, state, traversing_player, reach_p0, reach_p1):
        """One reference traversal.  Only root states are supported (that is how the reference and its
        experiment runner call it: run_vanilla_cfr_experiment.py:87-91)."""
        if state.is_terminal():
            return state.rewards()[traversing_player]
        words, order = state.env.packed()
        if (tuple(int(w) for w in words), int(order)) != self._root:
            raise NotImplementedError("_cfr_recursive on a non-root state: build a CFRTrainer for a game rooted there")
        v = self.solver.cfr_traverse(traversing_player, reach_p0, reach_p1)
        self._dirty = True
        return v

    def get_openspiel_policy(self):
        pol = LearnedCFRPolicy(self.game, self.info_set_map)
        pol._solver = self.solver          # lets evaluate_agent play the episodes on the GPU
        return pol

    def exploitability(self):
        """Best-response exploitability of the current average policy (device sweep)."""
        return self.solver.exploitability(0)

    def train(self, steps: int, eval_interval: int = 1000, compute_exploitability: bool = False):
        exploitability_history = []
        t = 0
        while t < steps:
            chunk = steps - t
            if compute_exploitability:
                chunk = min(chunk, eval_interval - (t % eval_interval))
            self.solver.cfr_iterate(chunk)
            t += chunk
            if compute_exploitability and t % eval_interval == 0:
                exploitability_history.append((t, self.solver.exploitability(0)))
        self._dirty = True
        return exploitability_history


class LearnedCFRPolicy(Policy):
    def __init__(self, game, info_set_map):
        super().__init__(game, list(range(game.num_players())))
        self.info_set_map = info_set_map
        self._solver = None

    def _device_table(self, solver):
        return solver.policy_table_from_dict(self.info_set_map, lambda p, s: s, lambda node: node.policy)

    def action_probabilities(self, state):
        if state.is_terminal():
            return {}
        player = state.current_player()
        info_state = state.information_state_string(player)
        legal_actions = state.legal_actions()
        if info_state in self.info_set_map:
            probs = self.info_set_map[info_state].policy
            return {action: probs[i] for i, action in enumerate(legal_actions)}
        prob = 1.0 / len(legal_actions)
        return {action: prob for action in legal_actions}


class RandomPolicy(Policy):
    """Policy that chooses actions uniformly at random."""

    def __init__(self, game):
        super().__init__(game, list(range(game.num_players())))

    def _device_table(self, solver):
        return solver.uniform_policy()

    def action_probabilities(self, state):
        legal_actions = state.legal_actions()
        prob = 1.0 / len(legal_actions)
        return {action: prob for action in legal_actions}


def _evaluate_agent_device(solver, trained_policy, opponent_policy, num_episodes, seed):
    """All episodes in two launches (agent in seat 0 for the first half, seat 1 for the second)."""
    t_tab, o_tab = trained_policy._device_table(solver), opponent_policy._device_table(solver)
    n0 = int(np.ceil(num_episodes / 2))                      # episodes with `episode < num_episodes / 2`
    n1 = num_episodes - n0
    r_a, s_a = solver.evaluate(t_tab, o_tab, n0, philox_seed=seed, first_game=0)
    r_b, s_b = solver.evaluate(o_tab, t_tab, n1, philox_seed=seed, first_game=n0)
    rew = np.concatenate([r_a.cpu().numpy().astype(np.float64), -r_b.cpu().numpy().astype(np.float64)])
    s_a, s_b = s_a.cpu().numpy().astype(np.int64), s_b.cpu().numpy().astype(np.int64)
    trained = np.concatenate([s_a[:, 0], s_b[:, 1]])
    opp = np.concatenate([s_a[:, 1], s_b[:, 0]])
    k = np.arange(1, num_episodes + 1)
    ct, co = np.cumsum(trained), np.cumsum(opp)
    avg_reward_history = (np.cumsum(rew) / k).tolist()
    scopa_history = {'trained': (ct / k).tolist(), 'opponent': (co / k).tolist(), 'diff': ((ct - co) / k).tolist()}
    avg_t, avg_o = trained.sum() / num_episodes, opp.sum() / num_episodes
    scopa_stats = {'trained_avg': avg_t, 'opponent_avg': avg_o, 'difference': avg_t - avg_o, 'history': scopa_history,
                   'data_collected': num_episodes > 0}
    return rew.sum() / num_episodes, avg_reward_history, scopa_stats


def evaluate_agent(game, trained_policy, opponent_policy, num_episodes=10000, seed=None):
    """Reference vanilla_cfr.py:157-216: episodes vs an opponent, seats swapped at half time.

    When both policies come from this package (tabular policies of a CUDA trainer, RandomPolicy) the episodes
    are played on the GPU in two launches (Philox stream `seed`, default drawn from np.random so that
    np.random.seed() still makes runs repeatable); any other policy object takes the reference's scalar loop."""
    solver = getattr(trained_policy, "_solver", None) or getattr(opponent_policy, "_solver", None)
    if (solver is not None and num_episodes > 0 and hasattr(trained_policy, "_device_table")
            and hasattr(opponent_policy, "_device_table")):
        if seed is None:
            seed = int(np.random.randint(0, 2**31 - 1))
        return _evaluate_agent_device(solver, trained_policy, opponent_policy, num_episodes, seed)
    total_winnings = 0
    avg_reward_history = []
    trained_scopas = 0
    opponent_scopas = 0
    scopa_history = {'trained': [], 'opponent': [], 'diff': []}
    for episode in range(num_episodes):
        if episode < num_episodes / 2:
            agent_seat = 0
            policies = [trained_policy, opponent_policy]
        else:
            agent_seat = 1
            policies = [opponent_policy, trained_policy]
        state = game.new_initial_state()
        while not state.is_terminal():
            player = state.current_player()
            action_probs = policies[player].action_probabilities(state)
            actions, probs = zip(*action_probs.items())
            action = np.random.choice(actions, p=probs)
            state.apply_action(action)
        total_winnings += state.rewards()[agent_seat]
        avg_reward_history.append(total_winnings / (episode + 1))
        players = state.env.game.players
        trained_scopas += players[agent_seat].scopas
        opponent_scopas += players[1 - agent_seat].scopas
        scopa_history['trained'].append(trained_scopas / (episode + 1))
        scopa_history['opponent'].append(opponent_scopas / (episode + 1))
        scopa_history['diff'].append((trained_scopas - opponent_scopas) / (episode + 1))
    avg_reward = total_winnings / num_episodes
    avg_trained_scopas = trained_scopas / num_episodes
    avg_opponent_scopas = opponent_scopas / num_episodes
    scopa_stats = {
        'trained_avg': avg_trained_scopas, 'opponent_avg': avg_opponent_scopas,
        'difference': avg_trained_scopas - avg_opponent_scopas, 'history': scopa_history,
        'data_collected': len(scopa_history['trained']) > 0,
    }
    return avg_reward, avg_reward_history, scopa_stats
