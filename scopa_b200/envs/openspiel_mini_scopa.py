"""Drop-in for the reference's src/envs/openspiel_mini_scopa.py (MiniScopaState / MiniScopaGame,
registered as "mini_scopa"), backed by the CUDA env.  /root/reference/src/envs/openspiel_mini_scopa.py:5-186
"""
import numpy as np

from .. import _lib, codec
from .. import pyspiel_compat as pyspiel
from .mini_scopa_game import MiniScopaEnv, MiniScopaGame as _RulesGame, Player, _Discrete


class MiniScopaState(pyspiel.State):
    """OpenSpiel-compatible state wrapper around MiniScopaEnv."""

    def __init__(self, game, env=None, num_players=2, skip_reset=False):
        super().__init__(game)
        self.num_players = num_players
        self.env = env or MiniScopaEnv(num_players=num_players)
        if not skip_reset and env is not None:
            self.env.reset()
        # (a freshly constructed env has just been reset with the same seed: the reference's second
        #  reset() at :13 reproduces the identical deal, so it is skipped when we built the env here)
        self._is_terminal = False
        self.action_history = []

    # -- device queries (n = 1) ---------------------------------------------------------------------
    def _packed_np(self):
        words, order = self.env.packed()
        return np.array([words], dtype=np.uint32), np.array([order], dtype=np.uint32)

    def current_player(self):
        if self._is_terminal:
            return pyspiel.PlayerId.TERMINAL
        return self.env.agent_name_mapping[self.env.agent_selection]

    def legal_actions(self, player=None):
        """Returns legal actions based on cards in player's hand (hand order; [0] fallback)."""
        if self._is_terminal:
            return []
        if player is None:
            player = self.current_player()
        st, order = self._packed_np()
        ordered = np.zeros((1, 4), dtype=np.uint8)
        count = np.zeros(1, dtype=np.uint8)
        _lib.check(_lib.load().ms_legal_actions_host(st.ctypes.data, order.ctypes.data, int(player), None,
                                                     ordered.ctypes.data, count.ctypes.data, None, 1))
        return [int(a) for a in ordered[0, :count[0]]]

    def apply_action(self, action):
        """Applies action to environment and updates terminal flag."""
        self.action_history.append(action)
        self.env.step(action)
        self._is_terminal = all(self.env.terminations.values())

    def _apply_action(self, action):
        self.apply_action(action)

    def is_terminal(self):
        return self._is_terminal

    def is_chance_node(self):
        return False

    def chance_outcomes(self):
        return []

    def history_str(self):
        history_str = "-".join(map(str, self.action_history))
        if self._is_terminal:
            rewards_str = ",".join(f"{r:.2f}" for r in self.rewards())
            return f"TERMINAL:{history_str}:{rewards_str}"
        return f"H:{history_str}:P{self.current_player()}"

    def rewards(self):
        if not self._is_terminal:
            return [0] * self.num_players
        return [self.env.rewards[f"player_{i}"] for i in range(self.num_players)]

    def returns(self):
        return self.rewards()

    def information_state_string(self, player=None):
        if player is None:
            player = self.current_player()
        if self._is_terminal or player < 0:
            return "TERMINAL"
        st, order = self._packed_np()
        keys = np.zeros(1, dtype=np.uint64)
        _lib.check(_lib.load().ms_infoset_keys_host(st.ctypes.data, int(player), keys.ctypes.data, 1))
        return codec.key_to_string(keys[0], int(order[0]))

    def clone(self):
        """CFR-safe copy via state serialization (max_steps becomes 16 like the reference, :108)."""
        new_env = MiniScopaEnv.__new__(MiniScopaEnv)
        new_env.num_players = self.num_players
        new_env.game = _RulesGame.__new__(_RulesGame)      # no deck shuffle: set_state overwrites everything
        new_env.game.num_players = self.num_players
        new_env.game.deck = self.env.game.deck
        new_env.game.players = [Player(f"player_{i}") for i in range(self.num_players)]
        new_env.game.table = []
        new_env.game.last_capture = None
        new_env.possible_agents = [f"player_{i}" for i in range(self.num_players)]
        new_env.agent_name_mapping = {name: i for i, name in enumerate(new_env.possible_agents)}
        new_env._action_spaces = {a: _Discrete(16) for a in new_env.possible_agents}
        new_env.max_steps = 16
        new_env.seed = self.env.seed
        new_env.set_state(self.env.get_state())
        new_state = MiniScopaState(self.get_game(), env=new_env, num_players=self.num_players, skip_reset=True)
        new_state._is_terminal = self._is_terminal
        new_state.action_history = self.action_history.copy()
        return new_state


class MiniScopaGame(pyspiel.Game):
    """Game wrapper for OpenSpiel registration."""

    def __init__(self, num_players=2):
        self._num_players = num_players
        super().__init__(_game_type(num_players), pyspiel.GameInfo(
            num_distinct_actions=16, max_chance_outcomes=0, num_players=num_players, min_utility=-10.0,
            max_utility=10.0, utility_sum=0.0, max_game_length=num_players * 4), {})

    def num_players(self):
        return self._num_players

    def new_initial_state(self):
        return MiniScopaState(self, num_players=self._num_players)


def _game_type(num_players=2):
    return pyspiel.GameType(
        short_name="mini_scopa", long_name="Two-Player Mini-Scopa",
        dynamics=pyspiel.GameType.Dynamics.SEQUENTIAL, chance_mode=pyspiel.GameType.ChanceMode.DETERMINISTIC,
        information=pyspiel.GameType.Information.IMPERFECT_INFORMATION, utility=pyspiel.GameType.Utility.ZERO_SUM,
        reward_model=pyspiel.GameType.RewardModel.TERMINAL, max_num_players=num_players, min_num_players=num_players,
        provides_information_state_string=True, provides_information_state_tensor=False,
        provides_observation_string=False, provides_observation_tensor=False, parameter_specification={},
        default_loadable=True, provides_factored_observation_string=False)


def _mini_scopa_factory(params=None):
    return MiniScopaGame()


pyspiel.register_game(_game_type(), _mini_scopa_factory)
