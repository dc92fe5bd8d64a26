"""OpenSpiel-protocol view of the CUDA Miniscopa env: `MiniScopaState`, `MiniScopaGame`, registered as
"mini_scopa".  Drop-in for the reference module of the same name
(/root/reference/src/envs/openspiel_mini_scopa.py:5-186): same class names, same method set
(current_player, legal_actions, apply_action, is_terminal, is_chance_node, chance_outcomes, history_str,
rewards, returns, information_state_string, clone) and the same observable quirks:

  * legal actions are listed in HAND (= deal) order and fall back to [0] on an empty non-terminal hand (:22-47);
  * a clone's env has max_steps = 16 instead of 8 (:108);
  * the info string is "P{p}:H[..]_T[..]" with cards as rank + suit initial, "TERMINAL" once the game is over
    (:86-95).

Every query is answered by a kernel through the C ABI (one launch per call: this scalar API exists for
compatibility, throughput work goes through scopa_b200.batch / scopa_b200.solver).
"""
import numpy as np

from .. import _lib, codec
from .. import pyspiel_compat as pyspiel
from . import mini_scopa_game as _msg

_TERMINAL_PLAYER = pyspiel.PlayerId.TERMINAL


def _np_state(env):
    """(state words [1,4] u32, hand order [1] u32) of an env, as host arrays for the *_host entry points."""
    words, order = env.packed()
    return np.asarray([words], dtype=np.uint32), np.asarray([order], dtype=np.uint32)


def _clone_env(src, num_players):
    """A detached MiniScopaEnv holding src's state; no deck is shuffled (the reference builds a fresh deck per
    clone, 32 % of its CFR run time) and max_steps becomes 16 as in the reference."""
    env = _msg.MiniScopaEnv.__new__(_msg.MiniScopaEnv)
    game = _msg.MiniScopaGame.__new__(_msg.MiniScopaGame)
    game.num_players, game.deck, game.table, game.last_capture = num_players, src.game.deck, [], None
    game.players = [_msg.Player(f"player_{i}") for i in range(num_players)]
    env.num_players, env.game = num_players, game
    env.possible_agents = [p.name for p in game.players]
    env.agent_name_mapping = {name: i for i, name in enumerate(env.possible_agents)}
    env._action_spaces = {name: _msg._Discrete(16) for name in env.possible_agents}
    env.max_steps, env.seed = 16, src.seed
    env.set_state(src.get_state())
    return env


class MiniScopaState(pyspiel.State):
    def __init__(self, game, env=None, num_players=2, skip_reset=False):
        super().__init__(game)
        self.num_players = num_players
        fresh = env is None
        self.env = _msg.MiniScopaEnv(num_players=num_players) if fresh else env
        if not (skip_reset or fresh):
            self.env.reset()        # a freshly built env has just dealt this very deal: no second shuffle needed
        self._is_terminal = False
        self.action_history = []

    # ---- protocol ---------------------------------------------------------------------------------------
    def current_player(self):
        return _TERMINAL_PLAYER if self._is_terminal else self.env.agent_name_mapping[self.env.agent_selection]

    def is_terminal(self):
        return self._is_terminal

    def is_chance_node(self):
        return False                                    # the deal is fixed before the first state exists

    def chance_outcomes(self):
        return []

    def legal_actions(self, player=None):
        if self._is_terminal:
            return []
        who = self.current_player() if player is None else player
        st, order = _np_state(self.env)
        ids, cnt = np.zeros((1, 4), dtype=np.uint8), np.zeros(1, dtype=np.uint8)
        _lib.check(_lib.load().ms_legal_actions_host(st.ctypes.data, order.ctypes.data, int(who), None,
                                                     ids.ctypes.data, cnt.ctypes.data, None, 1))
        return ids[0, :int(cnt[0])].astype(int).tolist()

    def apply_action(self, action):
        self.action_history.append(action)
        self.env.step(action)
        self._is_terminal = all(self.env.terminations.values())

    _apply_action = apply_action                        # the name OpenSpiel's own algorithms call

    def rewards(self):
        if self._is_terminal:
            return [self.env.rewards[name] for name in self.env.possible_agents[:self.num_players]]
        return [0] * self.num_players

    def returns(self):
        return self.rewards()

    def information_state_string(self, player=None):
        who = self.current_player() if player is None else player
        if self._is_terminal or who < 0:
            return "TERMINAL"
        st, order = _np_state(self.env)
        key = np.zeros(1, dtype=np.uint64)
        _lib.check(_lib.load().ms_infoset_keys_host(st.ctypes.data, int(who), key.ctypes.data, 1))
        return codec.key_to_string(key[0], int(order[0]))

    def history_str(self):
        played = "-".join(str(a) for a in self.action_history)
        if not self._is_terminal:
            return f"H:{played}:P{self.current_player()}"
        return "TERMINAL:%s:%s" % (played, ",".join("%.2f" % r for r in self.rewards()))

    def clone(self):
        twin = MiniScopaState(self.get_game(), env=_clone_env(self.env, self.num_players),
                              num_players=self.num_players, skip_reset=True)
        twin._is_terminal = self._is_terminal
        twin.action_history = list(self.action_history)
        return twin


def _game_type(num_players=2):
    gt = pyspiel.GameType
    return gt(short_name="mini_scopa", long_name="Two-Player Mini-Scopa", dynamics=gt.Dynamics.SEQUENTIAL,
              chance_mode=gt.ChanceMode.DETERMINISTIC, information=gt.Information.IMPERFECT_INFORMATION,
              utility=gt.Utility.ZERO_SUM, reward_model=gt.RewardModel.TERMINAL, max_num_players=num_players,
              min_num_players=num_players, provides_information_state_string=True,
              provides_information_state_tensor=False, provides_observation_string=False,
              provides_observation_tensor=False, parameter_specification={}, default_loadable=True,
              provides_factored_observation_string=False)


class MiniScopaGame(pyspiel.Game):
    def __init__(self, num_players=2):
        self._num_players = num_players
        info = pyspiel.GameInfo(num_distinct_actions=16, max_chance_outcomes=0, num_players=num_players,
                                min_utility=-10.0, max_utility=10.0, utility_sum=0.0, max_game_length=4 * num_players)
        super().__init__(_game_type(num_players), info, {})

    def num_players(self):
        return self._num_players

    def new_initial_state(self):
        return MiniScopaState(self, num_players=self._num_players)


def _mini_scopa_factory(params=None):
    return MiniScopaGame()


pyspiel.register_game(_game_type(), _mini_scopa_factory)
