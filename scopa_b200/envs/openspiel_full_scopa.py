"""OpenSpiel view of the CUDA 40-card Scopa env, registered as "full_scopa".  Drop-in for
/root/reference/src/envs/openspiel_full_scopa.py:4-185: legal actions are the mover's cards in hand order ([0] on an
empty hand), the info string is "P{p}:R{round}:H[sorted hand]:T[sorted table]:C[capture counts]:S[scopa counts]",
histories are the played action ids.  Transitions run on the GPU through FullScopaEnv; this module formats.

`clone()` works here.  In the reference it raises AttributeError (inside that module the name FullScopaGame is
rebound to the pyspiel.Game subclass before clone() uses it to build the env's game, :97-99 / :113), so nothing
there can be compared against; the clone continues from FullDeck()'s seed-42 order like set_state (:314-317)."""
from .. import full
from .. import pyspiel_compat as pyspiel
from . import full_scopa_game as _full


def _sorted_cards(cards):
    return "-".join(f"{r}{s[0]}" for r, s in sorted((c.rank, c.suit) for c in cards))


class FullScopaState(pyspiel.State):
    def __init__(self, game, env=None, num_players=2, skip_reset=False):
        super().__init__(game)
        self.num_players = num_players
        fresh = env is None
        self.env = _full.FullScopaEnv(num_players=num_players) if fresh else env
        if not (skip_reset or fresh):
            self.env.reset()
        self._is_terminal = False
        self.action_history = []

    def current_player(self):
        if self._is_terminal:
            return pyspiel.PlayerId.TERMINAL
        return self.env.agent_name_mapping[self.env.agent_selection]

    def legal_actions(self, player=None):
        if self._is_terminal:
            return []
        if player is None:
            player = self.current_player()
        ids = [full.card_id(c.rank, c.suit) for c in self.env.game.players[player].hand]
        return ids or [0]

    def apply_action(self, action):
        self.action_history.append(action)
        self.env.step(action)
        self._is_terminal = all(self.env.terminations.values())

    def is_terminal(self):
        return self._is_terminal

    def is_chance_node(self):
        return False

    def chance_outcomes(self):
        return []

    def history_str(self):
        played = "-".join(str(a) for a in self.action_history)
        if self._is_terminal:
            return "TERMINAL:%s:%s" % (played, ",".join("%.2f" % r for r in self.rewards()))
        return f"H:{played}:P{self.current_player()}"

    def rewards(self):
        if not self._is_terminal:
            return [0] * self.num_players
        return [self.env.rewards[f"player_{i}"] for i in range(self.num_players)]

    def returns(self):
        return self.rewards()

    def information_state_string(self, player):
        g = self.env.game
        return "P%d:R%d:H[%s]:T[%s]:C[%s]:S[%s]" % (
            player, g.round_number, _sorted_cards(g.players[player].hand), _sorted_cards(g.table),
            ",".join(str(len(p.captures)) for p in g.players), ",".join(str(p.scopas) for p in g.players))

    def clone(self):
        env = _full.FullScopaEnv.__new__(_full.FullScopaEnv)
        game = _full.FullScopaGame.__new__(_full.FullScopaGame)       # no shuffle here: set_state installs everything
        game.num_players, game.cards_per_hand = self.num_players, 3
        game.players = [_full.Player(f"player_{i}") for i in range(self.num_players)]
        game.table, game.last_capture, game.round_number, game.deck = [], None, 0, None
        env.num_players, env.game = self.num_players, game
        env.possible_agents = [f"player_{i}" for i in range(self.num_players)]
        env.agent_name_mapping = {name: i for i, name in enumerate(env.possible_agents)}
        env._action_spaces = {a: _full._Discrete(40) for a in env.possible_agents}
        env.max_steps, env.seed = 200, self.env.seed
        env.set_state(self.env.get_state())
        twin = FullScopaState(self.get_game(), env=env, num_players=self.num_players, skip_reset=True)
        twin._is_terminal = self._is_terminal
        twin.action_history = list(self.action_history)
        return twin


def _full_type(num_players=2):
    gt = pyspiel.GameType
    return gt(short_name="full_scopa", long_name="Full Italian Scopa", dynamics=gt.Dynamics.SEQUENTIAL,
              chance_mode=gt.ChanceMode.DETERMINISTIC, information=gt.Information.IMPERFECT_INFORMATION,
              utility=gt.Utility.ZERO_SUM, reward_model=gt.RewardModel.TERMINAL, max_num_players=num_players,
              min_num_players=num_players, provides_information_state_string=True,
              provides_information_state_tensor=False, provides_observation_string=False,
              provides_observation_tensor=False, parameter_specification={}, default_loadable=True,
              provides_factored_observation_string=False)


class FullScopaGame(pyspiel.Game):
    def __init__(self, num_players=2):
        self._num_players = num_players
        super().__init__(_full_type(num_players),
                         pyspiel.GameInfo(num_distinct_actions=40, max_chance_outcomes=0, num_players=num_players,
                                          min_utility=-10.0, max_utility=10.0, utility_sum=0.0, max_game_length=40), {})

    def num_players(self):
        return self._num_players

    def new_initial_state(self):
        return FullScopaState(self, num_players=self._num_players)


pyspiel.register_game(_full_type(), lambda params=None: FullScopaGame())
