"""TPI (team-public-information, two coordinators) OpenSpiel view of the CUDA team Miniscopa env, registered as
"team_mini_scopa_tpi".  Drop-in for /root/reference/src/envs/openspiel_team_mini_scopa.py:6-264: a state's
"player" is the TEAM whose member moves next; legal actions are that member's cards in hand order ([0] on an
empty hand); rewards are per team (mean of its two players' rewards); the info string is
"Team{t}:P{pid}:H[sorted hand]:T[sorted table]:A[action history]".  Transitions run on the GPU through
TeamMiniScopaEnv; what this module adds is formatting of the state the env already exposes."""
from .. import codec
from .. import pyspiel_compat as pyspiel
from . import team_mini_scopa_game as _team

_TEAM_MEMBERS = ((0, 1), (2, 3))


def _sorted_cards(cards):
    return "-".join(f"{r}{s[0]}" for r, s in sorted((c.rank, c.suit) for c in cards))


class TPIMiniScopaState(pyspiel.State):
    def __init__(self, game, env=None, skip_reset=False):
        super().__init__(game)
        fresh = env is None
        self.env = _team.TeamMiniScopaEnv() if fresh else env
        if not (skip_reset or fresh):
            self.env.reset()
        self._is_terminal = False
        self.action_history = []
        self._current_team = 0
        self._team_private_states = [None, None]

    def _mover(self):
        return self.env.agent_name_mapping[self.env.agent_selection]

    def current_player(self):
        return pyspiel.PlayerId.TERMINAL if self._is_terminal else self.env.game.get_team(self._mover())

    def _get_private_state_id(self, player_id):
        return tuple(sorted((c.rank, c.suit) for c in self.env.game.players[player_id].hand))

    def _get_team_private_states(self, team_id):
        return tuple(self._get_private_state_id(pid) for pid in _TEAM_MEMBERS[team_id])

    def legal_actions(self, player=None):
        if self._is_terminal:
            return []
        ids = [codec.card_id(c.rank, c.suit) for c in self.env.game.players[self._mover()].hand]
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

    def rewards(self):
        if not self._is_terminal:
            return [0, 0]
        r = [self.env.rewards[name] for name in self.env.possible_agents]
        return [(r[0] + r[1]) / 2, (r[2] + r[3]) / 2]

    def returns(self):
        return self.rewards()

    def history_str(self):
        played = "-".join(str(a) for a in self.action_history)
        if self._is_terminal:
            return "TERMINAL:%s:%s" % (played, ",".join("%.2f" % x for x in self.rewards()))
        return f"H:{played}:T{self.current_player()}"

    def information_state_string(self, player):
        members = _TEAM_MEMBERS[player]
        pid = self._mover()
        if pid not in members:
            pid = members[0]                            # not this team's turn: its first member's view
        g = self.env.game
        return "Team%d:P%d:H[%s]:T[%s]:A[%s]" % (player, pid, _sorted_cards(g.players[pid].hand), _sorted_cards(g.table),
                                                "-".join(str(a) for a in self.action_history))

    def clone(self):
        env = _team.TeamMiniScopaEnv.__new__(_team.TeamMiniScopaEnv)
        game = _team.TeamMiniScopaGame.__new__(_team.TeamMiniScopaGame)      # no shuffle: set_state overwrites all
        game.deck, game.table, game.last_capture_team = self.env.game.deck, [], None
        game.players = [_team.Player(f"player_{i}", team_id=i // 2) for i in range(4)]
        env.game = game
        env.possible_agents = [p.name for p in game.players]
        env.agent_name_mapping = {name: i for i, name in enumerate(env.possible_agents)}
        env._action_spaces = {name: _team._Discrete(16) for name in env.possible_agents}
        env.max_steps, env.seed = 16, self.env.seed
        env.set_state(self.env.get_state())
        twin = TPIMiniScopaState(self.get_game(), env=env, skip_reset=True)
        twin._is_terminal, twin._current_team = self._is_terminal, self._current_team
        twin.action_history = list(self.action_history)
        return twin


def _tpi_type():
    gt = pyspiel.GameType
    return gt(short_name="team_mini_scopa_tpi", long_name="Team Mini Scopa - TPI Representation",
              dynamics=gt.Dynamics.SEQUENTIAL, chance_mode=gt.ChanceMode.DETERMINISTIC,
              information=gt.Information.IMPERFECT_INFORMATION, utility=gt.Utility.ZERO_SUM,
              reward_model=gt.RewardModel.TERMINAL, max_num_players=2, min_num_players=2,
              provides_information_state_string=True, provides_information_state_tensor=False,
              provides_observation_string=False, provides_observation_tensor=False, parameter_specification={},
              default_loadable=True, provides_factored_observation_string=False)


class TPIMiniScopaGame(pyspiel.Game):
    """Two coordinators (teams) playing the 2v2 game, after Carminati et al. (ICML 2022)."""

    def __init__(self):
        super().__init__(_tpi_type(), pyspiel.GameInfo(num_distinct_actions=16, max_chance_outcomes=0, num_players=2,
                                                       min_utility=-20.0, max_utility=20.0, utility_sum=0.0,
                                                       max_game_length=16), {})

    def num_players(self):
        return 2

    def new_initial_state(self):
        return TPIMiniScopaState(self)


pyspiel.register_game(_tpi_type(), lambda params=None: TPIMiniScopaGame())
