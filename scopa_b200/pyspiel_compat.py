"""The few pyspiel names the reference touches (registry, base classes, enums).

If the real `pyspiel` (open_spiel) is installed it is used as is and "mini_scopa" is registered
with it; otherwise this module provides the same names with the same call shapes, so that
`pyspiel.load_game("mini_scopa")` works either way:

    from scopa_b200 import pyspiel_compat as pyspiel
"""
try:  # pragma: no cover - open_spiel is not part of this image
    import pyspiel as _real
    HAVE_OPEN_SPIEL = True
except ImportError:
    _real = None
    HAVE_OPEN_SPIEL = False

if HAVE_OPEN_SPIEL:  # pragma: no cover
    PlayerId, GameType, GameInfo, Game, State = _real.PlayerId, _real.GameType, _real.GameInfo, _real.Game, _real.State
    register_game, load_game = _real.register_game, _real.load_game
else:
    import enum

    class PlayerId:
        TERMINAL = -4
        CHANCE = -1
        INVALID = -3

    class _Bag:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class GameType(_Bag):
        class Dynamics(enum.Enum):
            SEQUENTIAL = 0
            SIMULTANEOUS = 1

        class ChanceMode(enum.Enum):
            DETERMINISTIC = 0
            EXPLICIT_STOCHASTIC = 1
            SAMPLED_STOCHASTIC = 2

        class Information(enum.Enum):
            ONE_SHOT = 0
            PERFECT_INFORMATION = 1
            IMPERFECT_INFORMATION = 2

        class Utility(enum.Enum):
            ZERO_SUM = 0
            CONSTANT_SUM = 1
            GENERAL_SUM = 2
            IDENTICAL = 3

        class RewardModel(enum.Enum):
            REWARDS = 0
            TERMINAL = 1

    class GameInfo(_Bag):
        pass

    class Game:
        def __init__(self, game_type, game_info, params):
            self._type, self._info, self._params = game_type, game_info, params

        def get_type(self):
            return self._type

        def utility_sum(self):
            return getattr(self._info, "utility_sum", 0.0)

    class State:
        def __init__(self, game):
            self._game = game

        def get_game(self):
            return self._game

        def child(self, action):
            c = self.clone()
            c.apply_action(action)
            return c

        def player_return(self, player):
            return self.returns()[player]

    _REGISTRY = {}

    def register_game(game_type, factory):
        _REGISTRY[game_type.short_name] = factory

    def load_game(name, params=None):
        return _REGISTRY[name](params)
