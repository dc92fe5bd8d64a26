"""Shim of the pyspiel names the reference touches (registry + base classes only)."""
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


class State:
    def __init__(self, game):
        self._game = game

    def get_game(self):
        return self._game

    def child(self, action):
        c = self.clone()
        c.apply_action(action)
        return c


_REGISTRY = {}


def register_game(game_type, factory):
    _REGISTRY[game_type.short_name] = factory


def load_game(name, params=None):
    return _REGISTRY[name](params)
