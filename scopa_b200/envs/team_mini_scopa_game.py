"""2v2 team Miniscopa under the reference's names (drop-in for /root/reference/src/envs/team_mini_scopa_game.py):
`Card`, `MiniDeck`, `Player(name, team_id)`, `TeamMiniScopaGame`, `TeamMiniScopaEnv`.  Rules run on the GPU
(csrc/ms_team.cu); the Python lists mirror the packed 32-byte device state after each call."""
import numpy as np

from .. import _lib, codec
from ..team import pack_team_state, unpack_team_state
from .mini_scopa_game import MiniDeck, _AECEnv, _Discrete, _PASS


class Card:
    """Value-comparable card, as in the reference's team module (:5-17)."""

    def __init__(self, rank: int, suit: str):
        self.rank, self.suit = rank, suit

    def __repr__(self):
        return f"{self.rank}_{self.suit}"

    def __eq__(self, other):
        return (self.rank, self.suit) == (other.rank, other.suit)

    def __hash__(self):
        return hash((self.rank, self.suit))


def _cid(card):
    return codec.card_id(card.rank, card.suit)


def _ids(cards):
    return [_cid(c) for c in cards]


class Player:
    def __init__(self, name, team_id):
        self.name, self.team_id = name, team_id
        self.hand, self.captures, self.scopas = [], [], 0

    def reset(self):
        self.hand.clear()
        self.captures.clear()
        self.scopas = 0


class TeamMiniScopaGame:
    def __init__(self):
        self.deck = MiniDeck()
        self.players = [Player(f"player_{i}", team_id=i // 2) for i in range(4)]
        self.table, self.last_capture_team = [], None

    def reset(self, seed=42):
        self.deck = MiniDeck(seed)
        self.table.clear()
        self.last_capture_team = None
        for p in self.players:
            p.reset()
            p.hand = [Card(c.rank, c.suit) for c in self.deck.deal(4)]

    def get_team(self, player_id):
        return self.players[player_id].team_id

    def _pack(self, cur, step_count=0, terminal=False, max_steps=16):
        return pack_team_state([codec.mask_of(_ids(p.hand)) for p in self.players], _ids(self.table),
                               [codec.mask_of(_ids(p.captures)) for p in self.players],
                               [p.scopas for p in self.players], step_count, cur, terminal, self.last_capture_team, max_steps)

    def _step(self, words, action):
        st = np.asarray([words], dtype=np.uint32)
        act = np.asarray([action], dtype=np.uint8)
        rew, done = np.zeros((1, 4), dtype=np.float32), np.zeros(1, dtype=np.uint8)
        _lib.check(_lib.load().ms_team_step_host(st.ctypes.data, act.ctypes.data, rew.ctypes.data, done.ctypes.data, 1))
        return unpack_team_state(st[0]), [float(x) for x in rew[0]], bool(done[0])

    def _absorb(self, u, table_before, played, mover, finished):
        pl = self.players[mover]
        if played is not None:
            card = next(c for c in pl.hand if _cid(c) == played)
            gone = [c for c in table_before if c not in u["table"]]
            if gone:
                by_id = {_cid(c): c for c in self.table}
                pl.captures.extend([by_id[c] for c in gone] + [card])
                self.table = [c for c in self.table if _cid(c) not in gone]
            else:
                self.table.append(card)
            pl.hand.remove(card)
        self.last_capture_team = u["last_capture_team"]
        for p, n in zip(self.players, u["scopas"]):
            p.scopas = n
        if finished and self.table and self.last_capture_team is not None:
            self.players[2 * self.last_capture_team].captures.extend(self.table)   # the sweep (:126-132); table stays

    def play_card(self, card, player):
        if card not in player.hand:
            raise ValueError("list.remove(x): x not in list")
        mover = self.players.index(player)
        before = _ids(self.table)
        u, _, _ = self._step(self._pack(mover, 0, False, 31), _cid(card))
        self._absorb(u, before, _cid(card), mover, False)


class TeamMiniScopaEnv(_AECEnv):
    metadata = {"name": "Team-Mini-Scopa-v0"}

    def __init__(self, seed=42):
        super().__init__()
        self.game = TeamMiniScopaGame()
        self.possible_agents = [p.name for p in self.game.players]
        self.agent_name_mapping = {name: i for i, name in enumerate(self.possible_agents)}
        self._action_spaces = {name: _Discrete(16) for name in self.possible_agents}
        self.max_steps = 16
        self.seed = seed
        self.reset(seed)

    def reset(self, seed=None):
        self.game.reset(seed or self.seed)
        self.agents = list(self.possible_agents)
        self.agent_selection = self.agents[0]
        self.rewards = dict.fromkeys(self.agents, 0)
        self.terminations = dict.fromkeys(self.agents, False)
        self.truncations = dict.fromkeys(self.agents, False)
        self.step_count = 0

    def step(self, action):
        if self.terminations[self.agent_selection]:
            self._was_dead_step(action)
            return
        mover = self.agent_name_mapping[self.agent_selection]
        g = self.game
        a = int(action)
        in_hand = 0 <= a < 16 and a in _ids(g.players[mover].hand)
        before = _ids(g.table)
        u, rew, done = g._step(g._pack(mover, self.step_count, False, min(self.max_steps, 31)), a if 0 <= a < 16 else _PASS)
        g._absorb(u, before, a if in_hand else None, mover, done)
        self.step_count = u["step_count"]
        if done:
            scored = any(len(p.captures) + 2 * p.scopas for p in g.players)
            for i, name in enumerate(self.agents):
                self.rewards[name] = rew[i] if scored else 0
                self.terminations[name] = True
        self.agent_selection = self.agents[(mover + 1) % 4]

    def get_state(self):
        g = self.game
        pairs = lambda cards: [(c.rank, c.suit) for c in cards]
        snap = {"table": pairs(g.table), "hands": [pairs(p.hand) for p in g.players],
                "captures": [pairs(p.captures) for p in g.players], "scopas": [p.scopas for p in g.players],
                "last_capture_team": g.last_capture_team, "agent_selection": self.agent_selection,
                "step_count": self.step_count, "agents": list(self.agents)}
        for k in ("rewards", "terminations", "truncations"):
            snap[k] = dict(getattr(self, k))
        return snap

    def set_state(self, state):
        g = self.game
        g.table = [Card(*rs) for rs in state["table"]]
        for p, hand, caps, n in zip(g.players, state["hands"], state["captures"], state["scopas"]):
            p.hand, p.captures, p.scopas = [Card(*rs) for rs in hand], [Card(*rs) for rs in caps], n
        g.last_capture_team = state["last_capture_team"]
        self.agent_selection, self.step_count = state["agent_selection"], state["step_count"]
        self.agents = list(state["agents"])
        for k in ("rewards", "terminations", "truncations"):
            setattr(self, k, dict(state[k]))
