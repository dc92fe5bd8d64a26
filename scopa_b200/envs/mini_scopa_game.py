"""Miniscopa rules and env under the reference's names, evaluated on the GPU.

Drop-in for /root/reference/src/envs/mini_scopa_game.py:6-194 -- `Card`, `MiniDeck`, `Player`,
`MiniScopaGame` (reset / card_in_table / play_card / evaluate_game) and the PettingZoo-style `MiniScopaEnv`
(reset / step / get_state / set_state with `agent_selection`, `rewards`, `terminations`, `truncations`,
`step_count`, `agents`, `possible_agents`, `agent_name_mapping`, `max_steps`, `seed`, `game`).

Division of labour: every rule evaluation -- the seeded shuffle, capture resolution, the transition, terminal
scoring -- is a kernel behind the C ABI (include/scopa_b200.h).  The Python lists only exist because the
reference exposes them (`env.game.players[i].hand / .captures / .scopas`, `env.game.table`); they are re-formatted
from the packed 16-byte device state after each call (scopa_b200/codec.py).  One kernel launch per call: this
scalar API is for compatibility, batches go through scopa_b200.batch / scopa_b200.solver.
"""
import numpy as np

from .. import _lib, codec

try:  # pragma: no cover - neither package is part of this image
    from gymnasium.spaces import Discrete as _Discrete
except ImportError:
    class _Discrete:
        def __init__(self, n):
            self.n = n

try:  # pragma: no cover
    from pettingzoo import AECEnv as _AECEnv
except ImportError:
    class _AECEnv:
        def __init__(self):
            pass

        def _was_dead_step(self, action):
            return None

_HAND = 4            # cards dealt to each player
_PASS = 255          # an action id no hand contains: the device treats it as the reference's silent pass


class Card:
    def __init__(self, rank: int, suit: str):
        self.rank, self.suit = rank, suit

    def __repr__(self):
        return f"{self.rank}_of_{self.suit}"


def _card(cid):
    return Card(codec.RANK_OF[cid], codec.SUIT_OF[cid])


def _cid(card):
    return codec.card_id(card.rank, card.suit)


def _ids(cards):
    return [_cid(c) for c in cards]


def _pairs(cards):
    return [(c.rank, c.suit) for c in cards]


def _i64(seed):
    seed = int(seed)
    if not -(2 ** 63) <= seed < 2 ** 63:
        raise OverflowError("scopa_b200 supports seeds in the signed 64-bit range")
    return seed


def _step_on_device(words, action, want_reward=False):
    """One ms_step on a single packed state -> (unpacked state, rewards[2] or None, done flag)."""
    st = np.asarray([words], dtype=np.uint32)
    act = np.asarray([action], dtype=np.uint8)
    rew = np.zeros((1, 2), dtype=np.float32)
    done = np.zeros(1, dtype=np.uint8)
    _lib.check(_lib.load().ms_step_host(st.ctypes.data, act.ctypes.data, rew.ctypes.data, done.ctypes.data, 1))
    return codec.unpack_state(st[0]), ([float(rew[0, 0]), float(rew[0, 1])] if want_reward else None), bool(done[0])


class MiniDeck:
    """The 16 cards (4 suits x 4 ranks, every rank on exactly two cards), shuffled by `seed` exactly like the
    reference's `random.seed(seed); random.shuffle(cards)` -- on the device (ms_deck_from_seeds)."""
    suits = list(codec.SUITS)
    ranks = {s: list(r) for s, r in codec.RANKS.items()}

    def __init__(self, seed=42):
        import torch
        s = torch.tensor([_i64(seed)], dtype=torch.int64, device="cuda")
        out = torch.empty(1, dtype=torch.int64, device="cuda")
        _lib.check(_lib.load().ms_deck_from_seeds(s.data_ptr(), 1, out.data_ptr(), _lib.stream_ptr()))
        perm = int(out.cpu().numpy().view(np.uint64)[0])
        self.cards = [_card(c) for c in codec.nibbles(perm, 16)]

    def deal(self, n):
        head, self.cards = self.cards[:n], self.cards[n:]
        return head


class Player:
    def __init__(self, name):
        self.name, self.hand, self.captures, self.scopas = name, [], [], 0

    def reset(self):
        self.hand.clear()
        self.captures.clear()
        self.scopas = 0


class MiniScopaGame:
    def __init__(self, num_players=2):
        if num_players != 2:
            raise ValueError("the CUDA Miniscopa path implements the 2-player game the reference's solvers use")
        self.num_players = num_players
        self.deck = MiniDeck()
        self.players = [Player(f"player_{i}") for i in range(num_players)]
        self.table, self.last_capture = [], None

    def reset(self, seed=42):
        self.deck = MiniDeck(seed)
        self.table.clear()
        self.last_capture = None
        for p in self.players:
            p.reset()
            p.hand = self.deck.deal(_HAND)

    # ---- packed-state plumbing ----------------------------------------------------------------------------
    def _pack(self, cur, step_count=0, terminal=False, max_steps=31):
        hands = [_ids(p.hand) for p in self.players]
        table = _ids(self.table)
        if len(table) > 8 or any(len(h) > _HAND for h in hands):
            raise ValueError("Miniscopa state out of range (table > 8 cards or hand > 4 cards)")
        words = codec.pack_state([codec.mask_of(h) for h in hands], table,
                                 [codec.mask_of(_ids(p.captures)) for p in self.players],
                                 [p.scopas for p in self.players], step_count, cur, terminal, max_steps)
        order = []
        for h in hands:                                  # hand order = list order, padded with cards not in the hand
            order += (h + [c for c in range(16) if c not in h])[:_HAND]
        return words, codec.pack_nibbles(order)

    def _absorb(self, unpacked, table_before, played, mover):
        """Bring the lists in line with the device state after one transition of `mover`."""
        pl = self.players[mover]
        if played is not None:
            card = next(c for c in pl.hand if _cid(c) == played)
            gone = [c for c in table_before if c not in unpacked["table"]]
            if gone:                                     # capture: captured cards in table order, then the card
                by_id = {_cid(c): c for c in self.table}
                pl.captures.extend([by_id[c] for c in gone] + [card])
                self.table = [c for c in self.table if _cid(c) not in gone]
                self.last_capture = pl
            else:
                self.table.append(card)
            pl.hand.remove(card)
        for p, n in zip(self.players, unpacked["scopas"]):
            p.scopas = n

    # ---- the reference's rule API -------------------------------------------------------------------------
    def card_in_table(self, card):
        """(captures?, [captured table cards]) for playing `card` on the current table (reference :66-91)."""
        import torch
        if card.rank <= 0 or not self.table:
            return False, []
        cid = _cid(card)
        if cid < 0:
            raise ValueError(f"{card!r} is not a Miniscopa card")
        words, _ = self._pack(0)
        st = torch.tensor(np.asarray([words], dtype=np.uint32).view(np.int32), device="cuda")
        cd = torch.tensor([cid], dtype=torch.uint8, device="cuda")
        out = torch.empty(1, dtype=torch.uint8, device="cuda")
        _lib.check(_lib.load().ms_capture(st.data_ptr(), cd.data_ptr(), out.data_ptr(), 1, _lib.stream_ptr()))
        mask = int(out.item())
        return bool(mask), [c for i, c in enumerate(self.table) if (mask >> i) & 1]

    def play_card(self, card, player):
        if card not in player.hand:
            raise ValueError("list.remove(x): x not in list")
        mover = self.players.index(player)
        before = _ids(self.table)
        u, _, _ = _step_on_device(self._pack(mover)[0], _cid(card))
        self._absorb(u, before, _cid(card), mover)

    def evaluate_game(self):
        """Zero-sum terminal rewards (+1 per captured card, +2 per scopa, minus the mean), scored on the device:
        a pass at step 30 of 31 ends the game and makes the kernel emit the rewards."""
        _, rew, _ = _step_on_device(self._pack(0, step_count=30, max_steps=31)[0], _PASS, want_reward=True)
        if not any(len(p.captures) + 2 * p.scopas for p in self.players):
            return [0] * self.num_players
        return rew


class MiniScopaEnv(_AECEnv):
    metadata = {"name": "Mini-Scopa-v0"}

    def __init__(self, seed=42, num_players=2):
        super().__init__()
        self.num_players = num_players
        self.game = MiniScopaGame(num_players=num_players)
        self.possible_agents = [p.name for p in self.game.players]
        self.agent_name_mapping = {name: i for i, name in enumerate(self.possible_agents)}
        self._action_spaces = {name: _Discrete(16) for name in self.possible_agents}
        self.max_steps = _HAND * num_players
        self.seed = seed
        self.reset(seed)

    def _fresh_flags(self):
        self.agents = list(self.possible_agents)
        self.agent_selection = self.agents[0]
        self.rewards = dict.fromkeys(self.agents, 0)
        self.terminations = dict.fromkeys(self.agents, False)
        self.truncations = dict.fromkeys(self.agents, False)
        self.step_count = 0

    def reset(self, seed=None):
        self.game.reset(seed or self.seed)               # reset(0) and reset(None) both mean self.seed (:132)
        self._fresh_flags()

    def step(self, action):
        if self.terminations[self.agent_selection]:
            self._was_dead_step(action)
            return
        mover = self.agent_name_mapping[self.agent_selection]
        g = self.game
        a = int(action)
        in_hand = 0 <= a < 16 and a in _ids(g.players[mover].hand)
        before = _ids(g.table)
        words, _ = g._pack(mover, self.step_count, False, min(self.max_steps, 31))
        u, rew, done = _step_on_device(words, a if 0 <= a < 16 else _PASS, want_reward=True)
        g._absorb(u, before, a if in_hand else None, mover)          # not in hand: silent pass (:155-157)
        self.step_count = u["step_count"]
        if done:
            scored = any(len(p.captures) + 2 * p.scopas for p in g.players)
            for i, name in enumerate(self.agents):
                self.rewards[name] = rew[i] if scored else 0
                self.terminations[name] = True
        self.agent_selection = self.agents[(self.agents.index(self.agent_selection) + 1) % self.num_players]

    def get_state(self):
        g = self.game
        snap = {"table": _pairs(g.table), "hands": [_pairs(p.hand) for p in g.players],
                "captures": [_pairs(p.captures) for p in g.players], "scopas": [p.scopas for p in g.players],
                "agent_selection": self.agent_selection, "step_count": self.step_count, "agents": list(self.agents)}
        for k in ("rewards", "terminations", "truncations"):
            snap[k] = dict(getattr(self, k))
        return snap

    def set_state(self, state):
        g = self.game
        g.table = [Card(*rs) for rs in state["table"]]
        for p, hand, caps, n in zip(g.players, state["hands"], state["captures"], state["scopas"]):
            p.hand, p.captures, p.scopas = [Card(*rs) for rs in hand], [Card(*rs) for rs in caps], n
        self.agent_selection, self.step_count = state["agent_selection"], state["step_count"]
        self.agents = list(state["agents"])
        for k in ("rewards", "terminations", "truncations"):
            setattr(self, k, dict(state[k]))

    def packed(self):
        """(packed state words, hand order) of the current position -- what the solver classes root on."""
        mover = self.agent_name_mapping[self.agent_selection]
        return self.game._pack(mover, self.step_count, all(self.terminations.values()), min(self.max_steps, 31))
