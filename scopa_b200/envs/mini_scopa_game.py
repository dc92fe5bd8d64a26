"""Drop-in for the reference's src/envs/mini_scopa_game.py, backed by the CUDA kernels.

Same classes, attributes and behaviour as the reference (Card, MiniDeck, Player, MiniScopaGame,
MiniScopaEnv -- /root/reference/src/envs/mini_scopa_game.py:6-194), but every rule evaluation (the
shuffle, capture resolution, the transition, terminal scoring) runs on the GPU through the C ABI
(include/scopa_b200.h).  The Python lists exist because the reference exposes them
(`env.game.players[i].hand`, `.captures`, `.scopas`, `env.game.table`); they are re-formatted from
the packed device state after each call.  This scalar API costs one kernel launch per call -- it is
for compatibility; throughput work uses scopa_b200.batch / scopa_b200.solver.
"""
import ctypes as C

import numpy as np

from .. import _lib, codec


class _Discrete:
    """gymnasium.spaces.Discrete stand-in (the reference only stores it)."""

    def __init__(self, n):
        self.n = n


try:  # pragma: no cover - not in this image
    from gymnasium.spaces import Discrete as _Discrete  # noqa: F811
except ImportError:
    pass

try:  # pragma: no cover - not in this image
    from pettingzoo import AECEnv as _AECEnv
except ImportError:
    class _AECEnv:
        def __init__(self):
            pass

        def _was_dead_step(self, action):
            return None


class Card:
    def __init__(self, rank: int, suit: str):
        self.rank = rank
        self.suit = suit

    def __repr__(self):
        return f"{self.rank}_of_{self.suit}"


def _card(cid):
    return Card(codec.RANK_OF[cid], codec.SUIT_OF[cid])


def _cid(card):
    return codec.card_id(card.rank, card.suit)


# ------------------------------------------------------------------------------- device helpers
def _deck_from_seed(seed):
    """MiniDeck(seed).cards as 16 card ids (ms_deck_from_seeds)."""
    import torch
    lib = _lib.load()
    s = torch.tensor([int(seed)], dtype=torch.int64, device="cuda")
    out = torch.empty(1, dtype=torch.int64, device="cuda")
    _lib.check(lib.ms_deck_from_seeds(s.data_ptr(), 1, out.data_ptr(), _lib.stream_ptr()))
    perm = int(out.cpu().numpy().view(np.uint64)[0])
    return codec.nibbles(perm, 16)


def _coerce_seed(seed):
    seed = int(seed)
    if not -(2 ** 63) <= seed < 2 ** 63:
        raise OverflowError("scopa_b200 supports seeds in the signed 64-bit range")
    return seed


class MiniDeck:
    """16-card deck: 4 suits 4 ranks each, pairwise duplicated ranks across suits."""
    suits = list(codec.SUITS)
    ranks = {k: list(v) for k, v in codec.RANKS.items()}

    def __init__(self, seed=42):
        self.cards = [_card(c) for c in _deck_from_seed(_coerce_seed(seed))]

    def deal(self, n):
        dealt = self.cards[:n]
        self.cards = self.cards[n:]
        return dealt


class Player:
    def __init__(self, name):
        self.name = name
        self.hand = []
        self.captures = []
        self.scopas = 0

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
        self.table = []
        self.last_capture = None

    def reset(self, seed=42):
        self.deck = MiniDeck(seed)
        self.table.clear()
        cards_per_player = 4
        for p in self.players:
            p.reset()
            p.hand = self.deck.deal(cards_per_player)
        self.last_capture = None

    # -- packed-state plumbing ------------------------------------------------------------------
    def _pack(self, cur, step_count=0, terminal=False, max_steps=31):
        hands = [[_cid(c) for c in p.hand] for p in self.players]
        caps = [[_cid(c) for c in p.captures] for p in self.players]
        table = [_cid(c) for c in self.table]
        if len(table) > 8 or any(len(h) > 4 for h in hands):
            raise ValueError("Miniscopa state out of range (table > 8 cards or hand > 4 cards)")
        words = codec.pack_state([codec.mask_of(h) for h in hands], table, [codec.mask_of(c) for c in caps],
                                 [p.scopas for p in self.players], step_count, cur, terminal, max_steps)
        order = []
        for h in hands:
            pad = [c for c in range(16) if c not in h]
            order += (h + pad)[:4]
        return words, codec.pack_nibbles(order)

    def _apply_unpacked(self, u, old_table_ids, played_id, mover):
        """Re-format the lists from an unpacked device state after one transition."""
        pl = self.players[mover]
        new_table = u["table"]
        if played_id is not None:
            card_obj = next(c for c in pl.hand if _cid(c) == played_id)
            captured_ids = [c for c in old_table_ids if c not in new_table]
            if captured_ids or (played_id not in new_table):
                objs = {_cid(c): c for c in self.table}
                pl.captures.extend([objs[c] for c in captured_ids] + [card_obj])   # captured + [card] (:98)
                self.last_capture = pl
                self.table = [c for c in self.table if _cid(c) not in captured_ids]
            else:
                self.table.append(card_obj)
            pl.hand.remove(card_obj)
        for i, p in enumerate(self.players):
            p.scopas = u["scopas"][i]

    def card_in_table(self, card):
        """Find subset of table cards that sum to card's rank (reference :66-91), on the device."""
        import torch
        lib = _lib.load()
        if card.rank <= 0 or not self.table:
            return False, []
        cid = _cid(card)
        if cid < 0:
            raise ValueError(f"{card!r} is not a Miniscopa card")
        words, _ = self._pack(0)
        st = torch.tensor(np.array([words], dtype=np.uint32).view(np.int32), device="cuda")
        cd = torch.tensor([cid], dtype=torch.uint8, device="cuda")
        out = torch.empty(1, dtype=torch.uint8, device="cuda")
        _lib.check(lib.ms_capture(st.data_ptr(), cd.data_ptr(), out.data_ptr(), 1, _lib.stream_ptr()))
        mask = int(out.item())
        combo = [self.table[i] for i in range(len(self.table)) if (mask >> i) & 1]
        return bool(mask), combo

    def play_card(self, card, player):
        mover = self.players.index(player)
        if card not in player.hand:
            raise ValueError("list.remove(x): x not in list")
        words, _ = self._pack(mover)
        st = np.array([words], dtype=np.uint32)
        act = np.array([_cid(card)], dtype=np.uint8)
        _lib.check(_lib.load().ms_step_host(st.ctypes.data, act.ctypes.data, None, None, 1))
        self._apply_unpacked(codec.unpack_state(st[0]), [_cid(c) for c in self.table], _cid(card), mover)

    def evaluate_game(self):
        """Final reward as zero-sum vector (+1 per capture, +2 per scopa), scored on the device."""
        words, _ = self._pack(0, step_count=30, max_steps=31)
        st = np.array([words], dtype=np.uint32)
        act = np.array([255], dtype=np.uint8)          # a pass that ends the game: step 31 >= max_steps 31
        rew = np.zeros((1, 2), dtype=np.float32)
        _lib.check(_lib.load().ms_step_host(st.ctypes.data, act.ctypes.data, rew.ctypes.data, None, 1))
        r = [float(rew[0, 0]), float(rew[0, 1])]
        if r[0] == 0.0 and r[1] == 0.0 and sum(len(p.captures) + 2 * p.scopas for p in self.players) == 0:
            return [0] * self.num_players
        return r


class MiniScopaEnv(_AECEnv):
    metadata = {"name": "Mini-Scopa-v0"}

    def __init__(self, seed=42, num_players=2):
        super().__init__()
        self.num_players = num_players
        self.game = MiniScopaGame(num_players=num_players)
        self.possible_agents = [f"player_{i}" for i in range(num_players)]
        self.agent_name_mapping = {name: i for i, name in enumerate(self.possible_agents)}
        self._action_spaces = {a: _Discrete(16) for a in self.possible_agents}
        self.max_steps = num_players * 4
        self.seed = seed
        self.reset(seed)

    def reset(self, seed=None):
        self.game.reset(seed or self.seed)
        self.agents = self.possible_agents[:]
        self.agent_selection = self.agents[0]
        self.rewards = {a: 0 for a in self.agents}
        self.terminations = {a: False for a in self.agents}
        self.truncations = {a: False for a in self.agents}
        self.step_count = 0

    def step(self, action):
        if self.terminations[self.agent_selection]:
            self._was_dead_step(action)
            return
        agent = self.agent_selection
        mover = self.agent_name_mapping[agent]
        g = self.game
        words, _ = g._pack(mover, self.step_count, False, min(self.max_steps, 31))
        st = np.array([words], dtype=np.uint32)
        a = int(action)
        act = np.array([a if 0 <= a < 16 else 255], dtype=np.uint8)
        rew = np.zeros((1, 2), dtype=np.float32)
        done = np.zeros(1, dtype=np.uint8)
        old_table = [_cid(c) for c in g.table]
        hand_ids = [_cid(c) for c in g.players[mover].hand]
        _lib.check(_lib.load().ms_step_host(st.ctypes.data, act.ctypes.data, rew.ctypes.data, done.ctypes.data, 1))
        u = codec.unpack_state(st[0])
        played = a if (0 <= a < 16 and a in hand_ids) else None      # otherwise: silent pass (:155-157)
        g._apply_unpacked(u, old_table, played, mover)
        self.step_count = u["step_count"]
        if done[0]:
            for i, ag in enumerate(self.agents):
                r = float(rew[0, i])
                self.rewards[ag] = 0 if (rew[0, 0] == 0 and rew[0, 1] == 0 and not any(
                    len(p.captures) + 2 * p.scopas for p in g.players)) else r
                self.terminations[ag] = True
        self.agent_selection = self.agents[(self.agents.index(agent) + 1) % self.num_players]

    def get_state(self):
        return {
            "table": [(c.rank, c.suit) for c in self.game.table],
            "hands": [[(c.rank, c.suit) for c in p.hand] for p in self.game.players],
            "captures": [[(c.rank, c.suit) for c in p.captures] for p in self.game.players],
            "scopas": [p.scopas for p in self.game.players],
            "agent_selection": self.agent_selection,
            "step_count": self.step_count,
            "agents": self.agents[:],
            "rewards": dict(self.rewards),
            "terminations": dict(self.terminations),
            "truncations": dict(self.truncations),
        }

    def set_state(self, state):
        self.game.table = [Card(r, s) for r, s in state["table"]]
        for i, p in enumerate(self.game.players):
            p.hand = [Card(r, s) for r, s in state["hands"][i]]
            p.captures = [Card(r, s) for r, s in state["captures"][i]]
            p.scopas = state["scopas"][i]
        self.agent_selection = state["agent_selection"]
        self.step_count = state["step_count"]
        self.agents = state["agents"][:]
        self.rewards = dict(state["rewards"])
        self.terminations = dict(state["terminations"])
        self.truncations = dict(state["truncations"])

    # -- used by the solver classes: the packed form of the current state ---------------------------
    def packed(self):
        mover = self.agent_name_mapping[self.agent_selection]
        term = all(self.terminations.values())
        return self.game._pack(mover, self.step_count, term, min(self.max_steps, 31))
