"""40-card Scopa under the reference's names, evaluated on the GPU (csrc/ms_full.cu).

Drop-in for /root/reference/src/envs/full_scopa_game.py:6-342 -- `Card`, `FullDeck`, `Player`, `FullScopaGame`
and the PettingZoo-style `FullScopaEnv` (reset / step / get_state / set_state, `agent_selection`, `rewards`,
`terminations`, `truncations`, `step_count`, `max_steps`, `seed`, `game`), two players.

Every rule evaluation -- the seeded shuffle, capture resolution, dealing of new hands, the sweep and the
traditional scoring -- is a kernel behind the C ABI; the Python lists exist because the reference exposes them
(`env.game.players[i].hand / .captures / .scopas`, `env.game.table`, `env.game.deck.cards`) and follow the packed
device state after each call.  One launch per call: this scalar API is for compatibility, batches go through
scopa_b200.full.BatchedFullScopa.

Differences, both on error paths: an action id outside 0..39 is a silent pass (IndexError in the reference, :264),
and `play_card` only takes the first capture combination (the only one `FullScopaEnv.step` ever asks for).
"""
import numpy as np

from .. import _lib, full
from .mini_scopa_game import _AECEnv, _Discrete

_PASS = 255


class Card:
    def __init__(self, rank: int, suit: str):
        self.rank, self.suit = rank, suit

    def __repr__(self):
        return f"{self.rank}_{self.suit}"

    def __eq__(self, other):
        return (self.rank, self.suit) == (other.rank, other.suit)

    def __hash__(self):
        return hash((self.rank, self.suit))


def _card(c):
    return Card(*full.card_rank_suit(c))


def _cid(card):
    return full.card_id(card.rank, card.suit)


def _ids(cards):
    return [_cid(c) for c in cards]


class FullDeck:
    """Standard Italian 40-card deck; the seeded shuffle runs on the device (CPython random.seed / shuffle)."""
    suits = list(full.SUITS)
    ranks = list(range(1, 11))
    primiera_values = {7: 21, 6: 18, 1: 16, 5: 15, 4: 14, 3: 13, 2: 12, 10: 10, 9: 10, 8: 10}

    def __init__(self, seed=42):
        self.cards = [_card(c) for c in full.deck_from_seed(int(seed))]

    def deal(self, n):
        dealt, self.cards = self.cards[:n], self.cards[n:]
        return dealt

    def cards_remaining(self):
        return len(self.cards)


class Player:
    def __init__(self, name):
        self.name = name
        self.hand, self.captures, self.scopas = [], [], 0

    def reset(self):
        self.hand.clear()
        self.captures.clear()
        self.scopas = 0


class FullScopaGame:
    def __init__(self, num_players=2):
        if num_players != 2:
            raise NotImplementedError("the CUDA 40-card env is built for two players (the registered OpenSpiel game)")
        self.num_players = num_players
        self.deck = FullDeck()
        self.players = [Player(f"player_{i}") for i in range(num_players)]
        self.table, self.last_capture, self.round_number, self.cards_per_hand = [], None, 0, 3

    def reset(self, seed=42):
        self.deck = FullDeck(seed)
        self.round_number = 0
        for p in self.players:
            p.reset()
        self.table = self.deck.deal(4)
        for p in self.players:
            p.hand = self.deck.deal(self.cards_per_hand)
        self.last_capture = None

    def can_deal_new_round(self):
        return self.deck.cards_remaining() >= self.num_players * self.cards_per_hand

    def deal_new_round(self):
        if not self.can_deal_new_round():
            return False
        for p in self.players:
            p.hand = self.deck.deal(self.cards_per_hand)
        self.round_number += 1
        return True

    # ---- packed device form of the lists above.  The device addresses hands through the deck (round r deals
    # deck[4 + 6 r + 3 p + i]); whatever the lists hold is laid out that way: the current hands at this round's
    # positions, the undealt cards behind them, so set_state() may install any position.
    def _pack(self, cur, step_count=0, terminal=False):
        r = self.round_number
        if not 0 <= r <= 5 or any(len(p.hand) > 3 for p in self.players):
            raise ValueError("state outside the 40-card game (round 0..5, at most three cards in hand)")
        deck = [0] * 40
        bits = []
        for pi, p in enumerate(self.players):
            ids = _ids(p.hand)
            deck[4 + 6 * r + 3 * pi: 4 + 6 * r + 3 * pi + len(ids)] = ids
            bits.append((1 << len(ids)) - 1)
        rest = _ids(self.deck.cards)[: 40 - (10 + 6 * r)]
        deck[10 + 6 * r: 10 + 6 * r + len(rest)] = rest
        caps = [sum(1 << c for c in set(_ids(p.captures))) for p in self.players]
        last = None if self.last_capture is None else self.players.index(self.last_capture)
        words = full.pack_full_state(_ids(self.table), caps, bits, last, cur, terminal, [p.scopas for p in self.players], r,
                                     step_count)
        return words, full.pack_deck(deck)

    def _absorb(self, u, played, mover, table_before):
        """Bring the lists up to the unpacked device state `u` after the mover played `played` (None = pass)."""
        pl = self.players[mover]
        if played is not None:
            card = next(c for c in pl.hand if _cid(c) == played)
            gone = [c for c in table_before if c not in u["table"]]
            captured = bool(gone) or (u["last_capture"] == mover and played not in u["table"])
            if captured:
                by_id = {_cid(c): c for c in self.table}
                pl.captures.extend([by_id[c] for c in gone] + [card])
                self.table = [c for c in self.table if _cid(c) not in gone]
                self.last_capture = pl
            else:
                self.table.append(card)
            pl.hand.remove(card)
        for p, n in zip(self.players, u["scopas"]):
            p.scopas = n
        if u["round_number"] > self.round_number:
            self.deal_new_round()

    def _sweep(self, times):
        if self.table and self.last_capture is not None:
            for _ in range(times):
                self.last_capture.captures.extend(self.table)      # the table itself stays (:187-188)

    def _device_step(self, cur, step_count, action):
        words, deck = self._pack(cur, step_count)
        st = np.asarray([words], dtype=np.uint32)
        dk = np.asarray([deck], dtype=np.uint64)
        act = np.asarray([action], dtype=np.uint8)
        rew, done = np.zeros((1, 2), dtype=np.float32), np.zeros(1, dtype=np.uint8)
        _lib.check(_lib.load().ms_full_step_host(st.ctypes.data, dk.ctypes.data, act.ctypes.data, rew.ctypes.data,
                                                 done.ctypes.data, 1))
        return full.unpack_full_state(st[0]), [float(x) for x in rew[0]], bool(done[0])

    def find_capture_combinations(self, card):
        """The combination play_card would take, as a one-element list ([] when the card would be placed).  The
        reference lists every subset here but FullScopaEnv only ever uses the first (:137-141)."""
        if not self.table or card.rank <= 0:
            return []
        probe = FullScopaGame.__new__(FullScopaGame)
        probe.num_players, probe.players, probe.cards_per_hand = 2, [Player("a"), Player("b")], 3
        probe.players[0].hand = [card]
        probe.table, probe.last_capture, probe.round_number, probe.deck = list(self.table), None, 0, self.deck
        before = _ids(probe.table)
        u, _, _ = probe._device_step(0, 0, _cid(card))
        gone = [c for c in before if c not in u["table"]]
        by_id = {_cid(c): c for c in self.table}
        return [[by_id[c] for c in gone]] if gone else []

    def play_card(self, card, player, capture_choice=None):
        if capture_choice not in (None, 0):
            raise NotImplementedError("only the first capture combination is supported (the one the env takes)")
        if card not in player.hand:
            raise ValueError("list.remove(x): x not in list")
        mover = self.players.index(player)
        before = _ids(self.table)
        saved_round = self.round_number
        u, _, _ = self._device_step(mover, 0, _cid(card))
        u["round_number"] = saved_round            # dealing and scoring belong to the env's step, not to play_card
        self._absorb(u, _cid(card), mover, before)

    @staticmethod
    def _device_evaluate(words):
        """-> rewards [2], detail [cards0, cards1, denari0, denari1, primiera0, primiera1, score0, score1]"""
        st = np.asarray([words], dtype=np.uint32)
        rew, det = np.zeros((1, 2), dtype=np.float32), np.zeros((1, 8), dtype=np.int32)
        _lib.check(_lib.load().ms_full_evaluate_host(st.ctypes.data, rew.ctypes.data, det.ctypes.data, 1))
        return [float(x) for x in rew[0]], [int(x) for x in det[0]]

    def calculate_primiera_score(self, captured_cards):
        mask = sum(1 << c for c in set(_ids(captured_cards)))
        words = full.pack_full_state([], [mask, 0], [0, 0], None, 0, False, [0, 0], 0, 0)
        return self._device_evaluate(words)[1][4]

    def evaluate_game(self):
        rew, det = self._device_evaluate(self._pack(0, 0)[0])
        self._sweep(1)
        return rew if det[6] + det[7] else [0] * self.num_players


class FullScopaEnv(_AECEnv):
    metadata = {"name": "Full-Scopa-v0"}

    def __init__(self, seed=42, num_players=2):
        super().__init__()
        self.num_players = num_players
        self.game = FullScopaGame(num_players=num_players)
        self.possible_agents = [f"player_{i}" for i in range(num_players)]
        self.agent_name_mapping = {name: i for i, name in enumerate(self.possible_agents)}
        self._action_spaces = {a: _Discrete(40) for a in self.possible_agents}
        self.max_steps = 200
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
        if self.max_steps != 200:
            raise NotImplementedError("the device env implements the reference's fixed 200-step safety limit")
        g = self.game
        mover = self.agent_name_mapping[self.agent_selection]
        a = int(action)
        in_hand = 0 <= a < 40 and a in _ids(g.players[mover].hand)
        before = _ids(g.table)
        u, rew, done = g._device_step(mover, self.step_count, a if 0 <= a < 40 else _PASS)
        g._absorb(u, a if in_hand else None, mover, before)
        self.step_count = u["step_count"]
        if done:
            g._sweep(2 if u["evaluated_twice"] else 1)
            for i, name in enumerate(self.agents):
                self.rewards[name] = rew[i]
                self.terminations[name] = True
        self.agent_selection = self.agents[(mover + 1) % self.num_players]

    def get_state(self):
        g = self.game
        pairs = lambda cards: [(c.rank, c.suit) for c in cards]
        return {
            "table": pairs(g.table), "hands": [pairs(p.hand) for p in g.players],
            "captures": [pairs(p.captures) for p in g.players], "scopas": [p.scopas for p in g.players],
            "deck_remaining": g.deck.cards_remaining(), "round_number": g.round_number,
            "last_capture": g.players.index(g.last_capture) if g.last_capture else None,
            "agent_selection": self.agent_selection, "step_count": self.step_count, "agents": self.agents[:],
            "rewards": dict(self.rewards), "terminations": dict(self.terminations), "truncations": dict(self.truncations),
        }

    def set_state(self, state):
        g = self.game
        g.deck = FullDeck()                                   # like the reference: the seed-42 deck (:314-317)
        g.deck.cards = g.deck.cards[40 - state["deck_remaining"]:]
        g.table = [Card(r, s) for r, s in state["table"]]
        for p, hand, caps, n in zip(g.players, state["hands"], state["captures"], state["scopas"]):
            p.hand, p.captures, p.scopas = [Card(r, s) for r, s in hand], [Card(r, s) for r, s in caps], n
        g.round_number = state["round_number"]
        g.last_capture = None if state["last_capture"] is None else g.players[state["last_capture"]]
        self.agent_selection, self.step_count = state["agent_selection"], state["step_count"]
        self.agents = state["agents"][:]
        for k in ("rewards", "terminations", "truncations"):
            setattr(self, k, dict(state[k]))
