"""Batched 40-card Scopa on the GPU (csrc/ms_full.cu) and the codec of its packed state.

One row = one game; states are torch.int32 [n, 8] (bit patterns of the uint32 words), decks torch.int64 [n, 4]
(ten 6-bit card ids per word: the table is deck[0..3], round r deals deck[4 + 6 r + 3 p + i] to player p).
Replaces, n games at a time, FullScopaEnv.reset/step of /root/reference/src/envs/full_scopa_game.py:243-296.
Card id = suit_idx * 10 + rank - 1 = the reference's action id.
"""
import torch

from . import _lib

SUITS = ["denari", "coppe", "spade", "bastoni"]
PLIES = 36
MAX_TABLE = 16


def card_id(rank, suit):
    return SUITS.index(suit) * 10 + (rank - 1)


def card_rank_suit(c):
    return c % 10 + 1, SUITS[c // 10]


def unpack_deck(words):
    out = []
    for w in words:
        w = int(w) & 0xFFFFFFFFFFFFFFFF
        out.extend((w >> (6 * k)) & 0x3F for k in range(10))
    return out


def pack_deck(cards):
    cards = list(cards) + [0] * (40 - len(cards))
    return [sum((cards[10 * q + k] & 0x3F) << (6 * k) for k in range(10)) for q in range(4)]


def unpack_full_state(words, deck=None):
    """packed words -> dict; with `deck` (40 ids) the hands are listed as card ids in hand order"""
    w = [int(x) & 0xFFFFFFFF for x in words]
    n = (w[5] >> 22) & 0x1F
    lo, hi = w[0] | (w[1] << 32), w[2] | (w[7] << 32)
    table = [((lo >> (6 * i)) if i < 10 else (hi >> (6 * (i - 10)))) & 0x3F for i in range(n)]
    last = (w[5] >> 27) & 3
    diff = (w[6] >> 23) & 0xFF
    out = {
        "table": table,
        "cap_mask": [w[3] | ((w[5] & 0xFF) << 32), w[4] | (((w[5] >> 8) & 0xFF) << 32)],
        "hand_bits": [(w[5] >> 16) & 7, (w[5] >> 19) & 7],
        "last_capture": None if last == 0 else last - 1,
        "cur": (w[5] >> 29) & 1, "terminal": bool((w[5] >> 30) & 1), "evaluated_twice": bool((w[5] >> 31) & 1),
        "scopas": [w[6] & 0x3F, (w[6] >> 6) & 0x3F], "round_number": (w[6] >> 12) & 7, "step_count": (w[6] >> 15) & 0xFF,
        "score_diff": diff - 256 if diff >= 128 else diff,
    }
    if deck is not None:
        base = 4 + 6 * out["round_number"]
        out["hands"] = [[deck[base + 3 * p + i] for i in range(3) if (out["hand_bits"][p] >> i) & 1] for p in range(2)]
    return out


def pack_full_state(table, cap_mask, hand_bits, last_capture, cur, terminal, scopas, round_number, step_count,
                    score_diff=0, evaluated_twice=False):
    if len(table) > MAX_TABLE:
        raise ValueError(f"the packed state holds at most {MAX_TABLE} table cards")
    lo = sum((c & 0x3F) << (6 * i) for i, c in enumerate(table[:10]))
    hi = sum((c & 0x3F) << (6 * i) for i, c in enumerate(table[10:]))
    w5 = ((cap_mask[0] >> 32) & 0xFF) | (((cap_mask[1] >> 32) & 0xFF) << 8) | ((hand_bits[0] & 7) << 16) | ((hand_bits[1] & 7) << 19) \
        | (len(table) << 22) | (((-1 if last_capture is None else last_capture) + 1) << 27) | ((cur & 1) << 29) \
        | ((1 if terminal else 0) << 30) | ((1 if evaluated_twice else 0) << 31)
    w6 = (scopas[0] & 0x3F) | ((scopas[1] & 0x3F) << 6) | ((round_number & 7) << 12) | ((step_count & 0xFF) << 15) \
        | ((score_diff & 0xFF) << 23)
    return (lo & 0xFFFFFFFF, lo >> 32, hi & 0xFFFFFFFF, cap_mask[0] & 0xFFFFFFFF, cap_mask[1] & 0xFFFFFFFF, w5, w6, hi >> 32)


class BatchedFullScopa:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MsError("scopa_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        self.states = None
        self.decks = None

    @property
    def n(self):
        return 0 if self.states is None else self.states.shape[0]

    def reset(self, seeds):
        seeds = torch.as_tensor(seeds, dtype=torch.int64).to(self.device).contiguous()
        n = seeds.numel()
        with torch.cuda.device(self.device):
            self.states = torch.empty((n, 8), dtype=torch.int32, device=self.device)
            self.decks = torch.empty((n, 4), dtype=torch.int64, device=self.device)
            _lib.check(self.lib.ms_full_deal_from_seeds(seeds.data_ptr(), n, self.states.data_ptr(), self.decks.data_ptr(),
                                                        _lib.stream_ptr()))
        return self

    def step(self, actions):
        """actions [n] u8 -> rewards [n, 2] f32 (zero while running), done [n] u8"""
        n = self.n
        actions = actions.to(self.device, torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            rewards = torch.empty((n, 2), dtype=torch.float32, device=self.device)
            done = torch.empty((n,), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.ms_full_step(self.states.data_ptr(), self.decks.data_ptr(), actions.data_ptr(), rewards.data_ptr(),
                                             done.data_ptr(), n, _lib.stream_ptr()))
        return rewards, done

    def legal_actions(self, player=-1):
        """-> ordered [n, 3] u8 (hand order, 0xFF padded), count [n] u8; player -1 = each game's mover"""
        n = self.n
        with torch.cuda.device(self.device):
            ordered = torch.empty((n, 3), dtype=torch.uint8, device=self.device)
            count = torch.empty((n,), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.ms_full_legal_actions(self.states.data_ptr(), self.decks.data_ptr(), player, ordered.data_ptr(),
                                                      count.data_ptr(), n, _lib.stream_ptr()))
        return ordered, count

    def rollout_random(self, philox_seed=0, game_offset=0):
        """36 uniform-random legal plies per game -> actions [n, 36] u8, rewards [n, 2] f32, final states [n, 8]."""
        n = self.n
        with torch.cuda.device(self.device):
            actions = torch.empty((n, PLIES), dtype=torch.uint8, device=self.device)
            rewards = torch.empty((n, 2), dtype=torch.float32, device=self.device)
            final = torch.empty((n, 8), dtype=torch.int32, device=self.device)
            _lib.check(self.lib.ms_full_rollout_random(self.states.data_ptr(), self.decks.data_ptr(), n, philox_seed, game_offset,
                                                       actions.data_ptr(), rewards.data_ptr(), final.data_ptr(), _lib.stream_ptr()))
        return actions, rewards, final

    def table_overflow(self):
        import ctypes as C
        flag = C.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ms_full_table_overflow(C.byref(flag), _lib.stream_ptr()))
        return bool(flag.value)


def deck_from_seed(seed, device="cuda", slow_path=False):
    """FullDeck(seed).cards as 40 card ids (no seed substitution)."""
    lib = _lib.load()
    dev = torch.device(device)
    s = torch.tensor([seed], dtype=torch.int64, device=dev)
    d = torch.empty((1, 4), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ms_full_deck_from_seeds(s.data_ptr(), 1, d.data_ptr(), 1 if slow_path else 0, _lib.stream_ptr()))
    return unpack_deck(d.cpu().numpy().view("uint64")[0])
