"""Host-side codec between the packed device formats and the reference's Python-visible forms.

Pure data re-formatting (no game rules): card id <-> (rank, suit), packed 16-byte state <-> lists,
64-bit infoset key <-> "P0:H[9f-6p-5f-7f]_T[]" strings
(reference src/envs/openspiel_mini_scopa.py:86-95, src/envs/mini_scopa_game.py:17-23, :149-153).
"""
SUITS = ["cuori", "fiori", "picche", "bello"]
RANKS = {"cuori": [2, 5, 8, 10], "fiori": [2, 5, 7, 9], "picche": [3, 6, 8, 9], "bello": [3, 6, 7, 10]}
RANK_OF = [RANKS[SUITS[c // 4]][c % 4] for c in range(16)]
SUIT_OF = [SUITS[c // 4] for c in range(16)]
CARD_STR = [f"{RANK_OF[c]}{SUIT_OF[c][0]}" for c in range(16)]
CARD_ID = {(RANK_OF[c], SUIT_OF[c]): c for c in range(16)}
TERMINAL_KEY = 0xFFFFFFFFFFFFFFFF


def card_id(rank, suit):
    return CARD_ID.get((rank, suit), -1)


def nibbles(word, n):
    return [(word >> (4 * i)) & 0xF for i in range(n)]


def pack_nibbles(cards):
    w = 0
    for i, c in enumerate(cards):
        w |= (c & 0xF) << (4 * i)
    return w


def mask_of(cards):
    m = 0
    for c in cards:
        m |= 1 << c
    return m


def hand_in_order(hand_mask, hand_order, player):
    """Cards of `hand_mask` listed in deal order (nibbles 4*player .. 4*player+3 of hand_order)."""
    out = []
    for c in nibbles(hand_order >> (16 * player), 4):
        if (hand_mask >> c) & 1 and c not in out:
            out.append(c)
    return out


def unpack_state(words):
    """(hands, table, captures, meta) uint32 words -> dict of plain Python fields."""
    x, y, z, w = (int(v) & 0xFFFFFFFF for v in words)
    tlen = w & 0xF
    return {
        "hand_mask": [x & 0xFFFF, x >> 16],
        "table": nibbles(y, tlen),
        "cap_mask": [z & 0xFFFF, z >> 16],
        "scopas": [(w >> 4) & 0xF, (w >> 8) & 0xF],
        "step_count": (w >> 12) & 0x1F,
        "cur": (w >> 17) & 1,
        "terminal": bool((w >> 18) & 1),
        "max_steps": (w >> 19) & 0x1F,
    }


def pack_state(hand_mask, table, cap_mask, scopas, step_count, cur, terminal, max_steps):
    x = (hand_mask[0] & 0xFFFF) | ((hand_mask[1] & 0xFFFF) << 16)
    y = pack_nibbles(table)
    z = (cap_mask[0] & 0xFFFF) | ((cap_mask[1] & 0xFFFF) << 16)
    w = (len(table) & 0xF) | ((scopas[0] & 0xF) << 4) | ((scopas[1] & 0xF) << 8) | ((step_count & 0x1F) << 12) \
        | ((cur & 1) << 17) | ((1 if terminal else 0) << 18) | ((max_steps & 0x1F) << 19)
    return (x, y, z, w)


def key_fields(key):
    key = int(key)
    return {"player": (key >> 52) & 1, "hand_mask": (key >> 36) & 0xFFFF, "table": nibbles(key & 0xFFFFFFFF, (key >> 32) & 0xF)}


def key_to_string(key, hand_order):
    """64-bit infoset key -> the reference's information_state_string."""
    if int(key) == TERMINAL_KEY:
        return "TERMINAL"
    f = key_fields(key)
    hand = hand_in_order(f["hand_mask"], hand_order, f["player"])
    return "P%d:H[%s]_T[%s]" % (f["player"], "-".join(CARD_STR[c] for c in hand), "-".join(CARD_STR[c] for c in f["table"]))


def string_to_key(info):
    """The reference's information_state_string -> 64-bit key (inverse of key_to_string)."""
    if info == "TERMINAL":
        return TERMINAL_KEY
    player = int(info[1])
    hand_part = info.split("H[")[1].split("]")[0]
    table_part = info.split("T[")[1].split("]")[0]
    sm = {"c": 0, "f": 1, "p": 2, "b": 3}

    def parse(part):
        out = []
        for cs in part.split("-") if part else []:
            out.append(CARD_ID[(int(cs[:-1]), SUITS[sm[cs[-1]]])])
        return out

    hand, table = parse(hand_part), parse(table_part)
    return (player << 52) | (mask_of(hand) << 36) | (len(table) << 32) | pack_nibbles(table)
