"""The PRODUCT's 40-card Scopa code on the CPU: scopa_b200/csrc/ms_full.cu (fs_step, fs_capture_mask, fs_evaluate,
fs_legal_list and the init / step / legal / evaluate / rollout kernels themselves) compiled for the host by
tests/emu/ms_full_host.cpp, checked bit for bit against the traces recorded from the unmodified reference
(full_scopa_game.py, incl. the 200-step limit and the double evaluation at step 200) and, for the fused 36-ply rollout,
against the oracle.  Same fixtures and assertions as tests/test_gpu_full.py, which runs the device build of this source."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import load_golden_json
from oracle import ms_oracle as ora
from scopa_b200 import full

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_build  # noqa: E402

vp = C.c_void_p


@pytest.fixture(scope="module")
def fs():
    lib = C.CDLL(emu_build.build_full_host())
    lib.host_full_init.argtypes = [vp, C.c_longlong, vp]
    lib.host_full_step.argtypes = [vp, vp, vp, vp, vp, C.c_longlong]
    lib.host_full_step.restype = C.c_uint
    lib.host_full_legal.argtypes = [vp, vp, C.c_int, vp, vp, C.c_longlong]
    lib.host_full_evaluate.argtypes = [vp, vp, vp, C.c_longlong]
    lib.host_full_rollout.argtypes = [vp, vp, C.c_longlong, C.c_ulonglong, C.c_ulonglong, vp, vp, vp]
    lib.host_full_rollout.restype = C.c_uint
    return lib


def trace_deck(t):
    """FullScopaEnv.reset(seed): `seed or self.seed` (full_scopa_game.py:243-250), the fixtures' envs were built with seed 42"""
    return ora.full_deck(t["seed"] or 42)


def pack_decks(decks):
    return np.array([full.pack_deck(d) for d in decks], dtype=np.uint64)


def init(fs, decks):
    st = np.zeros((len(decks), 8), np.uint32)
    fs.host_full_init(decks.ctypes.data, len(decks), st.ctypes.data)
    return st


def step(fs, st, decks, actions):
    n = st.shape[0]
    rew, done = np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
    actions = np.ascontiguousarray(actions, np.uint8)
    over = fs.host_full_step(st.ctypes.data, decks.ctypes.data, actions.ctypes.data, rew.ctypes.data, done.ctypes.data, n)
    return rew, done, over


def legal(fs, st, decks, player=-1):
    n = st.shape[0]
    ordered, count = np.zeros((n, 3), np.uint8), np.zeros(n, np.uint8)
    fs.host_full_legal(st.ctypes.data, decks.ctypes.data, player, ordered.ctypes.data, count.ctypes.data, n)
    return ordered, count


def test_initial_state_matches_codec_and_reference(fs):
    g = load_golden_json("full_env_traces.json.gz")
    for seed, deck in g["decks"].items():
        assert ora.full_deck(int(seed)) == deck                   # the oracle's shuffle is what feeds the tests below
    traces = g["traces"]
    cards = [trace_deck(t) for t in traces]
    decks = pack_decks(cards)
    st = init(fs, decks)
    for i, t in enumerate(traces):
        snap, u = t["snaps"][0], full.unpack_full_state(st[i], cards[i])
        assert u["table"] == snap["table"] == cards[i][:4] and u["hands"] == snap["hands"], t["seed"]
        assert tuple(int(x) for x in st[i]) == full.pack_full_state(cards[i][:4], [0, 0], [7, 7], None, 0, False, [0, 0], 0, 0)


def test_kernels_follow_reference_traces(fs):
    """Every trace at once through full_step_kernel / full_legal_kernel, packed state against the recorded lists."""
    traces = load_golden_json("full_env_traces.json.gz")["traces"]
    n_steps = max(len(t["actions"]) for t in traces)
    cards = [trace_deck(t) for t in traces]
    decks = pack_decks(cards)
    st = init(fs, decks)
    acts = np.full((len(traces), n_steps), 255, dtype=np.uint8)
    for i, t in enumerate(traces):
        acts[i, :len(t["actions"])] = t["actions"]
    for k in range(n_steps):
        rew, done, over = step(fs, st, decks, acts[:, k])
        assert not over
        ordered, count = legal(fs, st, decks)
        for i, t in enumerate(traces):
            if k >= len(t["actions"]):
                continue
            snap, u = t["snaps"][k + 1], full.unpack_full_state(st[i], cards[i])
            assert u["table"] == snap["table"] and u["hands"] == snap["hands"], (t["seed"], k)
            assert u["cap_mask"] == [sum(1 << c for c in set(cs)) for cs in snap["caps"]], (t["seed"], k)
            assert u["scopas"] == snap["scopas"] and u["step_count"] == snap["step"] and u["round_number"] == snap["round"]
            assert u["last_capture"] == snap["last"] and f"player_{u['cur']}" == snap["agent"]
            assert [u["terminal"]] * 2 == snap["term"] and bool(done[i]) == snap["term"][0]
            assert rew[i].tolist() == snap["rew"], (t["seed"], k)
            # FullScopaState.legal_actions of the mover: the hand in deal order, [0] for an empty hand, [] when over
            hand = u["hands"][u["cur"]]
            want = [] if u["terminal"] else (hand if hand else [0])
            assert ordered[i, :count[i]].tolist() == want and np.all(ordered[i, count[i]:] == 0xFF), (t["seed"], k)


def test_rollout_matches_oracle(fs):
    n = 20_000
    seeds = np.arange(1, n + 1, dtype=np.int64)
    cards = [ora.full_deck(int(s)) for s in seeds]
    decks = pack_decks(cards)
    st = init(fs, decks)
    acts, rew, fin = np.zeros((n, 36), np.uint8), np.zeros((n, 2), np.float32), np.zeros((n, 8), np.uint32)
    over = fs.host_full_rollout(st.ctypes.data, decks.ctypes.data, n, 77, 3, acts.ctypes.data, rew.ctypes.data, fin.ctypes.data)
    o_act, o_rew, o_sc, o_nc, o_mt = ora.full_rollout_random(seeds, 77, game_offset=3)
    assert not over and o_mt.max() <= full.MAX_TABLE
    assert np.array_equal(acts, o_act) and np.array_equal(rew, o_rew)
    assert np.array_equal(np.stack([fin[:, 6] & 0x3F, (fin[:, 6] >> 6) & 0x3F], 1).astype(np.uint8), o_sc)
    pop = np.array([[bin(c).count("1") for c in full.unpack_full_state(w)["cap_mask"]] for w in fin[:3000]])
    assert np.array_equal(pop.astype(np.uint8), o_nc[:3000])
    assert np.all(rew[:, 0] + rew[:, 1] == 0) and np.all((fin[:, 5] >> 30) & 1 == 1)
    # step kernel == rollout kernel: replay the recorded actions ply by ply
    for k in range(full.PLIES):
        r2, done, over = step(fs, st, decks, acts[:, k])
        assert not over
    assert np.array_equal(st, fin) and np.array_equal(r2, rew) and done.all()
    # a finished game is a fixed point of step
    before = st.copy()
    r3, done3, _ = step(fs, st, decks, np.zeros(n, np.uint8))
    assert np.array_equal(st, before) and np.array_equal(r3, rew) and done3.all()


def test_scoring_details(fs):
    """full_evaluate_kernel on hand-made piles: carte, denari, sette bello, primiera (calculate_primiera_score
    :160-172: 7 = 21, 6 = 18, ace = 16, 5 = 15, 4 = 14, 3 = 13, 2 = 12, face = 10; a missing suit = 0), scope, sweep."""
    cid = full.card_id
    mask = lambda cs: sum(1 << c for c in cs)
    caps0 = [cid(7, "denari"), cid(6, "coppe"), cid(1, "spade"), cid(10, "bastoni"), cid(2, "bastoni")]
    caps1 = [cid(1, "denari"), cid(2, "denari"), cid(3, "denari")]
    table = [cid(5, "coppe"), cid(9, "spade")]
    words = full.pack_full_state(table, [mask(caps0), mask(caps1)], [0, 0], 1, 0, False, [1, 0], 5, 36)
    st = np.array([words], dtype=np.uint32)
    rew, det = np.zeros((1, 2), np.float32), np.zeros((1, 8), np.int32)
    fs.host_full_evaluate(st.ctypes.data, rew.ctypes.data, det.ctypes.data, 1)
    # the table goes to the last capturer (player 1): cards 5 v 5 (nobody), denari 1 v 3 (player 1), sette bello player 0,
    # primiera 21 + 18 + 16 + 12 = 67 v 0 (player 1 holds no bastoni: no primiera), scope 1 v 0  ->  3 v 1
    assert det[0].tolist() == [5, 5, 1, 3, 67, 0, 3, 1]
    assert rew[0].tolist() == [1.0, -1.0]
    u = full.unpack_full_state(st[0])
    assert u["terminal"] and u["score_diff"] == 2 and u["table"] == table            # the table is not cleared
    assert u["cap_mask"] == [mask(caps0), mask(caps1 + table)]
