"""The PRODUCT's env kernels on the CPU: scopa_b200/csrc/ms_env.cu compiled for the host by tests/emu/ms_env_host.cpp.

deal_kernel restates CPython's `random.seed(int)` + `random.shuffle` (MiniDeck, /root/reference/src/envs/
mini_scopa_game.py:25-28) with a recomputed, windowed MT19937 seeding and a re-ordered shuffle loop -- the most delicate
integer code in the package.  CPython's own `random` IS the reference arithmetic and is available wherever the tests run,
so the kernel is compared with it directly (and with the reference-recorded fixture), seed by seed; likewise the
40-card deck (FullDeck, full_scopa_game.py:32-35), the slow paths, and the fused 8-ply rollout against the oracle.
tests/test_rules_host.py runs step / legal / capture / keys kernels of the same build through the rule fixtures."""
import ctypes as C
import os
import random
import sys

import numpy as np
import pytest

from conftest import load_golden_json
from oracle import ms_oracle as ora
from scopa_b200 import codec, full

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_build  # noqa: E402

vp = C.c_void_p


@pytest.fixture(scope="module")
def env():
    lib = C.CDLL(emu_build.build_env_host())
    lib.host_env_init()
    lib.host_deal.argtypes = [vp, C.c_longlong, vp, vp, vp, C.c_int]
    lib.host_deal_slow.argtypes = [vp, C.c_longlong, vp, vp]
    lib.host_full_deck.argtypes = [vp, C.c_longlong, vp, C.c_int, C.c_int]
    lib.host_step.argtypes = [vp, vp, vp, vp, C.c_longlong]
    lib.host_rollout.argtypes = [vp, vp, C.c_longlong, C.c_ulonglong, C.c_ulonglong, vp, vp, vp]
    return lib


def deal(env, seeds, zero_means_42=1):
    seeds = np.ascontiguousarray(seeds, np.int64)
    n = len(seeds)
    st, ho, deck = np.zeros((n, 4), np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint64)
    env.host_deal(seeds.ctypes.data, n, st.ctypes.data, ho.ctypes.data, deck.ctypes.data, zero_means_42)
    return st, ho, deck


def py_shuffle(seed, ncards):
    """the reference: random.seed(seed); random.shuffle(cards) over the deck in id order"""
    random.seed(seed)
    cards = list(range(ncards))
    random.shuffle(cards)
    return cards


EDGE_SEEDS = [0, 42, 1, -1, -42, 2**63 - 1, -(2**63) + 1, 2**32, 2**32 - 1, 2**32 + 1, 2**31, 2**31 - 1, 7 << 32, (1 << 62) + 5]


def test_deal_is_cpython_seed_and_shuffle(env):
    rng = np.random.default_rng(5)
    seeds = np.concatenate([np.array(EDGE_SEEDS, dtype=np.int64), rng.integers(0, 2**31, 12_000, dtype=np.int64),
                            rng.integers(-(2**62), 2**62, 8_000, dtype=np.int64)])
    state = random.getstate()
    try:
        st, ho, deck = deal(env, seeds, zero_means_42=0)
        for i, s in enumerate(seeds.tolist()):
            cards = py_shuffle(s, 16)
            assert codec.nibbles(int(deck[i]) & 0xFFFFFFFF, 8) + codec.nibbles(int(deck[i]) >> 32, 8) == cards, s
            assert int(ho[i]) == int(deck[i]) & 0xFFFFFFFF
            if i < 500:
                u = codec.unpack_state(st[i])
                assert u["hand_mask"] == [codec.mask_of(cards[:4]), codec.mask_of(cards[4:8])]
                assert u["table"] == [] and u["cap_mask"] == [0, 0] and u["scopas"] == [0, 0]
                assert u["step_count"] == 0 and u["cur"] == 0 and not u["terminal"] and u["max_steps"] == 8
    finally:
        random.setstate(state)
    # MiniScopaEnv.reset(seed): `seed or self.seed` -- 0 deals like 42 (mini_scopa_game.py:132)
    _, ho42, _ = deal(env, [0, 42], zero_means_42=1)
    assert ho42[0] == ho42[1] and ho42[0] != ho[0]


def test_deal_matches_reference_fixture_and_oracle(env):
    g = load_golden_json("deals.json")["decks"]
    seeds = [int(s) for s in g if -(2**63) <= int(s) < 2**63]
    _, ho, deck = deal(env, seeds, zero_means_42=0)
    for i, s in enumerate(seeds):
        assert codec.nibbles(int(deck[i]) & 0xFFFFFFFF, 8) + codec.nibbles(int(deck[i]) >> 32, 8) == g[str(s)], s
    rng = np.random.default_rng(8)
    bulk = np.concatenate([np.array(EDGE_SEEDS, dtype=np.int64), rng.integers(-(2**62), 2**62, 30_000, dtype=np.int64)])
    _, ho, _ = deal(env, bulk, zero_means_42=1)
    assert np.array_equal(np.stack([(ho >> (4 * i)) & 0xF for i in range(8)], 1).astype(np.int32), ora.batch_deal(bulk))


def test_deal_slow_path_agrees(env):
    rng = np.random.default_rng(6)
    seeds = np.concatenate([np.array([0, 42, -7, 2**40 + 3], dtype=np.int64), rng.integers(-(2**62), 2**62, 1500, dtype=np.int64)])
    st, ho, _ = deal(env, seeds, zero_means_42=1)
    st2, ho2 = np.zeros_like(st), np.zeros_like(ho)
    env.host_deal_slow(seeds.ctypes.data, len(seeds), st2.ctypes.data, ho2.ctypes.data)
    assert np.array_equal(st, st2) and np.array_equal(ho, ho2)


def test_full_deck_is_cpython_seed_and_shuffle(env):
    rng = np.random.default_rng(7)
    seeds = np.concatenate([np.array(EDGE_SEEDS, dtype=np.int64), np.arange(1, 3000, dtype=np.int64),
                            rng.integers(-(2**62), 2**62, 3000, dtype=np.int64)])
    out = {}
    for slow in (0, 1):
        d = np.zeros((len(seeds), 4), np.uint64)
        env.host_full_deck(seeds.ctypes.data, len(seeds), d.ctypes.data, 0, slow)
        out[slow] = d
    assert np.array_equal(out[0], out[1]), "fast and slow shuffle paths differ"
    state = random.getstate()
    try:
        for i, s in enumerate(seeds.tolist()):
            assert full.unpack_deck(out[0][i]) == py_shuffle(s, 40), s
    finally:
        random.setstate(state)
    g = load_golden_json("full_env_traces.json.gz")["decks"]
    gs = np.array([int(s) for s in g], dtype=np.int64)
    d = np.zeros((len(gs), 4), np.uint64)
    env.host_full_deck(gs.ctypes.data, len(gs), d.ctypes.data, 0, 0)
    for i, s in enumerate(gs.tolist()):
        assert full.unpack_deck(d[i]) == g[str(s)] == ora.full_deck(s), s
    d0 = np.zeros((2, 4), np.uint64)                         # FullScopaEnv.reset(0): `seed or self.seed`
    z = np.array([0, 42], dtype=np.int64)
    env.host_full_deck(z.ctypes.data, 2, d0.ctypes.data, 1, 0)
    assert np.array_equal(d0[0], d0[1])


def test_rollout_bit_exact_vs_oracle_and_step_replay(env):
    rng = np.random.default_rng(9)
    seeds = rng.integers(1, 2**40, 50_000, dtype=np.int64)
    n = len(seeds)
    st, ho, _ = deal(env, seeds)
    acts, rew, fin = np.zeros((n, 8), np.uint8), np.zeros((n, 2), np.float32), np.zeros((n, 4), np.uint32)
    env.host_rollout(st.ctypes.data, ho.ctypes.data, n, 0xC0FFEE, 17, acts.ctypes.data, rew.ctypes.data, fin.ctypes.data)
    o_act, o_rew, o_scopas, o_ncaps = ora.rollout_random(seeds, 0xC0FFEE, game_offset=17)
    assert np.array_equal(acts, o_act) and np.array_equal(rew, o_rew)
    assert np.array_equal((fin[:, 3] >> 4) & 0xF, o_scopas[:, 0]) and np.array_equal((fin[:, 3] >> 8) & 0xF, o_scopas[:, 1])
    assert not np.signbit(rew[rew == 0]).any()              # ties are +0.0 for both players
    # replaying the recorded actions through step_kernel reaches the same final state
    r, done = np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
    for k in range(8):
        a = np.ascontiguousarray(acts[:, k])
        env.host_step(st.ctypes.data, a.ctypes.data, r.ctypes.data, done.ctypes.data, n)
    assert np.array_equal(st, fin) and np.array_equal(r, rew) and done.all()
