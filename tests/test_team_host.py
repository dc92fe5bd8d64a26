"""The PRODUCT's 2v2 team-Miniscopa code on the CPU: scopa_b200/csrc/ms_team.cu (tm_step, tm_finish, tm_legal_list and
the init / step / rollout kernels themselves) compiled for the host by tests/emu/ms_team_host.cpp, checked bit for bit
against the traces recorded from the unmodified reference (team_mini_scopa_game.py) and, for the fused rollout, against
the oracle.  Same fixtures and assertions as tests/test_gpu_team.py, which runs the device build of the same source."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import load_golden_json
from oracle import ms_oracle as ora
from scopa_b200 import codec
from scopa_b200.team import pack_team_state, team_hand_in_order, unpack_team_state

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_build  # noqa: E402

vp = C.c_void_p


@pytest.fixture(scope="module")
def team():
    lib = C.CDLL(emu_build.build_team_host())
    lib.host_team_init.argtypes = [vp, C.c_longlong, vp]
    lib.host_team_step.argtypes = [vp, vp, vp, vp, C.c_longlong]
    lib.host_team_rollout.argtypes = [vp, vp, C.c_longlong, C.c_ulonglong, C.c_ulonglong, vp, vp, vp]
    return lib


def init(team, decks):
    decks = np.ascontiguousarray(decks, np.uint64)
    st = np.zeros((len(decks), 8), np.uint32)
    team.host_team_init(decks.ctypes.data, len(decks), st.ctypes.data)
    return st


def step(team, st, actions):
    n = st.shape[0]
    rew, done = np.zeros((n, 4), np.float32), np.zeros(n, np.uint8)
    actions = np.ascontiguousarray(actions, np.uint8)
    team.host_team_step(st.ctypes.data, actions.ctypes.data, rew.ctypes.data, done.ctypes.data, n)
    return rew, done


def rollout(team, st, decks, philox_seed, game_offset):
    n = st.shape[0]
    acts, rew, fin = np.zeros((n, 16), np.uint8), np.zeros((n, 4), np.float32), np.zeros((n, 8), np.uint32)
    team.host_team_rollout(st.ctypes.data, decks.ctypes.data, n, philox_seed, game_offset, acts.ctypes.data, rew.ctypes.data,
                           fin.ctypes.data)
    return acts, rew, fin


def deck_word(cards16):
    return sum(int(c) << (4 * i) for i, c in enumerate(cards16))


def trace_decks(traces):
    """the shuffled deck of each trace = its four dealt hands in order (all 16 cards are dealt in the team game)"""
    return np.array([deck_word([c for h in t["snaps"][0]["hands"] for c in h]) for t in traces], dtype=np.uint64)


def test_team_init_matches_codec(team):
    rng = np.random.default_rng(0)
    perms = [rng.permutation(16) for _ in range(200)]
    st = init(team, [deck_word(p) for p in perms])
    for p, row in zip(perms, st):
        want = pack_team_state([codec.mask_of(p[4 * k:4 * k + 4].tolist()) for k in range(4)], [], [0] * 4, [0] * 4, 0, 0, False, None)
        assert tuple(int(x) for x in row) == tuple(want)


def test_team_kernels_follow_reference_traces(team):
    traces = load_golden_json("team_env_traces.json.gz")["traces"]
    decks = trace_decks(traces)
    st = init(team, decks)

    def check(k, rew):
        for i, t in enumerate(traces):
            snap, u = t["snaps"][k], unpack_team_state(st[i])
            assert u["table"] == snap["table"], (t["seed"], k)
            assert [team_hand_in_order(u["hand_mask"][p], decks[i], p) for p in range(4)] == snap["hands"]
            assert u["cap_mask"] == [codec.mask_of(c) for c in snap["caps"]], (t["seed"], k)
            assert u["scopas"] == snap["scopas"] and u["step_count"] == snap["step"]
            assert u["last_capture_team"] == snap["lct"] and f"player_{u['cur']}" == snap["agent"]
            assert [u["terminal"]] * 4 == snap["term"]
            if rew is not None:
                assert rew[i].tolist() == snap["rew"], (t["seed"], k)

    check(0, None)
    acts = np.array([t["actions"] for t in traces], dtype=np.uint8)
    for k in range(acts.shape[1]):
        rew, done = step(team, st, acts[:, k])
        check(k + 1, rew)
        assert done.tolist() == [int(t["snaps"][k + 1]["term"][0]) for t in traces]


def test_team_rollout_bit_exact_vs_oracle(team):
    seeds = np.random.default_rng(3).integers(1, 2**40, 20_000, dtype=np.int64)
    decks = np.zeros(len(seeds), np.uint64)
    for i, s in enumerate(seeds):
        e = ora.TeamEnv(int(s))
        decks[i] = deck_word([c for h in e.snapshot()["hands"] for c in h])
    st = init(team, decks)
    acts, rew, fin = rollout(team, st, decks, 77, 5)
    o_act, o_rew, o_sc = ora.team_rollout_random(seeds, 77, game_offset=5)
    assert np.array_equal(acts, o_act) and np.array_equal(rew, o_rew)
    assert np.array_equal(np.stack([(fin[:, 6] >> (4 * p)) & 0xF for p in range(4)], 1), o_sc)
    assert np.all(rew[:, 0] == rew[:, 1]) and np.all(rew[:, 2] == rew[:, 3]) and np.all(rew[:, 0] + rew[:, 2] == 0)
    # stepping the recorded actions one ply at a time reaches the same final state and the same rewards
    for k in range(16):
        r, done = step(team, st, acts[:, k])
    assert np.array_equal(st, fin) and np.array_equal(r, rew) and done.all()
    # a finished game is a fixed point of step, and its rewards stay what they were (no second sweep)
    before = st.copy()
    r2, done2 = step(team, st, np.zeros(len(seeds), np.uint8))
    assert np.array_equal(st, before) and np.array_equal(r2, rew) and done2.all()
