"""The PRODUCT's rule code on the CPU: scopa_b200/csrc/ms_state.cuh (capture_mask, step, legal_list, infoset_key,
reward0 -- the __device__ functions that every env and solver kernel calls) compiled for the host by
tests/emu/ms_state_host.cpp and checked bit for bit against the fixtures recorded from the unmodified reference
(the same fixtures and assertions as tests/test_gpu_env.py, which runs the device build of the same header)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import load_golden_json
from scopa_b200 import codec

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu_build  # noqa: E402

vp = C.c_void_p


@pytest.fixture(scope="module", params=["rule header + restated loops", "ms_env.cu kernels"])
def rules(request):
    """the same exports from two builds: tests/emu/ms_state_host.cpp (ms_state.cuh + the kernels' loops restated) and
    tests/emu/ms_env_host.cpp (step_kernel / legal_kernel / capture_kernel / keys_kernel of ms_env.cu themselves)"""
    lib = C.CDLL(emu_build.build_state_host() if request.param.startswith("rule") else emu_build.build_env_host())
    lib.host_step.argtypes = [vp, vp, vp, vp, C.c_longlong]
    lib.host_legal.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, C.c_longlong]
    lib.host_capture.argtypes = [vp, vp, vp, C.c_longlong]
    lib.host_keys.argtypes = [vp, C.c_int, vp, C.c_longlong]
    return lib


def deal_state(cards, max_steps=8):
    """packed root state + hand_order word of a deal (first 8 cards of the shuffled deck, 4 + 4, empty table)"""
    words = codec.pack_state([codec.mask_of(cards[:4]), codec.mask_of(cards[4:8])], [], [0, 0], [0, 0], 0, 0, False, max_steps)
    return np.array(words, dtype=np.uint32), codec.pack_nibbles(cards[:8])


def step(rules, st, actions):
    n = st.shape[0]
    rew, done = np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
    actions = np.ascontiguousarray(actions, np.uint8)
    rules.host_step(st.ctypes.data, actions.ctypes.data, rew.ctypes.data, done.ctypes.data, n)
    return rew, done


def legal(rules, st, ho, player):
    n = st.shape[0]
    mask, ordered = np.zeros(n, np.uint16), np.zeros((n, 4), np.uint8)
    count, cap = np.zeros(n, np.uint8), np.zeros((n, 4), np.uint8)
    rules.host_legal(st.ctypes.data, ho.ctypes.data, player, mask.ctypes.data, ordered.ctypes.data, count.ctypes.data,
                     cap.ctypes.data, n)
    return mask, ordered, count, cap


def keys(rules, st, player):
    out = np.zeros(st.shape[0], np.uint64)
    rules.host_keys(st.ctypes.data, player, out.ctypes.data, st.shape[0])
    return out


def test_capture_cases_from_reference(rules):
    cases = load_golden_json("capture_cases.json")
    n = len(cases)
    st, cards = np.zeros((n, 4), np.uint32), np.zeros(n, np.uint8)
    for i, (table, played, isin, mask) in enumerate(cases):
        st[i] = codec.pack_state([1 << played, 0], table, [0, 0], [0, 0], 0, 0, False, 8)
        cards[i] = played
    got = np.zeros(n, np.uint8)
    rules.host_capture(st.ctypes.data, cards.ctypes.data, got.ctypes.data, n)
    assert np.array_equal(got, np.array([c[3] for c in cases], np.uint8))
    assert (np.array([c[2] for c in cases]) == (got != 0)).all()
    # legal_list + capture_mask as legal_kernel combines them: the single card in hand, the same capture
    _, ordered, count, cap = legal(rules, st, cards.astype(np.uint32), 0)
    assert np.array_equal(cap[:, 0], got) and np.array_equal(ordered[:, 0], cards) and (count == 1).all()


def test_exhaustive_seed42_tree_bit_exact(rules):
    nodes = load_golden_json("env_tree_seed42.json.gz")["nodes"]
    deck = load_golden_json("deals.json")["decks"]["42"]
    root, ho = deal_state(deck)
    n = len(nodes)
    st = np.tile(root, (n, 1))
    rew = np.zeros((n, 2), np.float32)
    for t in range(8):
        rows = np.array([i for i, nd in enumerate(nodes) if len(nd["h"]) > t], dtype=np.int64)
        if rows.size == 0:
            break
        sub = np.ascontiguousarray(st[rows])
        r, _ = step(rules, sub, [nodes[i]["h"][t] for i in rows])
        st[rows], rew[rows] = sub, r
    hos = np.full(n, ho, np.uint32)
    res = {pl: legal(rules, st, hos, pl)[:3] + (keys(rules, st, pl),) for pl in (-1, 0, 1)}
    for i, nd in enumerate(nodes):
        u = codec.unpack_state(st[i])
        assert u["terminal"] == nd["term"] and u["table"] == nd["table"], nd["h"]
        assert [codec.hand_in_order(u["hand_mask"][p], ho, p) for p in range(2)] == nd["hands"]
        assert u["cap_mask"] == [codec.mask_of(c) for c in nd["caps"]]
        assert u["scopas"] == nd["scopas"] and u["step_count"] == nd["step"]
        assert f"player_{u['cur']}" == nd["agent"] and rew[i].tolist() == nd["rew"]
        for pl, lk, ik in ((-1, "legal", "info"), (0, "legal0", "info0"), (1, "legal1", "info1")):
            mask, ordered, count, ks = res[pl]
            want = nd[lk]
            assert int(count[i]) == len(want) and ordered[i, :len(want)].tolist() == want
            assert int(mask[i]) == codec.mask_of(want)
            assert codec.key_to_string(ks[i], ho) == nd[ik], (nd["h"], pl)
    live = np.array([not nd["term"] for nd in nodes])
    assert len(np.unique(res[-1][3][live])) == len({nd["info"] for nd in nodes if not nd["term"]}) == 738


def test_random_traces_with_illegal_actions_and_dead_steps(rules):
    traces = [t for t in load_golden_json("env_random_traces.json.gz")["traces"] if t["kind"] == "env"]
    roots = []
    for t in traces:                                   # the fixture's first snapshot is the deal (hands in deal order)
        h = t["snaps"][0]["hands"]
        roots.append(deal_state(h[0] + h[1]))
    st = np.stack([r[0] for r in roots])
    ho = [r[1] for r in roots]

    def check(k, rew):
        for i, t in enumerate(traces):
            snap, u = t["snaps"][k], codec.unpack_state(st[i])
            assert u["table"] == snap["table"], (t["seed"], k)
            assert [codec.hand_in_order(u["hand_mask"][p], ho[i], p) for p in range(2)] == snap["hands"]
            assert u["cap_mask"] == [codec.mask_of(c) for c in snap["caps"]]
            assert u["scopas"] == snap["scopas"] and u["step_count"] == snap["step"]
            assert f"player_{u['cur']}" == snap["agent"] and [u["terminal"]] * 2 == snap["term"]
            if rew is not None:
                assert rew[i].tolist() == snap["rew"]

    check(0, None)
    acts = np.array([t["actions"] for t in traces], dtype=np.uint8)
    for k in range(acts.shape[1]):
        r, _ = step(rules, st, acts[:, k])
        check(k + 1, r)


def test_clone_semantics_max_steps_16(rules):
    traces = [t for t in load_golden_json("env_random_traces.json.gz")["traces"] if t["kind"] == "spiel_clone"]
    deck = load_golden_json("deals.json")["decks"]["42"]
    root, ho = deal_state(deck, max_steps=16)
    for t in traces:
        st = root.copy().reshape(1, 4)
        hos = np.array([ho], np.uint32)
        for k, a in enumerate(t["actions"]):
            r, _ = step(rules, st, [a])
            rec, u = t["recs"][k], codec.unpack_state(st[0])
            _, ordered, count, _ = legal(rules, st, hos, -1)
            assert u["terminal"] == rec["term"] and u["step_count"] == rec["step"] and u["table"] == rec["table"]
            assert ordered[0, :count[0]].tolist() == rec["legal"], (t["actions"], k)
            assert codec.key_to_string(keys(rules, st, -1)[0], ho) == rec["info"]
            assert r[0].tolist() == rec["rew"]
