"""The CPU oracle (oracle/ms_oracle.c) pinned against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py): deals, capture resolution, the exhaustive seed-42 game tree and random
traces with illegal actions.  Bit-exact (integer / string work)."""
import numpy as np
import pytest

from conftest import load_golden_json
from oracle import ms_oracle as ora


def test_deck_matches_reference_shuffle():
    g = load_golden_json("deals.json")
    for seed, cards in g["decks"].items():
        assert ora.deck(int(seed)) == cards, seed
    # documented anchors: seed 42 -> P0 [9f,6p,5f,7f] = ids [7,9,5,6]
    assert ora.deck(42)[:8] == [7, 9, 5, 6, 14, 10, 12, 8]


def test_env_reset_seed_zero_and_none_mean_42():
    g = load_golden_json("deals.json")
    e = ora.Env()
    e.reset(0)
    assert e.snapshot()["hands"] == g["env_reset_0_hands"]
    e.reset(None)
    assert e.snapshot()["hands"] == g["env_reset_none_hands"]
    assert g["env_reset_0_hands"] == [ora.deck(42)[:4], ora.deck(42)[4:8]]


def test_card_in_table_cases():
    cases = load_golden_json("capture_cases.json")
    n_subset = 0
    for table, played, isin, mask in cases:
        got_in, pos = ora.card_in_table(table, played)
        assert int(got_in) == isin
        assert sum(1 << p for p in pos) == mask
        n_subset += len(pos) > 1
    assert n_subset > 200  # the fixture exercises real subset-sum captures


def test_exhaustive_seed42_tree():
    nodes = load_golden_json("env_tree_seed42.json.gz")["nodes"]
    assert len(nodes) == 2229
    fields = ["cp", "term", "legal", "legal0", "legal1", "info", "info0", "info1", "hist", "rew", "hands", "caps",
              "scopas", "table", "step", "agent"]
    for nd in nodes:
        s = ora.State(42)
        for a in nd["h"]:
            s = s.clone()
            s.apply_action(a)
        rec = s.record()
        for f in fields:
            assert rec[f] == nd[f], (nd["h"], f, rec[f], nd[f])


def test_random_traces_with_illegal_actions():
    traces = load_golden_json("env_random_traces.json.gz")["traces"]
    n_env = n_sp = 0
    for tr in traces:
        if tr["kind"] == "env":
            e = ora.Env(42)
            e.reset(tr["seed"])
            assert e.snapshot() == tr["snaps"][0]
            for a, snap in zip(tr["actions"], tr["snaps"][1:]):
                e.step(a)
                assert e.snapshot() == snap, (tr["seed"], tr["actions"])
            n_env += 1
        else:
            s = ora.State(42)
            for a, rec in zip(tr["actions"], tr["recs"]):
                s = s.clone()
                s.apply_action(a)
                got = s.record()
                for f, v in rec.items():
                    if f != "h":
                        assert got[f] == v, (tr["actions"], f)
            n_sp += 1
    assert n_env >= 300 and n_sp >= 50


def test_batch_deal_and_rollout_consistency():
    seeds = np.array([0, 1, 42, 43, 12345, 2**33 + 7], dtype=np.int64)
    hands = ora.batch_deal(seeds)
    assert hands[0].tolist() == ora.deck(42)[:8]
    assert hands[1].tolist() == ora.deck(1)[:8]
    actions, rewards, scopas, ncaps = ora.rollout_random(seeds, 7)
    # replay the recorded actions through the scalar env: same rewards
    for g, seed in enumerate(seeds):
        e = ora.Env(42)
        e.reset(int(seed))
        for a in actions[g]:
            e.step(int(a))
        snap = e.snapshot()
        assert snap["term"] == [True, True]
        assert np.allclose(snap["rew"], rewards[g])
        assert rewards[g].sum() == 0


def test_team_env_traces_from_reference():
    """2v2 team Miniscopa (a "next" row): the oracle against 300 traces recorded from the unmodified reference,
    incl. illegal actions, dead steps and the last-capturer sweep."""
    traces = load_golden_json("team_env_traces.json.gz")["traces"]
    for tr in traces:
        e = ora.TeamEnv(42)
        e.reset(tr["seed"])
        assert e.snapshot() == tr["snaps"][0]
        for a, snap in zip(tr["actions"], tr["snaps"][1:]):
            e.step(a)
            assert e.snapshot() == snap, (tr["seed"], tr["actions"])
