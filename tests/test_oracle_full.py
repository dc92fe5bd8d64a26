"""The oracle's 40-card Scopa restatement (oracle/ms_oracle.c "40-card Scopa") against traces recorded from the
UNMODIFIED reference (tests/golden/full_env_traces.json.gz, generator oracle/gen_golden_full.py): 71 shuffled decks,
163 env traces (8 837 steps with illegal actions, dead steps, the 200-step safety limit and the step-200 double
evaluation).  CPU only."""
import numpy as np

from conftest import load_golden_json
from oracle import ms_oracle as ora


def test_full_decks_match_reference():
    for seed, deck in load_golden_json("full_env_traces.json.gz")["decks"].items():
        assert ora.full_deck(int(seed)) == deck, seed


def test_full_env_follows_reference_traces():
    traces = load_golden_json("full_env_traces.json.gz")["traces"]
    steps = 0
    for t in traces:
        e = ora.FullEnv(42)
        e.reset(t["seed"])
        assert e.snapshot() == t["snaps"][0], t["seed"]
        for k, a in enumerate(t["actions"]):
            e.step(a)
            assert e.snapshot() == t["snaps"][k + 1], (t["seed"], k, a)
            steps += 1
    assert steps > 8000
    # the edge cases are really in the fixture
    assert traces[-3]["snaps"][-1]["hands"] != [[], []] and traces[-3]["snaps"][-1]["term"] == [True, True]
    assert sum(len(c) for c in traces[-1]["snaps"][-1]["caps"]) > 40


def test_full_rollout_shapes_and_bounds():
    a, r, sc, nc, mt = ora.full_rollout_random(np.arange(1, 20001), 5)
    assert mt.max() <= 16 and (nc.sum(1) == 40).mean() > 0.5 and np.all(r[:, 0] == -r[:, 1])
