#!/usr/bin/env python
"""Golden traces of the TPI (2-coordinator) OpenSpiel wrapper of team Miniscopa, recorded from the UNMODIFIED
reference (/root/reference/src/envs/openspiel_team_mini_scopa.py:6-264).  TEST INFRASTRUCTURE ONLY.
Writes tests/golden/team_tpi_traces.json.gz."""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference/src"]
import pyspiel  # noqa: E402
from envs import openspiel_team_mini_scopa  # noqa: E402,F401


def rec(state):
    return {"cp": state.current_player(), "term": state.is_terminal(), "legal": list(state.legal_actions()),
            "legal0": list(state.legal_actions(0)), "legal1": list(state.legal_actions(1)),
            "info0": state.information_state_string(0), "info1": state.information_state_string(1),
            "hist": state.history_str(), "rew": [float(x) for x in state.rewards()]}


def main():
    rng = random.Random(99)
    game = pyspiel.load_game("team_mini_scopa_tpi")
    traces = []
    for k in range(60):
        state = game.new_initial_state()
        acts, recs = [], [rec(state)]
        while not state.is_terminal() and len(acts) < 20:
            legal = state.legal_actions()
            a = rng.randrange(16) if rng.random() < 0.15 else rng.choice(legal)
            if k % 2:
                state = state.clone()
            state.apply_action(a)
            acts.append(a)
            recs.append(rec(state))
        traces.append({"actions": acts, "recs": recs})
    path = os.path.join(HERE, "..", "tests", "golden", "team_tpi_traces.json.gz")
    with open(path, "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
            f.write(json.dumps({"traces": traces}, separators=(",", ":"), sort_keys=True).encode())
    print("wrote", path, os.path.getsize(path), "bytes", "example:", traces[0]["recs"][3]["info0"], "|", traces[0]["recs"][3]["hist"])


if __name__ == "__main__":
    main()
