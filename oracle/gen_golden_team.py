#!/usr/bin/env python
"""Golden traces for the 2v2 team Miniscopa env, recorded from the UNMODIFIED reference
(/root/reference/src/envs/team_mini_scopa_game.py:44-243).  TEST INFRASTRUCTURE ONLY; same import shims as
oracle/gen_golden.py.  Writes tests/golden/team_env_traces.json.gz."""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference/src"]
from envs.team_mini_scopa_game import MiniDeck, TeamMiniScopaEnv  # noqa: E402

SUITS = MiniDeck.suits
CARD_ID = {(r, s): si * 4 + ci for si, s in enumerate(SUITS) for ci, r in enumerate(MiniDeck.ranks[s])}


def cid(x):
    return CARD_ID[x if isinstance(x, tuple) else (x.rank, x.suit)]


def snap(env):
    st = env.get_state()
    return {"table": [cid(t) for t in st["table"]], "hands": [[cid(c) for c in h] for h in st["hands"]],
            "caps": [[cid(c) for c in h] for h in st["captures"]], "scopas": list(st["scopas"]),
            "lct": st["last_capture_team"], "agent": st["agent_selection"], "step": st["step_count"],
            "rew": [float(st["rewards"][a]) for a in env.possible_agents],
            "term": [bool(st["terminations"][a]) for a in env.possible_agents]}


def main():
    rng = random.Random(2024)
    traces = []
    for k in range(300):
        seed = rng.choice([0, 42, rng.randrange(1, 2 ** 31), rng.randrange(1, 10 ** 6), rng.randrange(2 ** 32, 2 ** 62)])
        p_illegal = rng.choice([0.0, 0.0, 0.0, 0.15, 0.4])
        env = TeamMiniScopaEnv(seed=42)
        env.reset(seed)
        snaps, acts = [snap(env)], []
        for _ in range(18):
            pl = env.game.players[env.agent_name_mapping[env.agent_selection]]
            a = cid(rng.choice(pl.hand)) if (pl.hand and rng.random() >= p_illegal) else rng.randrange(16)
            acts.append(a)
            env.step(a)
            snaps.append(snap(env))
        traces.append({"seed": seed, "actions": acts, "snaps": snaps})
    path = os.path.join(HERE, "..", "tests", "golden", "team_env_traces.json.gz")
    with open(path, "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
            f.write(json.dumps({"traces": traces}, separators=(",", ":"), sort_keys=True).encode())
    n_sweep = sum(1 for t in traces if t["snaps"][-1]["table"] and t["snaps"][-1]["lct"] is not None)
    print("wrote", path, os.path.getsize(path), "bytes;", len(traces), "traces;", n_sweep, "end with cards left on the table")


if __name__ == "__main__":
    main()
