#!/usr/bin/env python
"""Golden traces for the 40-card Scopa env and its OpenSpiel wrapper, recorded from the UNMODIFIED reference
(/root/reference/src/envs/full_scopa_game.py:21-342, /root/reference/src/envs/openspiel_full_scopa.py:4-110).
TEST INFRASTRUCTURE ONLY; same import shims as oracle/gen_golden.py.  Writes tests/golden/full_env_traces.json.gz.

Card id = suit_idx * 10 + (rank - 1) = the reference's action id (full_scopa_game.py:262-266)."""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "stubs"), "/root/reference/src"]
import pyspiel  # noqa: E402
from envs import openspiel_full_scopa  # noqa: E402,F401  (registers "full_scopa")
from envs.full_scopa_game import FullDeck, FullScopaEnv  # noqa: E402

SUITS = FullDeck.suits


def cid(x):
    r, s = x if isinstance(x, tuple) else (x.rank, x.suit)
    return SUITS.index(s) * 10 + (r - 1)


def snap(env):
    st = env.get_state()
    return {"table": [cid(t) for t in st["table"]], "hands": [[cid(c) for c in h] for h in st["hands"]],
            "caps": [[cid(c) for c in h] for h in st["captures"]], "scopas": list(st["scopas"]),
            "deck": st["deck_remaining"], "round": st["round_number"], "last": st["last_capture"],
            "agent": st["agent_selection"], "step": st["step_count"],
            "rew": [float(st["rewards"][a]) for a in env.possible_agents],
            "term": [bool(st["terminations"][a]) for a in env.possible_agents]}


def env_trace(rng, seed, p_illegal, n_steps, lead_passes=0):
    env = FullScopaEnv(seed=42)
    env.reset(seed)
    snaps, acts = [snap(env)], []
    for k in range(n_steps):
        pl = env.game.players[env.agent_name_mapping[env.agent_selection]]
        if k < lead_passes:                      # a card the mover does not hold = silent pass (:269-271)
            a = next(x for x in range(40) if all(cid(c) != x for c in pl.hand))
        elif pl.hand and rng.random() >= p_illegal:
            a = cid(rng.choice(pl.hand))
        else:
            a = rng.randrange(40)
        acts.append(a)
        env.step(a)
        snaps.append(snap(env))
    return {"seed": seed, "actions": acts, "snaps": snaps}


def spiel_trace(rng):
    game = pyspiel.load_game("full_scopa")
    s = game.new_initial_state()
    rows = []

    def row(st):
        return {"cp": int(st.current_player()), "term": bool(st.is_terminal()), "legal": list(st.legal_actions()),
                "legal0": list(st.legal_actions(0)), "legal1": list(st.legal_actions(1)),
                "info0": st.information_state_string(0), "info1": st.information_state_string(1),
                "hist": st.history_str(), "rew": [float(x) for x in st.rewards()]}

    rows.append(row(s))
    acts = []
    while not s.is_terminal():
        a = rng.choice(s.legal_actions())
        # (no clone() here: the reference's FullScopaState.clone raises AttributeError -- inside
        # openspiel_full_scopa.py the name FullScopaGame is rebound to the pyspiel.Game subclass at :113)
        acts.append(a)
        s.apply_action(a)
        rows.append(row(s))
    return {"actions": acts, "rows": rows}


def main():
    rng = random.Random(4040)
    traces = []
    for k in range(160):
        seed = rng.choice([0, 42, rng.randrange(1, 2 ** 31), rng.randrange(1, 10 ** 6), rng.randrange(2 ** 32, 2 ** 62)])
        p_illegal = rng.choice([0.0, 0.0, 0.0, 0.1, 0.3])
        traces.append(env_trace(rng, seed, p_illegal, 40 if p_illegal == 0.0 else 70))
    # the safety limit (:286-290): nothing but passes ends the game at step 200 with the cards still in hand
    traces.append(env_trace(rng, 7, 0.0, 203, lead_passes=203))
    # 164 passes, then the 36 cards: the last card falls on step 200 and evaluate_game runs twice (:278-290)
    traces.append(env_trace(rng, 11, 0.0, 202, lead_passes=164))
    traces.append(env_trace(rng, 12, 0.0, 202, lead_passes=164))
    spiel = [spiel_trace(rng) for _ in range(12)]
    decks = {}
    for seed in [0, 1, 2, 42, 43, 12345, 2 ** 31 - 1, 2 ** 32, 2 ** 33 + 7, 10 ** 15 + 3, -5] + [rng.randrange(1, 2 ** 40) for _ in range(60)]:
        decks[str(seed)] = [cid(c) for c in FullDeck(seed).cards]
    path = os.path.join(HERE, "..", "tests", "golden", "full_env_traces.json.gz")
    with open(path, "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
            f.write(json.dumps({"traces": traces, "spiel": spiel, "decks": decks}, separators=(",", ":"), sort_keys=True).encode())
    longest = max(max(len(s["table"]) for s in t["snaps"]) for t in traces)
    print("wrote", path, os.path.getsize(path), "bytes;", len(traces), "env traces,", len(spiel), "wrapper traces,",
          len(decks), "decks; longest table", longest)


if __name__ == "__main__":
    main()
