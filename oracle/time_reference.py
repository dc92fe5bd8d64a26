#!/usr/bin/env python
"""oracle/time_reference.py -- time the UNMODIFIED reference (oracle/_ref, see make_ref.py) on one core.

TEST / BENCH INFRASTRUCTURE ONLY.  One process = one single-threaded Python interpreter running the reference's own
loops, imported behind oracle/stubs:
    mccfr N   MCCFRTrainer(game).iteration() x N           src/algorithms/mc_cfr.py:88-92     (172 updates, 703 visits each)
    cfr N     CFRTrainer(game).train(steps=N)              src/algorithms/vanilla_cfr.py:105-120
    env N     MiniScopaEnv.reset(seed=g) + 8 random legal steps, N games   src/envs/mini_scopa_game.py:131-167

    python oracle/time_reference.py mccfr 100         -> one JSON line
    python oracle/time_reference.py serve             -> reads "<what> <n>" lines on stdin, answers one JSON line each
                                                         (imports and warm-up happen before the first answer "ready")
"""
import json
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")
sys.path.insert(0, os.path.join(HERE, "stubs"))
if os.path.isfile(os.path.join(REF, "src", "algorithms", "mc_cfr.refc")):
    sys.path.insert(0, HERE)
    import ref_import
    ref_import.install([os.path.join(REF, "src"), REF])
    REF_KIND = "oracle/_ref (byte-compiled reference)"
elif os.path.isdir("/root/reference/src"):
    sys.path[1:1] = ["/root/reference/src", "/root/reference"]       # authoring container: the sources themselves
    REF_KIND = "/root/reference"
else:
    raise SystemExit("time_reference.py: neither oracle/_ref nor /root/reference is present")

import numpy as np  # noqa: E402
import pyspiel  # noqa: E402  (shim unless the real package is installed)
from envs import openspiel_mini_scopa  # noqa: E402,F401  (registers "mini_scopa")
from envs.mini_scopa_game import MiniDeck, MiniScopaEnv  # noqa: E402
import algorithms.vanilla_cfr as ref_cfr  # noqa: E402
import algorithms.mc_cfr as ref_mccfr  # noqa: E402

ref_cfr.tqdm = lambda it, **k: it
ref_mccfr.tqdm = lambda it, **k: it
GAME = pyspiel.load_game("mini_scopa")
TRAINER = None


def run(what, n):
    global TRAINER
    if what == "mccfr":
        if TRAINER is None:
            np.random.seed(os.getpid() & 0x7FFFFFFF)
            TRAINER = ref_mccfr.MCCFRTrainer(GAME)
            TRAINER.iteration()                      # warm-up: creates most InfoNodes
        t0 = time.perf_counter()
        for _ in range(n):
            TRAINER.iteration()
        dt = time.perf_counter() - t0
        return {"what": what, "n": n, "seconds": dt, "ms_per_iteration": dt / n * 1e3, "updates": 172 * n, "visits": 703 * n,
                "updates_per_sec": 172 * n / dt, "visits_per_sec": 703 * n / dt}
    if what == "cfr":
        tr = ref_cfr.CFRTrainer(GAME)
        t0 = time.perf_counter()
        tr.train(steps=n)
        dt = time.perf_counter() - t0
        return {"what": what, "n": n, "seconds": dt, "ms_per_iteration": dt / n * 1e3,
                "node_visits_per_sec": 2 * 2229 * n / dt}
    if what == "env":
        rng = random.Random(5)
        env = MiniScopaEnv()
        card_id = {(r, s): si * 4 + ci for si, s in enumerate(MiniDeck.suits) for ci, r in enumerate(MiniDeck.ranks[s])}
        steps = 0
        t0 = time.perf_counter()
        for g in range(1, n + 1):
            env.reset(seed=g)
            while not all(env.terminations.values()):
                hand = env.game.players[env.agent_name_mapping[env.agent_selection]].hand
                c = hand[rng.randrange(len(hand))]           # uniform-random legal action
                env.step(card_id[(c.rank, c.suit)])
                steps += 1
        dt = time.perf_counter() - t0
        return {"what": what, "n": n, "seconds": dt, "steps": steps, "steps_per_sec": steps / dt}
    raise SystemExit(f"unknown workload {what!r}")


def main():
    if len(sys.argv) >= 2 and sys.argv[1] == "serve":
        run("mccfr", 1)
        run("env", 20)
        print(json.dumps({"ready": True, "pid": os.getpid(), "reference": REF_KIND, "module_file": ref_mccfr.__file__}), flush=True)
        for line in sys.stdin:
            parts = line.split()
            if not parts or parts[0] == "quit":
                break
            print(json.dumps(run(parts[0], int(parts[1]))), flush=True)
        return
    print(json.dumps(run(sys.argv[1], int(sys.argv[2]))), flush=True)


if __name__ == "__main__":
    main()
