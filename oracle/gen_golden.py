#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the authoring container (where /root/reference exists);
the fixtures it writes are committed so that nothing on the GPU box needs /root/reference.

The reference has no tests or known-answer vectors of its own, so every
pin below is manufactured from the reference's own code, imported from /root/reference/src
behind the import shims in oracle/stubs/ (base classes + registry only, no game logic).

    python oracle/gen_golden.py            # rewrites tests/golden/*

Fixtures (all produced by reference code paths, cited per block):
  deals.json            MiniDeck(seed).cards                 src/envs/mini_scopa_game.py:15-28
  capture_cases.json    MiniScopaGame.card_in_table          src/envs/mini_scopa_game.py:66-91
  env_tree_seed42.json.gz   exhaustive seed-42 game tree through MiniScopaState
                                                             src/envs/openspiel_mini_scopa.py:8-115
  env_random_traces.json.gz MiniScopaEnv.reset/step incl. illegal actions
                                                             src/envs/mini_scopa_game.py:131-194
  cfr_seed42.npz        CFRTrainer.train                     src/algorithms/vanilla_cfr.py:56-120
  mccfr_npseed*.npz     MCCFRTrainer.iteration               src/algorithms/mc_cfr.py:37-92
  sdcfr_seed0.npz       DeepCFR._state_to_features/_external_sampling_cfr
                                                             src/algorithms/deep_cfr/deep_cfr.py:213-365
  policies_eval.json    exploitability (restated BR, oracle/ms_exploit.py) of reference policies
"""
import gzip
import hashlib
import io
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "..", "tests", "golden")

sys.path[:0] = [
    os.path.join(HERE, "stubs"),
    os.path.join(REF, "src"),
    REF,
    os.path.join(REF, "src", "algorithms", "deep_cfr"),
    HERE,
]

import pyspiel  # noqa: E402  (shim)
from envs import openspiel_mini_scopa  # noqa: E402,F401  (registers "mini_scopa")
from envs.mini_scopa_game import Card, MiniDeck, MiniScopaEnv, MiniScopaGame  # noqa: E402
import algorithms.vanilla_cfr as ref_cfr  # noqa: E402
import algorithms.mc_cfr as ref_mccfr  # noqa: E402

ref_cfr.tqdm = lambda it, **k: it
ref_mccfr.tqdm = lambda it, **k: it

SUITS = MiniDeck.suits
CARD_ID = {(r, s): si * 4 + ci for si, s in enumerate(SUITS) for ci, r in enumerate(MiniDeck.ranks[s])}


def cid(card):
    """Card object or (rank, suit) tuple -> action id 0..15 (mini_scopa_game.py:149-153)."""
    if isinstance(card, tuple):
        return CARD_ID[card]
    return CARD_ID[(card.rank, card.suit)]


def dump_json(name, obj, gz=False):
    path = os.path.join(OUT, name)
    data = json.dumps(obj, separators=(",", ":"), sort_keys=True)
    if gz:
        with open(path, "wb") as raw:
            with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
                f.write(data.encode())
    else:
        with open(path, "w") as f:
            f.write(data)
    print(f"wrote {name}: {os.path.getsize(path)} bytes")


# --------------------------------------------------------------------------- deals
def gen_deals():
    seeds = [1, 2, 3, 7, 42, 43, 12345, 99999, 2**31 - 1, 2**31, 2**32 - 1, 2**32, 2**33 + 7,
             2**40 + 12345, 2**62 + 1, 2**63 - 1, -1, -42, -(2**40)]
    rng = random.Random(20261018)
    seeds += [rng.randrange(0, 2**31) for _ in range(300)]
    seeds += [rng.randrange(2**32, 2**63) for _ in range(60)]
    seeds += list(range(100, 228))
    out = {}
    for s in seeds:
        out[str(s)] = [cid(c) for c in MiniDeck(s).cards]
    # env-level semantics: `seed or self.seed` (mini_scopa_game.py:132) -> reset(0) == reset(42)
    env = MiniScopaEnv()
    env.reset(0)
    env0 = [[cid(c) for c in p.hand] for p in env.game.players]
    env.reset(None)
    envn = [[cid(c) for c in p.hand] for p in env.game.players]
    dump_json("deals.json", {"decks": out, "env_reset_0_hands": env0, "env_reset_none_hands": envn})


# --------------------------------------------------------------------------- capture cases
def gen_capture_cases():
    rng = random.Random(7)
    cards = [(r, s) for s in SUITS for r in MiniDeck.ranks[s]]
    cases = []
    g = MiniScopaGame()
    for _ in range(6000):
        n = rng.choice([0, 1, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 8])
        pick = rng.sample(cards, n + 1)
        table, played = pick[:n], pick[n]
        g.table = [Card(r, s) for r, s in table]
        isin, captured = g.card_in_table(Card(*played))
        pos = [next(i for i, c in enumerate(g.table) if c is cc) for cc in captured]
        mask = sum(1 << p for p in pos)
        cases.append([[cid(t) for t in table], cid(played), int(isin), mask])
    dump_json("capture_cases.json", cases)


# --------------------------------------------------------------------------- env tree (seed 42)
def state_record(state, history):
    env = state.env
    g = env.game
    rec = {
        "h": list(history),
        "cp": state.current_player(),
        "term": bool(state.is_terminal()),
        "legal": list(state.legal_actions()),
        "legal0": list(state.legal_actions(0)),
        "legal1": list(state.legal_actions(1)),
        "info": state.information_state_string(),
        "info0": state.information_state_string(0),
        "info1": state.information_state_string(1),
        "hist": state.history_str(),
        "rew": [float(x) for x in state.rewards()],
        "hands": [[cid(c) for c in p.hand] for p in g.players],
        "caps": [[cid(c) for c in p.captures] for p in g.players],
        "scopas": [p.scopas for p in g.players],
        "table": [cid(c) for c in g.table],
        "step": env.step_count,
        "agent": env.agent_selection,
    }
    return rec


def gen_env_tree():
    game = pyspiel.load_game("mini_scopa")
    nodes = []

    def rec(state, history):
        nodes.append(state_record(state, history))
        if state.is_terminal():
            return
        for a in state.legal_actions():
            c = state.clone()
            c.apply_action(a)
            rec(c, history + [a])

    rec(game.new_initial_state(), [])
    n_term = sum(n["term"] for n in nodes)
    infos = {n["info"] for n in nodes if not n["term"]}
    print(f"seed-42 tree: {len(nodes)} nodes, {n_term} terminals, {len(infos)} infosets")
    dump_json("env_tree_seed42.json.gz", {"nodes": nodes}, gz=True)


# --------------------------------------------------------------------------- random env traces
def env_snapshot(env):
    st = env.get_state()
    return {
        "table": [cid(t) for t in st["table"]],
        "hands": [[cid(c) for c in h] for h in st["hands"]],
        "caps": [[cid(c) for c in h] for h in st["captures"]],
        "scopas": list(st["scopas"]),
        "agent": st["agent_selection"],
        "step": st["step_count"],
        "rew": [float(st["rewards"][a]) for a in env.possible_agents],
        "term": [bool(st["terminations"][a]) for a in env.possible_agents],
    }


def gen_env_random_traces():
    rng = random.Random(11)
    traces = []
    # (a) raw MiniScopaEnv: random actions, a share of them illegal (silent pass, :155-157)
    for k in range(400):
        seed = rng.choice([0, 42, rng.randrange(1, 2**31), rng.randrange(1, 10**6)])
        p_illegal = rng.choice([0.0, 0.0, 0.2, 0.5])
        env = MiniScopaEnv(seed=42)
        env.reset(seed)
        steps = [env_snapshot(env)]
        acts = []
        for _ in range(10):  # two past the end: dead steps are no-ops (:141-143)
            player = env.game.players[env.agent_name_mapping[env.agent_selection]]
            if player.hand and rng.random() >= p_illegal:
                a = cid(rng.choice(player.hand))
            else:
                a = rng.randrange(16)
            acts.append(a)
            env.step(a)
            steps.append(env_snapshot(env))
        traces.append({"kind": "env", "seed": seed, "actions": acts, "snaps": steps})
    # (b) OpenSpiel wrapper with clone(): max_steps becomes 16 on clones (openspiel_mini_scopa.py:108)
    game = pyspiel.load_game("mini_scopa")
    for k in range(60):
        state = game.new_initial_state()
        acts, recs = [], []
        for _ in range(18):
            if state.is_terminal():
                break
            legal = state.legal_actions()
            if rng.random() < 0.3:
                a = rng.randrange(16)
            else:
                a = rng.choice(legal)
            state = state.clone()
            state.apply_action(a)
            acts.append(a)
            recs.append(state_record(state, acts))
        traces.append({"kind": "spiel_clone", "seed": 42, "actions": acts, "recs": recs})
    dump_json("env_random_traces.json.gz", {"traces": traces}, gz=True)


# --------------------------------------------------------------------------- CFR
def table_arrays(info_map, key_fn=lambda k: k):
    keys = list(info_map.keys())
    n = len(keys)
    reg = np.zeros((n, 4))
    strat = np.zeros((n, 4))
    nl = np.zeros(n, dtype=np.int8)
    legal = np.full((n, 4), -1, dtype=np.int8)
    for i, k in enumerate(keys):
        node = info_map[k]
        m = node.legal_actions.size
        nl[i] = m
        legal[i, :m] = node.legal_actions
        reg[i, :m] = node.regret_sum
        strat[i, :m] = node.strategy_sum
    return [key_fn(k) for k in keys], reg, strat, nl, legal


def table_sha(keys, reg, strat, nl):
    h = hashlib.sha256()
    for i in np.argsort(np.array(keys, dtype=object)):
        h.update(keys[i].encode())
        h.update(reg[i, :nl[i]].tobytes())
        h.update(strat[i, :nl[i]].tobytes())
    return h.hexdigest()


def gen_cfr():
    game = pyspiel.load_game("mini_scopa")
    tr = ref_cfr.CFRTrainer(game)
    snaps = {}
    it = 0
    for target in (1, 2, 5, 20):
        tr.train(target - it)
        it = target
        keys, reg, strat, nl, legal = table_arrays(tr.info_set_map)
        snaps[f"reg_{target}"] = reg
        snaps[f"strat_{target}"] = strat
        print(f"CFR iter {target}: {len(keys)} infosets sha={table_sha(keys, reg, strat, nl)}")
    snaps["keys"] = np.array(keys)
    snaps["nlegal"] = nl
    snaps["legal"] = legal
    np.savez_compressed(os.path.join(OUT, "cfr_seed42.npz"), **snaps)
    print("wrote cfr_seed42.npz", os.path.getsize(os.path.join(OUT, "cfr_seed42.npz")))
    return tr


# --------------------------------------------------------------------------- MCCFR
def gen_mccfr():
    game = pyspiel.load_game("mini_scopa")
    trainers = {}
    for npseed in (0, 1):
        np.random.seed(npseed)
        tr = ref_mccfr.MCCFRTrainer(game)
        snaps = {}
        it = 0
        for target in (1, 5, 20, 100):
            for _ in range(target - it):
                tr.iteration()
            it = target
            keys, reg, strat, nl, legal = table_arrays(tr.info_sets, key_fn=lambda k: f"{k[0]}|{k[1]}")
            snaps[f"keys_{target}"] = np.array(keys)
            snaps[f"reg_{target}"] = reg
            snaps[f"strat_{target}"] = strat
            snaps[f"nlegal_{target}"] = nl
            snaps[f"legal_{target}"] = legal
        np.savez_compressed(os.path.join(OUT, f"mccfr_npseed{npseed}.npz"), **snaps)
        print(f"wrote mccfr_npseed{npseed}.npz: {len(keys)} infosets after {it} iterations")
        trainers[npseed] = tr
    return trainers


# --------------------------------------------------------------------------- exploitability
def gen_policies_eval(cfr_trainer_unused=None):
    import ms_exploit
    game = pyspiel.load_game("mini_scopa")
    out = {}
    uni = ref_cfr.RandomPolicy(game)
    out["uniform"] = ms_exploit.exploitability(game, uni)
    tr = ref_cfr.CFRTrainer(game)
    it = 0
    out["cfr"] = {}
    for target in (1, 2, 5, 10, 20, 50):
        tr.train(target - it)
        it = target
        out["cfr"][str(target)] = ms_exploit.exploitability(game, tr.get_openspiel_policy())
    np.random.seed(0)
    mt = ref_mccfr.MCCFRTrainer(game)
    it = 0
    out["mccfr_npseed0"] = {}
    for target in (5, 20, 50, 100, 200, 500):
        for _ in range(target - it):
            mt.iteration()
        it = target
        out["mccfr_npseed0"][str(target)] = ms_exploit.exploitability(game, mt.tabular_policy())
    print("exploitability:", json.dumps(out, indent=1))
    dump_json("policies_eval.json", out)


# --------------------------------------------------------------------------- SDCFR
def gen_sdcfr():
    import torch
    import deep_cfr as ref_dcfr  # reference module (src/algorithms/deep_cfr/deep_cfr.py)
    ref_dcfr.tqdm = lambda *a, **k: _NoBar()
    import contextlib
    game = pyspiel.load_game("mini_scopa")
    torch.manual_seed(0)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        d = ref_dcfr.DeepCFR(game, 2, "cpu")
    out = {}
    for p in range(2):
        sd = d.advantage_nets[p].net.state_dict()
        for k, v in sd.items():
            out[f"net{p}.{k}"] = v.numpy().copy()
    # features / masks / advantages / policy at every decision node of the seed-42 tree
    feats, masks, advs, pols, hists = [], [], [], [], []

    def walk(state, hist):
        if state.is_terminal():
            return
        cp = state.current_player()
        f = d._state_to_features(state, cp)
        m = d._get_legal_actions_mask(state, cp)
        a = d.advantage_nets[cp].get_advantages(f, m)[0]
        pol = ref_dcfr.positive_regret_policy(torch.FloatTensor(a).unsqueeze(0),
                                              torch.FloatTensor(m).unsqueeze(0)).numpy()[0]
        feats.append(f); masks.append(m); advs.append(a); pols.append(pol)
        hists.append(hist + [-1] * (8 - len(hist)))
        for act in state.legal_actions():
            c = state.clone(); c.apply_action(act)
            walk(c, hist + [act])

    walk(game.new_initial_state(), [])
    out["node_hist"] = np.array(hists, dtype=np.int8)
    out["node_feat"] = np.array(feats, dtype=np.float32)
    out["node_mask"] = np.array(masks, dtype=np.float32)
    out["node_adv"] = np.array(advs, dtype=np.float32)
    out["node_pol"] = np.array(pols, dtype=np.float32)
    # one traversal per player with the initial nets (np.random.seed fixed just before)
    for p in range(2):
        np.random.seed(100 + p)
        d.advantage_nets[p].buffer.clear()
        val = d._external_sampling_cfr(game.new_initial_state(), p)
        buf = list(d.advantage_nets[p].buffer)
        out[f"trav{p}_value"] = np.array(val, dtype=np.float64)
        out[f"trav{p}_feat"] = np.array([b[0] for b in buf], dtype=np.float32)
        out[f"trav{p}_target"] = np.array([b[1] for b in buf], dtype=np.float32)
        out[f"trav{p}_mask"] = np.array([b[2] for b in buf], dtype=np.float32)
        print(f"SDCFR traversal p{p}: {len(buf)} samples, value {val}")
    np.savez_compressed(os.path.join(OUT, "sdcfr_seed0.npz"), **out)
    print("wrote sdcfr_seed0.npz", os.path.getsize(os.path.join(OUT, "sdcfr_seed0.npz")))


class _NoBar:
    def update(self, *a, **k):
        pass

    def set_postfix(self, *a, **k):
        pass

    def close(self):
        pass


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["deals", "capture", "tree", "traces", "cfr", "mccfr", "eval", "sdcfr"]
    if "deals" in which:
        gen_deals()
    if "capture" in which:
        gen_capture_cases()
    if "tree" in which:
        gen_env_tree()
    if "traces" in which:
        gen_env_random_traces()
    if "cfr" in which:
        gen_cfr()
    if "mccfr" in which:
        gen_mccfr()
    if "eval" in which:
        gen_policies_eval()
    if "sdcfr" in which:
        gen_sdcfr()


# --------------------------------------------------------------------------- SDCFR curve (appended)
def gen_sdcfr_curve():
    """Exploitability (restated BR) of the reference DeepCFR's average policy after k iterations, 3 trials.
    Seeds follow the reference's run_experiments.py:33-34 (torch / numpy seeded with trial_id * 42)."""
    import contextlib
    import torch
    import deep_cfr as ref_dcfr
    import ms_exploit
    ref_dcfr.tqdm = lambda *a, **k: _NoBar()
    game = pyspiel.load_game("mini_scopa")

    class SdPolicy:
        def __init__(self, d):
            self.d, self.cache = d, {}

        def action_probabilities(self, state):
            cp = state.current_player()
            key = state.information_state_string(cp)
            legal = state.legal_actions(cp)
            if key not in self.cache:
                p = self.d.get_policy(state, cp)
                ap = np.array([p[a] for a in legal], dtype=np.float64)
                if np.any(np.isnan(ap)) or ap.sum() <= 0:          # evaluate_vs_random's fallback (:387-390)
                    ap = np.ones(len(legal)) / len(legal)
                else:
                    ap = ap / ap.sum()
                self.cache[key] = ap
            return dict(zip(legal, self.cache[key]))

    n_trials = int(os.environ.get("SDCFR_CURVE_TRIALS", "3"))
    out = {"iterations": [5, 10, 20, 30] if n_trials == 3 else [20, 30], "trials": []}
    for trial in range(n_trials):
        row = []
        for iters in out["iterations"]:
            torch.manual_seed(trial * 42)
            np.random.seed(trial * 42)
            random.seed(trial * 42)
            with contextlib.redirect_stdout(io.StringIO()):
                d = ref_dcfr.DeepCFR(game, 2, "cpu")
                d.train(iterations=iters, advantage_epochs=5, eval_freq=10 ** 9)
            row.append(ms_exploit.exploitability(game, SdPolicy(d)))
            print(f"SDCFR trial {trial} iters {iters}: exploitability {row[-1]:.4f}", flush=True)
        out["trials"].append(row)
    # 3 trials: the round-1 fixture; SDCFR_CURVE_TRIALS=12 writes sdcfr_curve12.json (the bf16 / fp32 curve tests take their
    # tolerance from the spread of these trials)
    dump_json("sdcfr_curve.json" if n_trials == 3 else f"sdcfr_curve{n_trials}.json", out)


if __name__ == "__main__" and "sdcfr_curve" in sys.argv[1:]:
    gen_sdcfr_curve()
