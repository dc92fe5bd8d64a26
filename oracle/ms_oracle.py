"""ctypes wrapper around oracle/_build/libms_oracle.so (the CPU restatement in oracle/ms_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under scopa_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libms_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "ms_oracle.c")
    hdr = os.path.join(HERE, "ms_oracle.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-s", "-B"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


class OraEnv(C.Structure):
    _fields_ = [("hand", (C.c_int * 4) * 2), ("nhand", C.c_int * 2),
                ("caps", (C.c_int * 16) * 2), ("ncaps", C.c_int * 2),
                ("scopas", C.c_int * 2),
                ("table", C.c_int * 16), ("ntable", C.c_int),
                ("agent", C.c_int), ("step_count", C.c_int), ("max_steps", C.c_int), ("seed", C.c_int64),
                ("rewards", C.c_double * 2), ("term", C.c_int * 2)]


class OraTeamEnv(C.Structure):
    _fields_ = [("hand", (C.c_int * 4) * 4), ("nhand", C.c_int * 4), ("caps", (C.c_int * 32) * 4), ("ncaps", C.c_int * 4),
                ("scopas", C.c_int * 4), ("table", C.c_int * 16), ("ntable", C.c_int), ("last_capture_team", C.c_int),
                ("agent", C.c_int), ("step_count", C.c_int), ("max_steps", C.c_int), ("seed", C.c_int64),
                ("rewards", C.c_double * 4), ("term", C.c_int * 4)]


class OraFullEnv(C.Structure):
    _fields_ = [("hand", (C.c_int * 3) * 2), ("nhand", C.c_int * 2), ("caps", (C.c_int * 96) * 2), ("ncaps", C.c_int * 2),
                ("scopas", C.c_int * 2), ("table", C.c_int * 40), ("ntable", C.c_int), ("deck", C.c_int * 40),
                ("deck_pos", C.c_int), ("last_capture", C.c_int), ("round_number", C.c_int), ("agent", C.c_int),
                ("step_count", C.c_int), ("max_steps", C.c_int), ("seed", C.c_int64), ("rewards", C.c_double * 2),
                ("term", C.c_int * 2)]


class OraState(C.Structure):
    _fields_ = [("env", OraEnv), ("is_terminal", C.c_int), ("history", C.c_int * 40), ("nhist", C.c_int)]


class OraMlp(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_float)) for n in ("w1", "b1", "w2", "b2", "w3", "b3")]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    i64, u64, dbl, vp, ci = C.c_int64, C.c_uint64, C.c_double, C.c_void_p, C.c_int
    P = C.POINTER
    L.ora_deck.argtypes = [i64, P(ci)]
    L.ora_card_in_table.argtypes = [P(ci), ci, ci, P(ci)]
    L.ora_card_in_table.restype = ci
    L.ora_env_init.argtypes = [P(OraEnv), i64]
    L.ora_env_reset.argtypes = [P(OraEnv), i64, ci]
    L.ora_env_step.argtypes = [P(OraEnv), ci]
    L.ora_state_init.argtypes = [P(OraState), i64]
    L.ora_state_clone.argtypes = [P(OraState), P(OraState)]
    L.ora_state_apply.argtypes = [P(OraState), ci]
    L.ora_state_current_player.argtypes = [P(OraState)]
    L.ora_state_current_player.restype = ci
    L.ora_state_legal.argtypes = [P(OraState), ci, P(ci)]
    L.ora_state_legal.restype = ci
    L.ora_state_info_string.argtypes = [P(OraState), ci, C.c_char_p, ci]
    L.ora_state_history_str.argtypes = [P(OraState), C.c_char_p, ci]
    L.ora_state_rewards.argtypes = [P(OraState), P(dbl)]
    L.ora_batch_deal.argtypes = [vp, i64, vp]
    L.ora_rollout_random.argtypes = [vp, i64, u64, u64, vp, vp, vp, vp, ci]
    L.ora_table_new.restype = vp
    L.ora_table_free.argtypes = [vp]
    L.ora_table_size.argtypes = [vp]
    L.ora_table_key.argtypes = [vp, ci]
    L.ora_table_key.restype = C.c_char_p
    L.ora_table_nlegal.argtypes = [vp, ci]
    L.ora_table_legal.argtypes = [vp, ci]
    L.ora_table_legal.restype = P(ci)
    L.ora_table_regret.argtypes = [vp, ci]
    L.ora_table_regret.restype = P(dbl)
    L.ora_table_strategy.argtypes = [vp, ci]
    L.ora_table_strategy.restype = P(dbl)
    L.ora_table_find.argtypes = [vp, C.c_char_p]
    L.ora_cfr_train.argtypes = [vp, i64, ci]
    L.ora_rng_new.argtypes = [ci, u64]
    L.ora_rng_new.restype = vp
    L.ora_rng_free.argtypes = [vp]
    L.ora_mccfr_iterate.argtypes = [vp, i64, ci, vp, u64]
    L.ora_mccfr_populate.argtypes = [vp, i64]
    L.ora_mccfr_batch.argtypes = [vp, i64, ci, u64, u64, i64, P(i64), P(i64)]
    L.ora_mccfr_batch_seq.argtypes = [vp, i64, ci, u64, u64, i64, P(i64), P(i64)]
    L.ora_exploitability.argtypes = [vp, ci, i64, P(dbl)]
    L.ora_exploitability.restype = dbl
    L.ora_features.argtypes = [P(OraState), ci, vp, vp]
    L.ora_mlp_forward.argtypes = [P(OraMlp), vp, vp]
    L.ora_advantages_policy.argtypes = [P(OraMlp), vp, vp, vp, vp]
    L.ora_sdcfr_traverse.argtypes = [P(OraMlp), i64, ci, vp, u64, vp, vp, vp, ci, P(ci)]
    L.ora_sdcfr_traverse.restype = C.c_float
    L.ora_mccfr_bench.argtypes = [i64, i64, ci, u64, P(i64), P(i64)]
    L.ora_team_init.argtypes = [P(OraTeamEnv), i64]
    L.ora_team_reset.argtypes = [P(OraTeamEnv), i64, ci]
    L.ora_team_step.argtypes = [P(OraTeamEnv), ci]
    L.ora_team_rollout_random.argtypes = [vp, i64, u64, u64, vp, vp, vp, ci]
    L.ora_mccfr_batch_mode.argtypes = [vp, i64, ci, ci, u64, u64, i64, P(i64), P(i64)]
    L.ora_full_deck.argtypes = [i64, P(ci)]
    L.ora_full_init.argtypes = [P(OraFullEnv), i64]
    L.ora_full_reset.argtypes = [P(OraFullEnv), i64, ci]
    L.ora_full_step.argtypes = [P(OraFullEnv), ci]
    L.ora_full_capture.argtypes = [P(ci), ci, ci, P(ci)]
    L.ora_full_capture.restype = ci
    L.ora_full_rollout_random.argtypes = [vp, i64, u64, u64, vp, vp, vp, vp, vp, ci]
    L.ora_md_new.restype = vp
    L.ora_md_free.argtypes = [vp]
    L.ora_md_size.argtypes = [vp]
    L.ora_md_size.restype = i64
    L.ora_md_key.argtypes = [vp, i64]
    L.ora_md_key.restype = C.c_char_p
    L.ora_md_packed_key.argtypes = [vp, i64]
    L.ora_md_packed_key.restype = u64
    L.ora_md_nlegal.argtypes = [vp, i64]
    L.ora_md_regret.argtypes = [vp, i64]
    L.ora_md_regret.restype = P(dbl)
    L.ora_md_strategy.argtypes = [vp, i64]
    L.ora_md_strategy.restype = P(dbl)
    L.ora_md_batch.argtypes = [vp, vp, i64, ci, u64, u64, i64, P(i64), P(i64)]
    L.ora_md_apply.argtypes = [vp]
    L.ora_md_populate.argtypes = [vp, vp, i64]
    L.ora_md_batch_blocked.argtypes = [vp, vp, i64, ci, u64, u64, i64, i64, P(i64), P(i64)]
    _lib = L
    return L


# --------------------------------------------------------------------------- convenience wrappers
def deck(seed):
    out = (C.c_int * 16)()
    lib().ora_deck(seed, out)
    return list(out)


def card_in_table(table, card):
    t = (C.c_int * 16)(*table)
    pos = (C.c_int * 8)()
    n = lib().ora_card_in_table(t, len(table), card, pos)
    return n > 0, [pos[i] for i in range(n)]


class Env:
    """ora_env with the attribute names of the reference MiniScopaEnv that the fixtures record."""

    def __init__(self, seed=42):
        self.e = OraEnv()
        lib().ora_env_init(C.byref(self.e), seed)

    def reset(self, seed=None):
        lib().ora_env_reset(C.byref(self.e), 0 if seed is None else seed, 0 if seed is None else 1)

    def step(self, action):
        lib().ora_env_step(C.byref(self.e), action)

    def snapshot(self):
        e = self.e
        return {
            "table": [e.table[i] for i in range(e.ntable)],
            "hands": [[e.hand[p][i] for i in range(e.nhand[p])] for p in range(2)],
            "caps": [[e.caps[p][i] for i in range(e.ncaps[p])] for p in range(2)],
            "scopas": [e.scopas[0], e.scopas[1]],
            "agent": f"player_{e.agent}",
            "step": e.step_count,
            "rew": [e.rewards[0], e.rewards[1]],
            "term": [bool(e.term[0]), bool(e.term[1])],
        }


class State:
    """ora_state with the pyspiel.State protocol used by the reference solvers."""

    def __init__(self, seed=42, _raw=None):
        self.s = OraState()
        if _raw is None:
            lib().ora_state_init(C.byref(self.s), seed)

    def clone(self):
        o = State(_raw=True)
        lib().ora_state_clone(C.byref(self.s), C.byref(o.s))
        return o

    def child(self, action):
        c = self.clone()
        c.apply_action(action)
        return c

    def apply_action(self, a):
        lib().ora_state_apply(C.byref(self.s), a)

    def current_player(self):
        return lib().ora_state_current_player(C.byref(self.s))

    def is_terminal(self):
        return bool(self.s.is_terminal)

    def is_chance_node(self):
        return False

    def legal_actions(self, player=None):
        out = (C.c_int * 4)()
        n = lib().ora_state_legal(C.byref(self.s), -1 if player is None else player, out)
        return [out[i] for i in range(n)]

    def information_state_string(self, player=None):
        buf = C.create_string_buffer(96)
        lib().ora_state_info_string(C.byref(self.s), -100 if player is None else player, buf, 96)
        return buf.value.decode()

    def history_str(self):
        buf = C.create_string_buffer(320)
        lib().ora_state_history_str(C.byref(self.s), buf, 320)
        return buf.value.decode()

    def rewards(self):
        out = (C.c_double * 2)()
        lib().ora_state_rewards(C.byref(self.s), out)
        return [out[0], out[1]]

    returns = rewards

    def record(self):
        e = self.s.env
        return {
            "cp": self.current_player(), "term": self.is_terminal(),
            "legal": self.legal_actions(), "legal0": self.legal_actions(0), "legal1": self.legal_actions(1),
            "info": self.information_state_string(), "info0": self.information_state_string(0),
            "info1": self.information_state_string(1), "hist": self.history_str(),
            "rew": self.rewards(),
            "hands": [[e.hand[p][i] for i in range(e.nhand[p])] for p in range(2)],
            "caps": [[e.caps[p][i] for i in range(e.ncaps[p])] for p in range(2)],
            "scopas": [e.scopas[0], e.scopas[1]],
            "table": [e.table[i] for i in range(e.ntable)],
            "step": e.step_count, "agent": f"player_{e.agent}",
        }


class Table:
    def __init__(self):
        self.t = lib().ora_table_new()

    def __del__(self):
        try:
            lib().ora_table_free(self.t)
        except Exception:
            pass

    def __len__(self):
        return lib().ora_table_size(self.t)

    def arrays(self):
        """-> keys (list[str], first-touch order), regret [n,4], strategy [n,4], nlegal [n], legal [n,4]"""
        L = lib()
        n = len(self)
        keys, reg, strat = [], np.zeros((n, 4)), np.zeros((n, 4))
        nl = np.zeros(n, dtype=np.int8)
        legal = np.full((n, 4), -1, dtype=np.int8)
        for i in range(n):
            keys.append(L.ora_table_key(self.t, i).decode())
            m = L.ora_table_nlegal(self.t, i)
            nl[i] = m
            r, s, lg = L.ora_table_regret(self.t, i), L.ora_table_strategy(self.t, i), L.ora_table_legal(self.t, i)
            for a in range(m):
                reg[i, a], strat[i, a], legal[i, a] = r[a], s[a], lg[a]
        return keys, reg, strat, nl, legal

    def set_arrays(self, reg, strat):
        L = lib()
        for i in range(len(self)):
            r, s = L.ora_table_regret(self.t, i), L.ora_table_strategy(self.t, i)
            for a in range(4):
                r[a], s[a] = reg[i, a], strat[i, a]

    def cfr_train(self, iters, seed=42):
        lib().ora_cfr_train(self.t, seed, iters)

    def mccfr_iterate(self, iters, rng, seed=42, first_iter=0):
        lib().ora_mccfr_iterate(self.t, seed, iters, rng.r, first_iter)

    def mccfr_populate(self, seed=42):
        lib().ora_mccfr_populate(self.t, seed)

    def mccfr_batch(self, player, philox_seed, first_trav, ntrav, seed=42):
        nu, nv = C.c_int64(), C.c_int64()
        lib().ora_mccfr_batch(self.t, seed, player, philox_seed, first_trav, ntrav, C.byref(nu), C.byref(nv))
        return nu.value, nv.value

    def mccfr_batch_seq(self, player, philox_seed, first_trav, ntrav, seed=42):
        """frozen-sigma batch on the sequential 32-bit Philox stream (the headline kernel's stream)"""
        nu, nv = C.c_int64(), C.c_int64()
        lib().ora_mccfr_batch_seq(self.t, seed, player, philox_seed, first_trav, ntrav, C.byref(nu), C.byref(nv))
        return nu.value, nv.value

    def mccfr_batch_mode(self, mode, player, philox_seed, first_trav, ntrav, seed=42):
        nu, nv = C.c_int64(), C.c_int64()
        lib().ora_mccfr_batch_mode(self.t, seed, mode, player, philox_seed, first_trav, ntrav, C.byref(nu), C.byref(nv))
        return nu.value, nv.value

    def exploitability(self, policy_kind, seed=42):
        br = (C.c_double * 2)()
        e = lib().ora_exploitability(self.t, policy_kind, seed, br)
        return e, [br[0], br[1]]


class Rng:
    def __init__(self, kind, seed):
        self.r = lib().ora_rng_new(kind, seed)

    def __del__(self):
        try:
            lib().ora_rng_free(self.r)
        except Exception:
            pass


def batch_deal(seeds):
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    out = np.zeros((len(seeds), 8), dtype=np.int32)
    lib().ora_batch_deal(seeds.ctypes.data, len(seeds), out.ctypes.data)
    return out


def rollout_random(seeds, philox_seed, nthreads=0, game_offset=0):
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    n = len(seeds)
    actions = np.zeros((n, 8), dtype=np.uint8)
    rewards = np.zeros((n, 2), dtype=np.float32)
    scopas = np.zeros((n, 2), dtype=np.uint8)
    ncaps = np.zeros((n, 2), dtype=np.uint8)
    lib().ora_rollout_random(seeds.ctypes.data, n, philox_seed, game_offset, actions.ctypes.data, rewards.ctypes.data,
                             scopas.ctypes.data, ncaps.ctypes.data, nthreads)
    return actions, rewards, scopas, ncaps


class Mlp:
    """Holds float32 copies of the six FlexibleNet tensors and the ora_mlp view of them."""

    def __init__(self, w1, b1, w2, b2, w3, b3):
        self.arr = [np.ascontiguousarray(a, dtype=np.float32) for a in (w1, b1, w2, b2, w3, b3)]
        assert self.arr[0].shape == (128, 34) and self.arr[2].shape == (64, 128) and self.arr[4].shape == (16, 64)
        self.c = OraMlp(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in self.arr])


def features(state, player):
    f = np.zeros(34, dtype=np.float32)
    m = np.zeros(16, dtype=np.float32)
    lib().ora_features(C.byref(state.s), player, f.ctypes.data, m.ctypes.data)
    return f, m


def advantages_policy(mlp, feat, mask):
    adv = np.zeros(16, dtype=np.float32)
    pol = np.zeros(16, dtype=np.float32)
    f = np.ascontiguousarray(feat, dtype=np.float32)
    m = np.ascontiguousarray(mask, dtype=np.float32)
    lib().ora_advantages_policy(C.byref(mlp.c), f.ctypes.data, m.ctypes.data, adv.ctypes.data, pol.ctypes.data)
    return adv, pol


def sdcfr_traverse(mlps, player, rng, trav_id=0, seed=42, cap=64):
    nets = (OraMlp * 2)(mlps[0].c, mlps[1].c)
    feat = np.zeros((cap, 34), dtype=np.float32)
    target = np.zeros((cap, 16), dtype=np.float32)
    mask = np.zeros((cap, 16), dtype=np.float32)
    n = C.c_int()
    v = lib().ora_sdcfr_traverse(nets, seed, player, rng.r, trav_id, feat.ctypes.data, target.ctypes.data,
                                 mask.ctypes.data, cap, C.byref(n))
    return float(v), feat[:n.value], target[:n.value], mask[:n.value]


def mccfr_bench(ntrav_per_thread, nthreads, philox_seed=0, seed=42):
    u, v = C.c_int64(), C.c_int64()
    lib().ora_mccfr_bench(seed, ntrav_per_thread, nthreads, philox_seed, C.byref(u), C.byref(v))
    return u.value, v.value


class TeamEnv:
    """ora_team_env with the snapshot shape of oracle/gen_golden_team.py."""

    def __init__(self, seed=42):
        self.e = OraTeamEnv()
        lib().ora_team_init(C.byref(self.e), seed)

    def reset(self, seed=None):
        lib().ora_team_reset(C.byref(self.e), 0 if seed is None else seed, 0 if seed is None else 1)

    def step(self, action):
        lib().ora_team_step(C.byref(self.e), action)

    def snapshot(self):
        e = self.e
        return {"table": [e.table[i] for i in range(e.ntable)],
                "hands": [[e.hand[p][i] for i in range(e.nhand[p])] for p in range(4)],
                "caps": [[e.caps[p][i] for i in range(e.ncaps[p])] for p in range(4)],
                "scopas": [e.scopas[p] for p in range(4)],
                "lct": None if e.last_capture_team < 0 else e.last_capture_team,
                "agent": f"player_{e.agent}", "step": e.step_count,
                "rew": [e.rewards[p] for p in range(4)], "term": [bool(e.term[p]) for p in range(4)]}


def team_rollout_random(seeds, philox_seed, nthreads=0, game_offset=0):
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    n = len(seeds)
    actions = np.zeros((n, 16), dtype=np.uint8)
    rewards = np.zeros((n, 4), dtype=np.float32)
    scopas = np.zeros((n, 4), dtype=np.uint8)
    lib().ora_team_rollout_random(seeds.ctypes.data, n, philox_seed, game_offset, actions.ctypes.data, rewards.ctypes.data,
                                  scopas.ctypes.data, nthreads)
    return actions, rewards, scopas


class MultiDealTable:
    """Multi-deal MCCFR table (chance root over `seeds`; parity unpinned beyond one deal)."""

    def __init__(self, seeds):
        self.seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        self.t = lib().ora_md_new()

    def __del__(self):
        try:
            lib().ora_md_free(self.t)
        except Exception:
            pass

    def __len__(self):
        return lib().ora_md_size(self.t)

    def batch(self, player, philox_seed, first_trav, ntrav):
        nu, nv = C.c_int64(0), C.c_int64(0)
        lib().ora_md_batch(self.t, self.seeds.ctypes.data, len(self.seeds), player, philox_seed, first_trav, ntrav,
                           C.byref(nu), C.byref(nv))
        return nu.value, nv.value

    def apply(self):
        lib().ora_md_apply(self.t)

    def populate(self):
        lib().ora_md_populate(self.t, self.seeds.ctypes.data, len(self.seeds))

    def batch_blocked(self, player, philox_seed, first_visit, n_visits, pairs):
        nu, nv = C.c_int64(0), C.c_int64(0)
        lib().ora_md_batch_blocked(self.t, self.seeds.ctypes.data, len(self.seeds), player, philox_seed, first_visit, n_visits,
                                   pairs, C.byref(nu), C.byref(nv))
        return nu.value, nv.value

    def arrays(self):
        """-> string keys, packed keys (uint64), regret [n,4], strategy [n,4], nlegal [n] (first-touch order)"""
        L, n = lib(), len(self)
        keys, packed = [], np.zeros(n, dtype=np.uint64)
        reg, strat, nl = np.zeros((n, 4)), np.zeros((n, 4)), np.zeros(n, dtype=np.int8)
        for i in range(n):
            keys.append(L.ora_md_key(self.t, i).decode())
            packed[i] = L.ora_md_packed_key(self.t, i)
            nl[i] = L.ora_md_nlegal(self.t, i)
            r, s = L.ora_md_regret(self.t, i), L.ora_md_strategy(self.t, i)
            for a in range(nl[i]):
                reg[i, a], strat[i, a] = r[a], s[a]
        return keys, packed, reg, strat, nl


# --------------------------------------------------------------------------- 40-card Scopa
def full_deck(seed):
    out = (C.c_int * 40)()
    lib().ora_full_deck(seed, out)
    return list(out)


class FullEnv:
    """ora_full_env with the snapshot format of tests/golden/full_env_traces.json.gz."""

    def __init__(self, seed=42):
        self.e = OraFullEnv()
        lib().ora_full_init(C.byref(self.e), seed)

    def reset(self, seed=None):
        lib().ora_full_reset(C.byref(self.e), 0 if seed is None else seed, 0 if seed is None else 1)

    def step(self, action):
        lib().ora_full_step(C.byref(self.e), action)

    def snapshot(self):
        e = self.e
        return {"table": [e.table[i] for i in range(e.ntable)],
                "hands": [[e.hand[p][i] for i in range(e.nhand[p])] for p in range(2)],
                "caps": [[e.caps[p][i] for i in range(e.ncaps[p])] for p in range(2)],
                "scopas": [e.scopas[0], e.scopas[1]], "deck": 40 - e.deck_pos, "round": e.round_number,
                "last": None if e.last_capture < 0 else e.last_capture, "agent": f"player_{e.agent}", "step": e.step_count,
                "rew": [e.rewards[0], e.rewards[1]], "term": [bool(e.term[0]), bool(e.term[1])]}


def full_rollout_random(seeds, philox_seed, nthreads=0, game_offset=0):
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    n = len(seeds)
    actions, rewards = np.zeros((n, 36), dtype=np.uint8), np.zeros((n, 2), dtype=np.float32)
    scopas, ncaps, maxt = np.zeros((n, 2), dtype=np.uint8), np.zeros((n, 2), dtype=np.uint8), np.zeros(n, dtype=np.uint8)
    lib().ora_full_rollout_random(seeds.ctypes.data, n, philox_seed, game_offset, actions.ctypes.data, rewards.ctypes.data,
                                  scopas.ctypes.data, ncaps.ctypes.data, maxt.ctypes.data, nthreads)
    return actions, rewards, scopas, ncaps, maxt
