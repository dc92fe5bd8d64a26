"""40-card Scopa: CUDA kernels, the drop-in env and the OpenSpiel wrapper against traces
recorded from the unmodified reference, and the fused rollout against the oracle at scale.  Bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden_json
from oracle import ms_oracle as ora
from scopa_b200 import full
from scopa_b200 import pyspiel_compat as pyspiel
from scopa_b200.envs import openspiel_full_scopa  # noqa: F401  (registers "full_scopa")
from scopa_b200.envs.full_scopa_game import Card, FullDeck, FullScopaEnv

pytestmark = pytest.mark.gpu


def _cid(x):
    r, s = x if isinstance(x, tuple) else (x.rank, x.suit)
    return full.card_id(r, s)


def _snap(env):
    st = env.get_state()
    return {"table": [_cid(t) for t in st["table"]], "hands": [[_cid(c) for c in h] for h in st["hands"]],
            "caps": [[_cid(c) for c in h] for h in st["captures"]], "scopas": list(st["scopas"]),
            "deck": st["deck_remaining"], "round": st["round_number"], "last": st["last_capture"],
            "agent": st["agent_selection"], "step": st["step_count"],
            "rew": [float(st["rewards"][a]) for a in env.possible_agents],
            "term": [bool(st["terminations"][a]) for a in env.possible_agents]}


def test_decks_match_reference_and_oracle():
    g = load_golden_json("full_env_traces.json.gz")["decks"]
    for seed, deck in g.items():
        assert full.deck_from_seed(int(seed)) == deck, seed
        assert [_cid(c) for c in FullDeck(int(seed)).cards] == deck
    seeds = np.concatenate([np.arange(0, 3000), np.random.default_rng(5).integers(2 ** 32, 2 ** 62, 1000)]).astype(np.int64)
    d_seeds = torch.from_numpy(seeds).cuda()
    lib = full._lib.load()
    out = {}
    for slow in (0, 1):
        d = torch.empty((len(seeds), 4), dtype=torch.int64, device="cuda")
        full._lib.check(lib.ms_full_deck_from_seeds(d_seeds.data_ptr(), len(seeds), d.data_ptr(), slow, full._lib.stream_ptr()))
        out[slow] = d.cpu().numpy().view(np.uint64)
    assert np.array_equal(out[0], out[1]), "fast and slow shuffle paths differ"
    for i in range(0, len(seeds), 7):
        assert full.unpack_deck(out[0][i]) == ora.full_deck(int(seeds[i])), seeds[i]


def test_drop_in_env_follows_reference_traces():
    traces = load_golden_json("full_env_traces.json.gz")["traces"]
    for t in traces[:60] + traces[-3:]:
        env = FullScopaEnv(seed=42)
        env.reset(t["seed"])
        assert _snap(env) == t["snaps"][0], t["seed"]
        for k, a in enumerate(t["actions"]):
            env.step(a)
            assert _snap(env) == t["snaps"][k + 1], (t["seed"], k, a)


def test_kernels_follow_reference_traces():
    """Batched device path: every trace at once through ms_full_step, packed state against the recorded lists."""
    traces = [t for t in load_golden_json("full_env_traces.json.gz")["traces"]]
    n_steps = max(len(t["actions"]) for t in traces)
    b = full.BatchedFullScopa().reset(np.array([t["seed"] for t in traces], dtype=np.int64))
    decks = [full.unpack_deck(w) for w in b.decks.cpu().numpy().view(np.uint64)]
    acts = np.full((len(traces), n_steps), 255, dtype=np.uint8)
    for i, t in enumerate(traces):
        acts[i, :len(t["actions"])] = t["actions"]
    d_acts = torch.from_numpy(acts).cuda()
    for k in range(n_steps):
        rew, done = b.step(d_acts[:, k].contiguous())
        st, rew, done = b.states.cpu().numpy().view(np.uint32), rew.cpu().numpy(), done.cpu().numpy()
        for i, t in enumerate(traces):
            if k >= len(t["actions"]):
                continue
            snap, u = t["snaps"][k + 1], full.unpack_full_state(st[i], decks[i])
            assert u["table"] == snap["table"] and u["hands"] == snap["hands"], (t["seed"], k)
            assert u["cap_mask"] == [sum(1 << c for c in set(cs)) for cs in snap["caps"]], (t["seed"], k)
            assert u["scopas"] == snap["scopas"] and u["step_count"] == snap["step"] and u["round_number"] == snap["round"]
            assert u["last_capture"] == snap["last"] and f"player_{u['cur']}" == snap["agent"]
            assert [u["terminal"]] * 2 == snap["term"] and bool(done[i]) == snap["term"][0]
            assert rew[i].tolist() == snap["rew"], (t["seed"], k)
    assert not b.table_overflow()


def test_openspiel_wrapper_follows_reference_traces():
    for t in load_golden_json("full_env_traces.json.gz")["spiel"]:
        s = pyspiel.load_game("full_scopa").new_initial_state()
        rows = t["rows"]
        for k in range(len(rows)):
            row = rows[k]
            assert int(s.current_player()) == row["cp"] and s.is_terminal() == row["term"]
            assert s.legal_actions() == row["legal"] and s.legal_actions(0) == row["legal0"] and s.legal_actions(1) == row["legal1"]
            assert s.information_state_string(0) == row["info0"] and s.information_state_string(1) == row["info1"]
            assert s.history_str() == row["hist"] and [float(x) for x in s.rewards()] == row["rew"]
            if k < len(t["actions"]):
                if k % 5 == 2:
                    s = s.clone()          # (raises in the reference; here the clone must carry on identically)
                s.apply_action(t["actions"][k])


def test_rollout_matches_oracle_at_scale():
    n = 200_000
    seeds = np.arange(1, n + 1, dtype=np.int64)
    b = full.BatchedFullScopa().reset(seeds)
    actions, rewards, final = b.rollout_random(philox_seed=77, game_offset=3)
    o_act, o_rew, o_sc, o_nc, o_mt = ora.full_rollout_random(seeds, 77, game_offset=3)
    assert np.array_equal(actions.cpu().numpy(), o_act)
    assert np.array_equal(rewards.cpu().numpy(), o_rew)
    fin = final.cpu().numpy().view(np.uint32)
    assert np.array_equal(np.stack([fin[:, 6] & 0x3F, (fin[:, 6] >> 6) & 0x3F], 1).astype(np.uint8), o_sc)
    caps = [fin[:, 3].astype(np.uint64) | ((fin[:, 5] & 0xFF).astype(np.uint64) << np.uint64(32)),
            fin[:, 4].astype(np.uint64) | (((fin[:, 5] >> 8) & 0xFF).astype(np.uint64) << np.uint64(32))]
    pop = np.stack([np.array([bin(int(x)).count("1") for x in c[:5000]]) for c in caps], 1)
    assert np.array_equal(pop.astype(np.uint8), o_nc[:5000])
    assert not b.table_overflow() and o_mt.max() <= full.MAX_TABLE
    # step kernel == rollout kernel: replay the recorded actions ply by ply
    b2 = full.BatchedFullScopa().reset(seeds[:20000])
    for k in range(full.PLIES):
        r2, done = b2.step(actions[:20000, k].contiguous())
    assert torch.equal(b2.states, final[:20000]) and torch.equal(r2, rewards[:20000]) and bool(done.all())
    # host-buffer entry point
    h_act, h_rew = np.zeros((1000, 36), dtype=np.uint8), np.zeros((1000, 2), dtype=np.float32)
    full._lib.check(full._lib.load().ms_full_rollout_random_host(seeds.ctypes.data, 1000, 77, 3, h_act.ctypes.data, h_rew.ctypes.data))
    assert np.array_equal(h_act, o_act[:1000]) and np.array_equal(h_rew, o_rew[:1000])
    lib = full._lib.load()
    try:                                                        # several pipeline stages, ragged last one
        lib.ms_debug_set_host_chunk(384)
        h_act[:], h_rew[:] = 0, 0
        full._lib.check(lib.ms_full_rollout_random_host(seeds.ctypes.data, 1000, 77, 3, h_act.ctypes.data, h_rew.ctypes.data))
        assert np.array_equal(h_act, o_act[:1000]) and np.array_equal(h_rew, o_rew[:1000])
    finally:
        lib.ms_debug_set_host_chunk(0)


def test_game_helpers_on_device():
    env = FullScopaEnv(seed=42)
    g = env.game
    g.table = [Card(3, "coppe"), Card(4, "spade"), Card(7, "bastoni"), Card(1, "denari")]
    assert g.find_capture_combinations(Card(7, "denari")) == [[Card(7, "bastoni")]]          # equal rank first
    assert g.find_capture_combinations(Card(8, "denari")) == [[Card(3, "coppe"), Card(4, "spade"), Card(1, "denari")]]
    assert g.find_capture_combinations(Card(2, "denari")) == []
    caps = [Card(7, "denari"), Card(6, "coppe"), Card(1, "spade"), Card(10, "bastoni"), Card(2, "bastoni")]
    assert g.calculate_primiera_score(caps) == 21 + 18 + 16 + 12
    assert g.calculate_primiera_score(caps[:3]) == 0
