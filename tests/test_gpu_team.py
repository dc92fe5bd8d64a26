"""2v2 team Miniscopa: CUDA kernels and the drop-in env against traces recorded from the unmodified
reference, and the fused rollout against the oracle at scale.  Bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden_json
from oracle import ms_oracle as ora
from scopa_b200 import codec
from scopa_b200.team import BatchedTeamMiniScopa, team_hand_in_order, unpack_team_state

pytestmark = pytest.mark.gpu


def test_team_kernels_follow_reference_traces():
    traces = load_golden_json("team_env_traces.json.gz")["traces"]
    seeds = np.array([t["seed"] for t in traces], dtype=np.int64)
    b = BatchedTeamMiniScopa().reset(seeds)
    ho = b.hand_order.cpu().numpy().view(np.uint64)

    def check(k, rew):
        st = b.states.cpu().numpy().view(np.uint32)
        for i, t in enumerate(traces):
            snap, u = t["snaps"][k], unpack_team_state(st[i])
            assert u["table"] == snap["table"], (t["seed"], k)
            assert [team_hand_in_order(u["hand_mask"][p], ho[i], p) for p in range(4)] == snap["hands"]
            assert u["cap_mask"] == [codec.mask_of(c) for c in snap["caps"]], (t["seed"], k)
            assert u["scopas"] == snap["scopas"] and u["step_count"] == snap["step"]
            assert u["last_capture_team"] == snap["lct"] and f"player_{u['cur']}" == snap["agent"]
            assert [u["terminal"]] * 4 == snap["term"]
            if rew is not None:
                assert rew[i].tolist() == snap["rew"], (t["seed"], k)

    check(0, None)
    acts = torch.tensor([t["actions"] for t in traces], dtype=torch.uint8, device="cuda")
    for k in range(acts.shape[1]):
        r, _ = b.step(acts[:, k].contiguous())
        check(k + 1, r.cpu().numpy())


def test_team_rollout_bit_exact_vs_oracle():
    seeds = np.random.default_rng(3).integers(1, 2**40, 100_000, dtype=np.int64)
    b = BatchedTeamMiniScopa().reset(seeds)
    actions, rewards, final = b.rollout_random(philox_seed=77, game_offset=5)
    o_act, o_rew, o_sc = ora.team_rollout_random(seeds, 77, game_offset=5)
    assert np.array_equal(actions.cpu().numpy(), o_act)
    assert np.array_equal(rewards.cpu().numpy(), o_rew)
    fin = final.cpu().numpy().view(np.uint32)
    assert np.array_equal(np.stack([(fin[:, 6] >> (4 * p)) & 0xF for p in range(4)], 1), o_sc)
    r = rewards.cpu().numpy()
    assert np.all(r[:, 0] == r[:, 1]) and np.all(r[:, 2] == r[:, 3]) and np.all(r[:, 0] + r[:, 2] == 0)
    assert not np.signbit(r).any() or np.all(r[np.signbit(r)] < 0)          # no negative zeros
    # stepping the recorded actions one ply at a time reaches the same final state
    for k in range(16):
        b.step(actions[:, k].contiguous())
    assert torch.equal(b.states, final)


def test_team_env_drop_in():
    from scopa_b200.envs.team_mini_scopa_game import TeamMiniScopaEnv
    traces = load_golden_json("team_env_traces.json.gz")["traces"][:25]
    ids = lambda lst: [codec.card_id(r, s) for r, s in lst]
    for tr in traces:
        env = TeamMiniScopaEnv(seed=42)
        env.reset(tr["seed"])
        for k, a in enumerate(tr["actions"]):
            env.step(a)
            st, snap = env.get_state(), tr["snaps"][k + 1]
            assert ids(st["table"]) == snap["table"] and [ids(h) for h in st["hands"]] == snap["hands"]
            assert [ids(h) for h in st["captures"]] == snap["caps"], (tr["seed"], k)     # order of captures incl. the sweep
            assert st["scopas"] == snap["scopas"] and st["last_capture_team"] == snap["lct"]
            assert st["agent_selection"] == snap["agent"] and st["step_count"] == snap["step"]
            assert [st["rewards"][n] for n in env.possible_agents] == snap["rew"]
            assert [st["terminations"][n] for n in env.possible_agents] == snap["term"]
    assert env.game.get_team(3) == 1


def test_tpi_wrapper_drop_in():
    """The two-coordinator OpenSpiel view of the team game against traces recorded from the unmodified reference."""
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_team_mini_scopa  # noqa: F401
    traces = load_golden_json("team_tpi_traces.json.gz")["traces"]
    game = pyspiel.load_game("team_mini_scopa_tpi")
    assert game.num_players() == 2

    def rec(state):
        return {"cp": state.current_player(), "term": state.is_terminal(), "legal": list(state.legal_actions()),
                "legal0": list(state.legal_actions(0)), "legal1": list(state.legal_actions(1)),
                "info0": state.information_state_string(0), "info1": state.information_state_string(1),
                "hist": state.history_str(), "rew": [float(x) for x in state.rewards()]}

    for k, tr in enumerate(traces[:30]):
        state = game.new_initial_state()
        assert rec(state) == tr["recs"][0]
        for a, want in zip(tr["actions"], tr["recs"][1:]):
            if k % 2:
                state = state.clone()
            state.apply_action(a)
            assert rec(state) == want, (k, tr["actions"])
