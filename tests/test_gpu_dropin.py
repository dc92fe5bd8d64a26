"""The drop-in Python classes (scopa_b200.envs / scopa_b200.algorithms) read like the reference's own
usage and are checked against fixtures recorded from the unmodified reference."""
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, load_golden_json
from scopa_b200 import codec

pytestmark = pytest.mark.gpu


def _ids(cards):
    return [codec.card_id(c.rank, c.suit) for c in cards]


def _tuples_to_ids(lst):
    return [codec.card_id(r, s) for r, s in lst]


def test_minideck_and_env_reset():
    from scopa_b200.envs.mini_scopa_game import MiniDeck, MiniScopaEnv
    g = load_golden_json("deals.json")
    for seed in ("1", "42", "43", "12345", "8589934599", "-42", "2147483648"):
        assert _ids(MiniDeck(int(seed)).cards) == g["decks"][seed]
    d = MiniDeck(42)
    first = d.deal(4)
    assert _ids(first) == [7, 9, 5, 6] and len(d.cards) == 12 and repr(first[0]) == "9_of_fiori"
    env = MiniScopaEnv()
    env.reset(0)
    assert [_ids(p.hand) for p in env.game.players] == g["env_reset_0_hands"]
    assert env.agent_selection == "player_0" and env.step_count == 0 and env.max_steps == 8
    assert env.rewards == {"player_0": 0, "player_1": 0}


def test_env_step_traces_match_reference_get_state():
    from scopa_b200.envs.mini_scopa_game import MiniScopaEnv
    traces = [t for t in load_golden_json("env_random_traces.json.gz")["traces"] if t["kind"] == "env"][:60]
    for tr in traces:
        env = MiniScopaEnv(seed=42)
        env.reset(tr["seed"])
        for k, a in enumerate(tr["actions"]):
            env.step(a)
            st = env.get_state()
            snap = tr["snaps"][k + 1]
            assert _tuples_to_ids(st["table"]) == snap["table"], (tr["seed"], k)
            assert [_tuples_to_ids(h) for h in st["hands"]] == snap["hands"]
            assert [_tuples_to_ids(h) for h in st["captures"]] == snap["caps"]      # ORDER of captures too
            assert st["scopas"] == snap["scopas"] and st["step_count"] == snap["step"]
            assert st["agent_selection"] == snap["agent"]
            assert [st["rewards"][a_] for a_ in env.possible_agents] == snap["rew"]
            assert [st["terminations"][a_] for a_ in env.possible_agents] == snap["term"]


def test_game_level_calls():
    from scopa_b200.envs.mini_scopa_game import Card, MiniScopaGame
    cases = load_golden_json("capture_cases.json")[:150]
    g = MiniScopaGame()
    for table, played, isin, mask in cases:
        g.table = [Card(codec.RANK_OF[c], codec.SUIT_OF[c]) for c in table]
        got_in, combo = g.card_in_table(Card(codec.RANK_OF[played], codec.SUIT_OF[played]))
        assert int(got_in) == isin
        assert [table.index(c) for c in _ids(combo)] == [i for i in range(len(table)) if (mask >> i) & 1]
    g.reset(42)
    p0 = g.players[0]
    g.play_card(p0.hand[0], p0)                       # 9f on an empty table: placed
    assert _ids(g.table) == [7] and _ids(p0.hand) == [9, 5, 6]
    with pytest.raises(ValueError):
        g.play_card(Card(2, "cuori"), p0)
    assert g.evaluate_game() == [0, 0]


def test_openspiel_state_exhaustive_tree():
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401
    nodes = load_golden_json("env_tree_seed42.json.gz")["nodes"]
    game = pyspiel.load_game("mini_scopa")
    assert game.num_players() == 2
    got = []

    def rec(state, history):
        env, gm = state.env, state.env.game
        got.append({
            "h": list(history), "cp": state.current_player(), "term": bool(state.is_terminal()),
            "legal": list(state.legal_actions()), "legal0": list(state.legal_actions(0)),
            "legal1": list(state.legal_actions(1)), "info": state.information_state_string(),
            "info0": state.information_state_string(0), "info1": state.information_state_string(1),
            "hist": state.history_str(), "rew": [float(x) for x in state.rewards()],
            "hands": [_ids(p.hand) for p in gm.players], "caps": [_ids(p.captures) for p in gm.players],
            "scopas": [p.scopas for p in gm.players], "table": _ids(gm.table), "step": env.step_count,
            "agent": env.agent_selection})
        if state.is_terminal():
            return
        for a in state.legal_actions():
            c = state.clone()
            c.apply_action(a)
            rec(c, history + [a])

    rec(game.new_initial_state(), [])
    want = {tuple(n["h"]): n for n in nodes}
    assert len(got) == 2229
    for g_ in got:
        assert g_ == want[tuple(g_["h"])], g_["h"]
    assert any(g_["term"] for g_ in got)


def test_cfr_trainer_drop_in():
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401
    from scopa_b200.algorithms.vanilla_cfr import CFRTrainer, RandomPolicy, evaluate_agent
    g = np.load(os.path.join(GOLDEN, "cfr_seed42.npz"))
    game = pyspiel.load_game("mini_scopa")
    trainer = CFRTrainer(game=game)
    history = trainer.train(steps=20, eval_interval=5, compute_exploitability=True)
    assert [h[0] for h in history] == [5, 10, 15, 20]
    ev = load_golden_json("policies_eval.json")["cfr"]
    assert abs(history[0][1] - ev["5"]) < 1e-9 and abs(history[1][1] - ev["10"]) < 1e-9 and abs(history[3][1] - ev["20"]) < 1e-9
    m = trainer.info_set_map
    assert list(m.keys()) == list(g["keys"]) and len(m) == 738
    for i, k in enumerate(g["keys"]):
        n = int(g["nlegal"][i])
        assert np.array_equal(m[k].regret_sum, g["reg_20"][i, :n]) and np.array_equal(m[k].strategy_sum, g["strat_20"][i, :n])
        assert m[k].legal_actions.tolist() == g["legal"][i, :n].tolist()
        assert np.array_equal(m[k].local_strategy, m[k].get_strategy())
    # the private entry point the reference's experiment runner calls directly
    t2 = CFRTrainer(game=game)
    for _ in range(2):
        for i in range(game.num_players()):
            t2._cfr_recursive(game.new_initial_state(), i, 1.0, 1.0)
    assert np.array_equal(t2.info_set_map[g["keys"][0]].regret_sum, g["reg_2"][0, :4])
    # a clone() of the fresh root is the same root (its step limit differs: the reference's clone hard-codes 16)
    t3 = CFRTrainer(game=game)
    for _ in range(2):
        for i in range(game.num_players()):
            t3._cfr_recursive(game.new_initial_state().clone(), i, 1.0, 1.0)
    assert np.array_equal(t3.info_set_map[g["keys"][0]].regret_sum, g["reg_2"][0, :4])
    child = game.new_initial_state()
    child.apply_action(child.legal_actions()[0])
    with pytest.raises(NotImplementedError):
        t3._cfr_recursive(child, 0, 1.0, 1.0)
    np.random.seed(3)
    avg, hist, stats = evaluate_agent(game, trainer.get_openspiel_policy(), RandomPolicy(game), num_episodes=60)
    assert len(hist) == 60 and stats["data_collected"] and -5 < avg < 5
    assert set(stats) == {"trained_avg", "opponent_avg", "difference", "history", "data_collected"}


def test_mccfr_trainer_drop_in():
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401
    from scopa_b200.algorithms.mc_cfr import MCCFRTrainer, RandomPolicy, evaluate_agent
    game = pyspiel.load_game("mini_scopa")
    trainer = MCCFRTrainer(game=game, seed=5)
    trainer.iteration()
    n1 = len(trainer.info_sets)
    assert 150 < n1 < 738 and all(isinstance(k, tuple) and k[1].startswith(f"P{k[0]}:") for k in trainer.info_sets)
    assert trainer.train(iterations=300) == []
    e300 = trainer.exploitability()
    assert 580 <= len(trainer.info_sets) <= 738          # reference runs report 593-732 after 500 iterations
    assert 0.35 < e300 < 0.75
    pol = trainer.tabular_policy()
    s = game.new_initial_state()
    probs = pol.action_probabilities(s)
    assert list(probs) == [7, 9, 5, 6] and abs(sum(probs.values()) - 1) < 1e-12
    np.random.seed(4)
    avg, hist, stats = evaluate_agent(game, pol, RandomPolicy(game), num_episodes=80)
    assert len(hist) == 80
    # batched mode: thousands of traversals per launch against a frozen table
    tb = MCCFRTrainer(game=game, seed=5, traversals_per_iteration=2048)
    tb.train(iterations=30)
    assert len(tb.info_sets) == 738 and tb.exploitability() < 0.8


def test_deep_cfr_drop_in():
    import torch
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401
    from scopa_b200.algorithms.deep_cfr import DeepCFR
    g = np.load(os.path.join(GOLDEN, "sdcfr_seed0.npz"))
    game = pyspiel.load_game("mini_scopa")
    torch.manual_seed(0)
    np.random.seed(0)
    d = DeepCFR(game, 2, "cuda")
    assert d.input_dim == 34
    # features / mask / advantages with the reference's weights loaded
    for p in range(2):
        sd = {k[len(f"net{p}."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"net{p}.")}
        d.advantage_nets[p].net.load_state_dict(sd)
    s = game.new_initial_state()
    f, m = d._state_to_features(s, 0), d._get_legal_actions_mask(s, 0)
    assert np.array_equal(f, g["node_feat"][0]) and np.array_equal(m, g["node_mask"][0])
    adv = d.advantage_nets[0].get_advantages(f, m)
    np.testing.assert_allclose(adv[0], g["node_adv"][0], rtol=2e-5, atol=2e-6)
    # one traversal per player like the reference: 41 samples each (deep_cfr.py:284-346)
    for p in range(2):
        v = d._external_sampling_cfr(game.new_initial_state(), p)
        assert len(d.advantage_nets[p].buffer) == 41 and np.isfinite(v)
    d.train(iterations=6, advantage_epochs=5, eval_freq=5, eval_episodes=10)
    h = d.training_history
    assert len(h["losses"][0]) == 6 and len(h["values"][1]) == 6 and h["buffer_sizes"][0][-1] == 41 * 7
    assert len(h["eval_rewards"]) == 2 and len(d.strategy_buffers[0].strategies) == 5
    assert d.strategy_buffers[0].weights == [2, 3, 4, 5, 6]
    pol = d.get_policy(game.new_initial_state(), 0)
    assert pol.dtype == np.float32 and pol.shape == (16,) and (pol >= 0).all() and pol.sum() < 1 + 1e-5
    assert pol[[0, 1, 2, 3, 4, 8, 10]].sum() == 0
    # batched + tensor-core configuration
    d2 = DeepCFR(game, 2, "cuda", precision="bf16", traversals_per_iteration=512)
    d2.train(iterations=3, advantage_epochs=4, eval_freq=100, eval_episodes=0)
    assert d2.training_history["buffer_sizes"][1][-1] == 3 * 512 * 41


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sdcfr_exploitability_curve_vs_reference(precision):
    """Statistical parity for SDCFR, for BOTH inference paths (fp32 on CUDA cores = the reference's precision; bf16 operands
    on the tcgen05 tensor cores = narrower than the reference).  tests/golden/sdcfr_curve12.json holds the exploitability
    (restated BR) of the UNMODIFIED reference's average policy after 20 / 30 iterations for 12 trials (seeds trial*42, as
    the reference's run_experiments.py:33-34; oracle/gen_golden.py sdcfr_curve with SDCFR_CURVE_TRIALS=12): mean 1.753 /
    1.551, standard deviation 0.49 per trial (uniform play = 2.26) -- SDCFR at 30 iterations of ONE traversal is noisy.
    Stated tolerance: |mean of 16 of our seeds - mean of the 12 reference trials| < 3 standard errors of that difference
    (sqrt(se_ref^2 + se_ours^2), both from the sample spreads: about 0.55), same hyper-parameters (1 traversal per player
    per iteration, 5 advantage epochs); and the curve must fall from 20 to 30 iterations like the reference's."""
    import torch
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401
    from scopa_b200.algorithms.deep_cfr import DeepCFR
    ref = load_golden_json("sdcfr_curve12.json")
    assert ref["iterations"] == [20, 30]
    ref_t = np.array(ref["trials"])
    game = pyspiel.load_game("mini_scopa")
    n_seeds = 16
    ours = {20: [], 30: []}
    for iters, offset in ((20, 0), (30, 1)):
        for seed in range(n_seeds):
            torch.manual_seed(seed * 42 + offset)
            np.random.seed(seed * 42 + offset)
            d = DeepCFR(game, 2, "cuda", seed=100 * offset + seed, precision=precision)
            d.train(iterations=iters, advantage_epochs=5, eval_freq=10 ** 9, eval_episodes=0)
            ours[iters].append(d.exploitability())
            if seed == 0 and iters == 20 and precision == "fp32":   # the batched table equals the per-state get_policy()
                tab = d.average_policy_table().cpu().numpy()
                st = d._solver.static_table()
                s = game.new_initial_state()
                for _ in range(3):
                    cp = s.current_player()
                    p16 = d.get_policy(s, cp)
                    slot = st["strings"].index(s.information_state_string(cp))
                    legal = s.legal_actions(cp)
                    want = p16[legal] / p16[legal].sum() if p16[legal].sum() > 0 else np.ones(len(legal)) / len(legal)
                    np.testing.assert_allclose(tab[slot, :len(legal)], want, rtol=1e-4, atol=1e-5)
                    s.apply_action(legal[0])
    means = {}
    for k, idx in ((20, 0), (30, 1)):
        o = np.array(ours[k])
        se = float(np.sqrt(ref_t[:, idx].var(ddof=1) / len(ref_t) + o.var(ddof=1) / len(o)))
        means[k] = float(o.mean())
        assert abs(means[k] - ref_t[:, idx].mean()) < 3.0 * se, (precision, k, means[k], float(ref_t[:, idx].mean()), se, ours[k])
        assert means[k] < 2.26
    assert means[30] < means[20] + 0.15, means          # learning continues (reference: 1.753 -> 1.551)
