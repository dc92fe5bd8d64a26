"""Against the only numbers the reference PUBLISHES: the MC-CFR reward-vs-random
results of src/experiments/experiments/results/MiniScopa_MCCFR_data.json (10 runs x 500 iterations, evaluation
every 5 iterations with 500 episodes, 5000 final episodes).  Same protocol, CUDA solver + GPU evaluator."""
import numpy as np
import pytest

from conftest import load_golden_json
from oracle import ms_oracle as ora

pytestmark = pytest.mark.gpu


def _reference_restated_exact_rewards(n_seeds=10, iterations=500):
    """Exact expected reward (seat-swapped protocol, uniform-random opponent) of the policies the reference's
    CURRENT code learns: oracle in numpy-RNG mode (bit-exact restatement of MCCFRTrainer, pinned by
    tests/test_oracle_solvers.py) + an exact walk of the reference-recorded game tree."""
    nodes = load_golden_json("env_tree_seed42.json.gz")["nodes"]
    by_h = {tuple(n["h"]): n for n in nodes}

    def value_p0(pol0, pol1):
        def val(h):
            n = by_h[h]
            if n["term"]:
                return n["rew"][0]
            probs = (pol0 if n["cp"] == 0 else pol1)(n)
            return sum(p * val(h + (a,)) for a, p in zip(n["legal"], probs))
        return val(())

    def uniform(n):
        return [1.0 / len(n["legal"])] * len(n["legal"])

    out = []
    for seed in range(n_seeds):
        t = ora.Table()
        t.mccfr_iterate(iterations, ora.Rng(0, seed))
        keys, _, strat, nl, _ = t.arrays()
        tab = {k: strat[i, :nl[i]] for i, k in enumerate(keys)}

        def pol(n):
            k = f"{n['cp']}|{n['info']}"
            if k in tab and tab[k].sum() > 1e-12:
                return list(tab[k] / tab[k].sum())
            return uniform(n)

        out.append((value_p0(pol, uniform) - value_p0(uniform, pol)) / 2)
    return np.array(out)

# statistics.final_metrics and statistics.rewards.mean[...] of the shipped file
REF_FINAL_MEAN, REF_FINAL_STD = 1.1545, 0.1163
REF_CURVE = {5: 0.473, 10: 0.703, 15: 0.816, 490: 1.151, 495: 1.175, 500: 1.129}
REF_SCOPA_TRAINED, REF_SCOPA_RANDOM = 0.4025, 0.1559


def test_mccfr_experiment_reproduces_the_published_results():
    from scopa_b200.experiments.run_mccfr_experiment import run_experiments
    data = run_experiments(num_runs=10, iterations=500, eval_interval=5, final_eval_episodes=5000, base_seed=7)
    # same file shape as the reference's ExperimentTracker output
    assert set(data) == {"experiment_name", "algorithm", "num_runs", "runs", "statistics"}
    assert set(data["runs"][0]) == {"run_id", "eval_iterations", "eval_rewards", "eval_scopas_trained",
                                    "eval_scopas_random", "eval_scopa_diff", "final_reward", "final_scopa_trained",
                                    "final_scopa_random", "final_scopa_diff", "num_info_sets"}
    st = data["statistics"]
    assert st["eval_iterations"] == list(range(5, 501, 5)) and len(st["rewards"]["mean"]) == 100
    fm = st["final_metrics"]
    # (1) against the reference's CURRENT code, restated bit-exactly and evaluated exactly (no episode noise):
    #     10 numpy seeds give 1.271 +- 0.108.  Stated tolerance: 3 standard errors of the difference + 0.03.
    ref_exact = _reference_restated_exact_rewards()
    se = np.sqrt(ref_exact.var(ddof=1) / len(ref_exact) + fm["reward_std"] ** 2 / data["num_runs"])
    assert abs(fm["reward_mean"] - ref_exact.mean()) < 3 * se + 0.03, (fm, ref_exact.mean(), se)
    # (2) against the PUBLISHED file (1.1545 +- 0.1163 over 10 runs of an unknown numpy seed / code revision):
    #     the current reference code itself sits 0.12 above it (2.4 sigma), so the band here is wide
    assert abs(fm["reward_mean"] - REF_FINAL_MEAN) < 0.30, fm
    assert 0.03 < fm["reward_std"] < 0.30
    assert abs(fm["scopa_trained_mean"] - REF_SCOPA_TRAINED) < 0.08 and abs(fm["scopa_random_mean"] - REF_SCOPA_RANDOM) < 0.03
    mean = dict(zip(st["eval_iterations"], st["rewards"]["mean"]))
    for it, ref in REF_CURVE.items():
        assert abs(mean[it] - ref) < 0.30, (it, mean[it], ref)      # 500-episode evaluations are noisy early on
    late = np.mean([mean[i] for i in range(400, 501, 5)])
    assert abs(late - ref_exact.mean()) < 0.12
    for r in data["runs"]:
        assert 560 <= r["num_info_sets"] <= 738                     # reference runs: 593-732 of 738
