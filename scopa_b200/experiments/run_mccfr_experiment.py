"""The reference's MC-CFR experiment protocol (src/experiments/run_mccfr_experiment.py:64-202) on the CUDA
solver, writing the same JSON shape as its shipped results file
(src/experiments/experiments/results/MiniScopa_MCCFR_data.json, produced by ExperimentTracker.save,
src/experiments/experiment_tracker.py:82-220) so the reference's plot_mccfr.py can read it.

Protocol per run: `iterations` x MCCFRTrainer.iteration(); every `eval_interval` iterations 500 episodes
vs a uniform-random opponent (seats swapped at half time); 5000 final episodes.  Episodes are played on the
GPU (ms_eval_policies).

    python -m scopa_b200.experiments.run_mccfr_experiment --runs 10 --out MiniScopa_MCCFR_data.json
"""
import argparse
import json

import numpy as np

from .. import pyspiel_compat as pyspiel
from ..envs import openspiel_mini_scopa  # noqa: F401  (registers the game)
from ..algorithms.mc_cfr import MCCFRTrainer


def _evaluate(solver, n, seed, policy_kind=1):
    """avg reward of the trained seat, avg scopas trained / random (evaluate_policy_quick, :24-61)."""
    tab, uni = solver.average_policy(policy_kind), solver.uniform_policy()
    n0 = int(np.ceil(n / 2))
    r_a, s_a = solver.evaluate(tab, uni, n0, philox_seed=seed, first_game=0)
    r_b, s_b = solver.evaluate(uni, tab, n - n0, philox_seed=seed, first_game=n0)
    reward = (float(r_a.double().sum().item()) - float(r_b.double().sum().item())) / n
    s_a, s_b = s_a.double().sum(0).cpu().numpy(), s_b.double().sum(0).cpu().numpy()
    return reward, (s_a[0] + s_b[1]) / n, (s_a[1] + s_b[0]) / n


def run_single(run_id, iterations=500, eval_interval=5, final_eval_episodes=5000, seed=0):
    game = pyspiel.load_game("mini_scopa")
    trainer = MCCFRTrainer(game=game, seed=seed)
    run = {"run_id": run_id, "eval_iterations": [], "eval_rewards": [], "eval_scopas_trained": [],
           "eval_scopas_random": [], "eval_scopa_diff": []}
    for t in range(0, iterations, eval_interval):
        trainer.iterate(min(eval_interval, iterations - t))
        it = min(t + eval_interval, iterations)
        r, st, sr = _evaluate(trainer.solver, 500, seed * 1_000_003 + it)
        run["eval_iterations"].append(it)
        run["eval_rewards"].append(r)
        run["eval_scopas_trained"].append(st)
        run["eval_scopas_random"].append(sr)
        run["eval_scopa_diff"].append(st - sr)
    r, st, sr = _evaluate(trainer.solver, final_eval_episodes, seed * 1_000_003 + 999_983)
    run.update({"final_reward": r, "final_scopa_trained": st, "final_scopa_random": sr, "final_scopa_diff": st - sr,
                "num_info_sets": len(trainer.info_sets)})
    return run


def run_experiments(num_runs=10, iterations=500, eval_interval=5, final_eval_episodes=5000, base_seed=0):
    runs = [run_single(i + 1, iterations, eval_interval, final_eval_episodes, seed=base_seed + i) for i in range(num_runs)]

    def stat(key, extra=False):
        a = np.array([r[key] for r in runs])
        out = {"mean": a.mean(0).tolist(), "std": a.std(0).tolist()}
        if extra:
            out.update({"min": a.min(0).tolist(), "max": a.max(0).tolist()})
        return out

    fin = lambda k: np.array([r[k] for r in runs])
    return {
        "experiment_name": "MiniScopa_MCCFR", "algorithm": "MC-CFR", "num_runs": num_runs, "runs": runs,
        "statistics": {
            "eval_iterations": runs[0]["eval_iterations"],
            "rewards": stat("eval_rewards", True), "scopas_trained": stat("eval_scopas_trained"),
            "scopas_random": stat("eval_scopas_random"), "scopa_diff": stat("eval_scopa_diff"),
            "final_metrics": {
                "reward_mean": float(fin("final_reward").mean()), "reward_std": float(fin("final_reward").std()),
                "scopa_trained_mean": float(fin("final_scopa_trained").mean()),
                "scopa_trained_std": float(fin("final_scopa_trained").std()),
                "scopa_random_mean": float(fin("final_scopa_random").mean()),
                "scopa_random_std": float(fin("final_scopa_random").std()),
            },
        },
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=10)
    ap.add_argument("--iterations", type=int, default=500)
    ap.add_argument("--eval-interval", type=int, default=5)
    ap.add_argument("--final-episodes", type=int, default=5000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="MiniScopa_MCCFR_data.json")
    ap.add_argument("--out-dir", default=None,
                    help="also write every file ExperimentTracker.save writes (pickle, JSON, per-run and statistics CSVs) here")
    a = ap.parse_args()
    data = run_experiments(a.runs, a.iterations, a.eval_interval, a.final_episodes, a.seed)
    with open(a.out, "w") as f:
        json.dump(data, f, indent=2)
    if a.out_dir:
        from . import tracker_output
        tracker_output.save("MiniScopa_MCCFR", "MC-CFR", data["runs"], a.out_dir)
    fm = data["statistics"]["final_metrics"]
    print(f"final reward vs random: {fm['reward_mean']:.4f} +- {fm['reward_std']:.4f}  "
          f"(reference's shipped file: 1.1545 +- 0.1163)")


if __name__ == "__main__":
    main()
