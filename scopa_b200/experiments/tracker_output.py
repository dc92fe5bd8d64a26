"""Result files in the shapes the reference's ExperimentTracker writes (src/experiments/experiment_tracker.py:71-220):

    <name>_data.json                  save_data_for_plotting (:82-160)  -- what the reference's plot_mccfr.py reads
    <name>_run_<i>.csv                save_data_as_csv (:162-176): Iteration, Reward, Scopas_Trained, Scopas_Random, Scopa_Diff
    <name>_run_<i>_exploitability.csv (:178-185), when a run carries exploitability points
    <name>_statistics.csv             (:187-216), when there is more than one run
    <name>.pkl                        (:71-77) a pickle of the runs -- here a list of plain dicts with the field names of
                                      ExperimentMetrics (:13-39): the reference pickles its own dataclass instances, which
                                      cannot be unpickled without its module on the path

Host-side bookkeeping only; the numbers come from the CUDA solver through the runners beside this file.
"""
import csv
import json
import os
import pickle

import numpy as np

RUN_KEYS = ("eval_iterations", "eval_rewards", "eval_scopas_trained", "eval_scopas_random", "eval_scopa_diff",
            "final_reward", "final_scopa_trained", "final_scopa_random", "final_scopa_diff", "num_info_sets")


def plot_data(experiment_name, algorithm, runs):
    """runs: list of dicts with RUN_KEYS (+ optional exploitability_iterations / exploitability_values)."""
    out = {"experiment_name": experiment_name, "algorithm": algorithm if runs else "Unknown", "num_runs": len(runs), "runs": []}
    for i, r in enumerate(runs):
        d = {"run_id": i + 1}
        d.update({k: r[k] for k in RUN_KEYS})
        if r.get("exploitability_iterations"):
            d["exploitability_iterations"] = r["exploitability_iterations"]
            d["exploitability_values"] = r["exploitability_values"]
        out["runs"].append(d)
    if len(runs) > 1:
        arr = lambda k: np.array([r[k] for r in runs])
        rew, st, sr, sd = arr("eval_rewards"), arr("eval_scopas_trained"), arr("eval_scopas_random"), arr("eval_scopa_diff")
        ms = lambda a: {"mean": a.mean(axis=0).tolist(), "std": a.std(axis=0).tolist()}
        out["statistics"] = {
            "eval_iterations": runs[0]["eval_iterations"],
            "rewards": dict(ms(rew), min=rew.min(axis=0).tolist(), max=rew.max(axis=0).tolist()),
            "scopas_trained": ms(st), "scopas_random": ms(sr), "scopa_diff": ms(sd),
            "final_metrics": {
                "reward_mean": float(np.mean([r["final_reward"] for r in runs])),
                "reward_std": float(np.std([r["final_reward"] for r in runs])),
                "scopa_trained_mean": float(np.mean([r["final_scopa_trained"] for r in runs])),
                "scopa_trained_std": float(np.std([r["final_scopa_trained"] for r in runs])),
                "scopa_random_mean": float(np.mean([r["final_scopa_random"] for r in runs])),
                "scopa_random_std": float(np.std([r["final_scopa_random"] for r in runs])),
            },
        }
    return out


def save(experiment_name, algorithm, runs, save_dir="experiments/results"):
    """ExperimentTracker.save: pickle + JSON + CSVs.  -> list of the files written."""
    os.makedirs(save_dir, exist_ok=True)
    p = lambda suffix: os.path.join(save_dir, experiment_name + suffix)
    written = []
    with open(p(".pkl"), "wb") as f:
        pickle.dump([dict(r, algorithm=algorithm, iterations=list(range(max(r["eval_iterations"] or [0])))) for r in runs], f)
    written.append(p(".pkl"))
    with open(p("_data.json"), "w") as f:
        json.dump(plot_data(experiment_name, algorithm, runs), f, indent=2)
    written.append(p("_data.json"))
    for i, r in enumerate(runs):
        with open(p(f"_run_{i + 1}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["Iteration", "Reward", "Scopas_Trained", "Scopas_Random", "Scopa_Diff"])
            for j, it in enumerate(r["eval_iterations"]):
                w.writerow([it, r["eval_rewards"][j], r["eval_scopas_trained"][j], r["eval_scopas_random"][j], r["eval_scopa_diff"][j]])
        written.append(p(f"_run_{i + 1}.csv"))
        if r.get("exploitability_iterations"):
            with open(p(f"_run_{i + 1}_exploitability.csv"), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["Iteration", "Exploitability"])
                for it, v in zip(r["exploitability_iterations"], r["exploitability_values"]):
                    w.writerow([it, v])
            written.append(p(f"_run_{i + 1}_exploitability.csv"))
    if len(runs) > 1:
        arr = lambda k: np.array([r[k] for r in runs])
        rew, st, sr, sd = arr("eval_rewards"), arr("eval_scopas_trained"), arr("eval_scopas_random"), arr("eval_scopa_diff")
        with open(p("_statistics.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["Iteration", "Reward_Mean", "Reward_Std", "Scopas_Trained_Mean", "Scopas_Trained_Std",
                        "Scopas_Random_Mean", "Scopas_Random_Std", "Scopa_Diff_Mean", "Scopa_Diff_Std"])
            for i, it in enumerate(runs[0]["eval_iterations"]):
                w.writerow([it, rew[:, i].mean(), rew[:, i].std(), st[:, i].mean(), st[:, i].std(),
                            sr[:, i].mean(), sr[:, i].std(), sd[:, i].mean(), sd[:, i].std()])
        written.append(p("_statistics.csv"))
    return written
