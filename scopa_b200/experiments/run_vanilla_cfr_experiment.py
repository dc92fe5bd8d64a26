"""The reference's vanilla-CFR experiment protocol (src/experiments/run_vanilla_cfr_experiment.py:60-167) on the CUDA
solver: one run (vanilla CFR is deterministic) of `iterations` iterations -- the reference drives the trainer through its
private `_cfr_recursive(new_initial_state, player, 1.0, 1.0)` (:87-91), which the drop-in CFRTrainer accepts -- with 500
evaluation episodes vs a uniform-random opponent every `eval_interval` iterations, 5000 at the end, and (an addition)
the exploitability of the average policy at every evaluation point; results in the reference's file shapes.

    python -m scopa_b200.experiments.run_vanilla_cfr_experiment --out-dir experiments/results
"""
import argparse

from .. import pyspiel_compat as pyspiel
from ..envs import openspiel_mini_scopa  # noqa: F401  (registers the game)
from ..algorithms.vanilla_cfr import CFRTrainer
from . import tracker_output
from .run_mccfr_experiment import _evaluate


def run_vanilla_cfr_experiment(iterations=500, eval_interval=5, final_eval_episodes=5000):
    game = pyspiel.load_game("mini_scopa")
    trainer = CFRTrainer(game=game)
    run = {"eval_iterations": [], "eval_rewards": [], "eval_scopas_trained": [], "eval_scopas_random": [], "eval_scopa_diff": [],
           "exploitability_iterations": [], "exploitability_values": []}
    for t in range(iterations):
        for player_id in range(game.num_players()):                  # :87-91, the reference's own loop
            trainer._cfr_recursive(game.new_initial_state(), player_id, 1.0, 1.0)
        if (t + 1) % eval_interval == 0:
            r, st, sr = _evaluate(trainer.solver, 500, 7_000_003 + t, policy_kind=0)
            run["eval_iterations"].append(t + 1)
            run["eval_rewards"].append(r)
            run["eval_scopas_trained"].append(st)
            run["eval_scopas_random"].append(sr)
            run["eval_scopa_diff"].append(st - sr)
            run["exploitability_iterations"].append(t + 1)
            run["exploitability_values"].append(trainer.solver.exploitability(0))
    r, st, sr = _evaluate(trainer.solver, final_eval_episodes, 999_983, policy_kind=0)
    run.update({"final_reward": r, "final_scopa_trained": st, "final_scopa_random": sr, "final_scopa_diff": st - sr,
                "num_info_sets": len(trainer.info_set_map)})
    return run


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=500)
    ap.add_argument("--eval-interval", type=int, default=5)
    ap.add_argument("--final-episodes", type=int, default=5000)
    ap.add_argument("--out-dir", default="experiments/results")
    a = ap.parse_args()
    run = run_vanilla_cfr_experiment(a.iterations, a.eval_interval, a.final_episodes)
    files = tracker_output.save("MiniScopa_VanillaCFR", "Vanilla CFR", [run], a.out_dir)
    print(f"final reward vs random {run['final_reward']:+.4f}, exploitability {run['exploitability_values'][-1]:.4f}, "
          f"{run['num_info_sets']} infosets; wrote {len(files)} files under {a.out_dir}")


if __name__ == "__main__":
    main()
