"""Host-side result files: the JSON written for a set of runs has the keys and statistics of the
reference's shipped MiniScopa_MCCFR_data.json (tests/golden/../profiles copy of round 1 has the same shape), and the CSV
files carry the reference's headers (experiment_tracker.py:162-216)."""
import csv
import json
import os
import pickle

import numpy as np

from scopa_b200.experiments import tracker_output


def _run(seed, n=4, expl=False):
    rng = np.random.default_rng(seed)
    r = {"eval_iterations": [5 * (i + 1) for i in range(n)], "eval_rewards": rng.normal(1, .1, n).tolist(),
         "eval_scopas_trained": rng.uniform(.3, .5, n).tolist(), "eval_scopas_random": rng.uniform(.1, .2, n).tolist(),
         "final_reward": 1.1, "final_scopa_trained": .4, "final_scopa_random": .15, "final_scopa_diff": .25, "num_info_sets": 700}
    r["eval_scopa_diff"] = (np.array(r["eval_scopas_trained"]) - np.array(r["eval_scopas_random"])).tolist()
    if expl:
        r["exploitability_iterations"], r["exploitability_values"] = r["eval_iterations"], rng.uniform(0, 1, n).tolist()
    return r


def test_files_and_shapes(tmp_path):
    runs = [_run(0, expl=True), _run(1), _run(2)]
    files = tracker_output.save("MiniScopa_MCCFR", "MC-CFR", runs, str(tmp_path))
    names = sorted(os.path.basename(f) for f in files)
    assert names == sorted(["MiniScopa_MCCFR.pkl", "MiniScopa_MCCFR_data.json", "MiniScopa_MCCFR_run_1.csv",
                            "MiniScopa_MCCFR_run_1_exploitability.csv", "MiniScopa_MCCFR_run_2.csv",
                            "MiniScopa_MCCFR_run_3.csv", "MiniScopa_MCCFR_statistics.csv"])
    d = json.load(open(tmp_path / "MiniScopa_MCCFR_data.json"))
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "..", "profiles", "MiniScopa_MCCFR_data_r01.json")))
    assert set(d) == set(ref) and set(d["statistics"]) == set(ref["statistics"])
    assert set(d["statistics"]["final_metrics"]) == set(ref["statistics"]["final_metrics"])
    assert set(d["runs"][1]) == set(ref["runs"][0])
    rew = np.array([r["eval_rewards"] for r in runs])
    assert np.allclose(d["statistics"]["rewards"]["mean"], rew.mean(0)) and np.allclose(d["statistics"]["rewards"]["std"], rew.std(0))
    rows = list(csv.reader(open(tmp_path / "MiniScopa_MCCFR_run_2.csv")))
    assert rows[0] == ["Iteration", "Reward", "Scopas_Trained", "Scopas_Random", "Scopa_Diff"] and len(rows) == 5
    assert float(rows[1][1]) == runs[1]["eval_rewards"][0]
    rows = list(csv.reader(open(tmp_path / "MiniScopa_MCCFR_statistics.csv")))
    assert rows[0][:3] == ["Iteration", "Reward_Mean", "Reward_Std"] and len(rows) == 5
    rows = list(csv.reader(open(tmp_path / "MiniScopa_MCCFR_run_1_exploitability.csv")))
    assert rows[0] == ["Iteration", "Exploitability"]
    assert len(pickle.load(open(tmp_path / "MiniScopa_MCCFR.pkl", "rb"))) == 3


def test_single_run_has_no_statistics(tmp_path):
    files = tracker_output.save("MiniScopa_VanillaCFR", "Vanilla CFR", [_run(3, expl=True)], str(tmp_path))
    assert not any(f.endswith("_statistics.csv") for f in files)
    d = json.load(open(tmp_path / "MiniScopa_VanillaCFR_data.json"))
    assert "statistics" not in d and d["runs"][0]["exploitability_values"]
