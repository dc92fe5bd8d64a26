#!/usr/bin/env python
"""profiles/summarise_capture.py -- turn an exported `ncu --set full` report into the entry bench.py reads.

    ncu -i gpurun_out/mccfr_r02b.ncu-rep --page raw --csv > profiles/mccfr_r02b_raw.csv
    python profiles/summarise_capture.py mccfr_headline profiles/mccfr_r02b_raw.csv --pairs 151552 \
        --sources scopa_b200/csrc/ms_static_walk.cuh scopa_b200/csrc/ms_solver.cu scopa_b200/csrc/ms_tree_walk.cuh

writes / updates profiles/captures.json[name] with the figures bench.py's roofline objects quote (instructions per
traversal pair, DRAM bytes per launch, pipe utilisations) and the sha256 of the kernel's source files AT CAPTURE TIME:
bench.py recomputes that hash and marks the capture `stale` when the sources have changed since, instead of printing
old profiler numbers as if they described the running kernel.
"""
import argparse
import csv
import hashlib
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0,
        "Ghz": 1e9, "Mhz": 1e6}


def sha16(files):
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("raw_csv", nargs="?")
    ap.add_argument("--row", type=int, default=0, help="which captured launch of the file")
    ap.add_argument("--pairs", type=float, default=None, help="traversal pairs the captured launch processed")
    ap.add_argument("--sources", nargs="+", required=True)
    ap.add_argument("--sm-clock-mhz", type=float, default=None, help="SM clock assumed by bench.py's issue-slot peak (default: measured in the capture)")
    ap.add_argument("--cas-stall-share-pct", type=float, default=None, help="share of stall samples on the fp64 atomicAdd lines (source page)")
    ap.add_argument("--note", default=None)
    ap.add_argument("--source-sha16", default=None,
                    help="hash of --sources recorded ON THE GPU BOX when the capture was taken (profiles/prof_*.sh write it with "
                         "--hash-only); default: the hash of the files as they are now")
    ap.add_argument("--hash-only", action="store_true", help="print the hash of --sources and exit")
    a = ap.parse_args()
    if a.hash_only:
        print(sha16(a.sources))
        return
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units, row = rows[0], rows[1], rows[2 + a.row]

    def get(metric, default=None):
        for i, h in enumerate(hdr):
            if h == metric:
                try:
                    return float(row[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                except ValueError:
                    return row[i]
        return default

    inst = get("smsp__inst_executed.sum")
    clock = get("sm__cycles_elapsed.avg.per_second")
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    except Exception:
        commit = None
    e = {
        "kernel": get("Kernel Name"), "file": os.path.relpath(os.path.abspath(a.raw_csv), ROOT), "commit": commit,
        "source_files": a.sources, "source_sha16": a.source_sha16 or sha16(a.sources),
        "duration_us_under_ncu": (get("gpu__time_duration.sum") or 0) * 1e6,
        "warp_inst_per_launch": inst, "sm_clock_mhz_in_capture": clock / 1e6 if clock else None,
        "sm_clock_mhz_assumed": a.sm_clock_mhz or (clock / 1e6 if clock else 1965.0),
        "dram_bytes_per_launch": (get("dram__bytes_read.sum") or 0) + (get("dram__bytes_write.sum") or 0),
        "issue_slots_active_pct": get("sm__inst_issued.avg.pct_of_peak_sustained_active"),
        "smem_wavefronts_pct_of_peak": get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "alu_pipe_pct": get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": get("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_active_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": get("launch__registers_per_thread"),
        "threads_per_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "cas_stall_share_pct": a.cas_stall_share_pct, "note": a.note,
    }
    if a.pairs:
        e["traversal_pairs_in_capture"] = a.pairs
        e["warp_inst_per_traversal_pair"] = inst / a.pairs
        e["thread_inst_per_traversal_pair"] = inst * (e["threads_per_inst"] or 32.0) / a.pairs
    path = os.path.join(ROOT, "profiles", "captures.json")
    try:
        allc = json.load(open(path))
    except Exception:
        allc = {}
    allc[a.name] = e
    json.dump(allc, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(e, indent=1))


if __name__ == "__main__":
    main()
