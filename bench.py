#!/usr/bin/env python
"""bench.py -- throughput of the Miniscopa hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mccfr|rollout]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

BASELINE.json's metric is double-barrelled ("MCCFR infoset-node updates/sec & env steps/sec"), so the one
JSON line carries both:
  * primary (metric/value/e2e/roofline/cpu_baseline): MCCFR infoset-node updates/s, config 3/5 of
    BASELINE.json -- a step = one batched MCCFR iteration on the seed-42 deal: `--trav` traversals per
    player per GPU against the frozen table, one exchange of the slot-aligned delta buffer when N > 1,
    then table += delta.  Traversal ids are global (rank-offset), so the union of all ranks' work is
    independent of N ("weak" scaling: per-GPU work is fixed).
  * "env": env steps/s, config 2 of BASELINE.json -- 1 M concurrent random-policy games per GPU, a step =
    one fused rollout launch (8 plies per game) over deals already resident in HBM; its own e2e
    (seeds on the host -> actions + rewards on the host), roofline and CPU baseline.
`--workload rollout` swaps which of the two is reported as the primary metric.

Layout of a run.  The two sections above are COLLECTIVE-SYMMETRIC: every rank executes the same sequence of
collectives, and they are the only sections that run when N > 1.  Everything else (other estimators, the
exploitability curves, the step-granular env API, vanilla CFR, multi-deal MCCFR, 40-card Scopa, SDCFR, the CPU
baselines) is single-GPU reporting: it runs only when N == 1, never issues a collective, and each section is
wrapped so that a failure shows up as {"error": ...} under its key instead of taking the line down.

--impl reference times the reference's CPU implementation of the path on the box's host cores, same metric and
config: oracle/_ref (the unmodified reference, byte-compiled by oracle/make_ref.py; single-threaded Python, one
process per core) when it is there, else the C restatement oracle/ms_oracle.c ("port") on all cores.
"""
import argparse
import datetime
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic bytes per unit of work
BYTES_PER_UPDATE_FP64 = 203.7      # 172 updates x 136 B + 291 opponent lookups x 40 B per reference iteration
BYTES_PER_ENV_STEP = 34.0          # 16 B state load + 1 B action + 16 B state store (+ rewards on the last ply)
FALLBACK_HBM_GBS = 6650.0
UPDATES_PER_PAIR, VISITS_PER_PAIR = 172, 703     # the reference estimator's recursion shape on a 4+4-card deal


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def kernel_source_sha(files):
    """sha256 over the kernel sources a committed ncu capture describes: a capture whose recorded hash differs from the
    tree's is reported as stale instead of being printed as if it described the running kernel."""
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(ROOT, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_capture(name):
    """profiles/captures.json: per kernel, figures read off a committed `ncu --set full` report (written by
    profiles/summarise_capture.py from the exported raw CSV), with the kernel sources' hash at capture time."""
    try:
        with open(os.path.join(ROOT, "profiles", "captures.json")) as f:
            cap = json.load(f)[name]
        cap = dict(cap)
        cap["stale"] = kernel_source_sha(cap["source_files"]) != cap["source_sha16"]
        return cap
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a thread
    (nvidia-smi -lms as a fallback), reported as the median SM clock under load and the set of reasons seen."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.sm, self.reasons, self.smax, self._stop = [], set(), None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 5 ms polling"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 20"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bind_to_gpu_numa_node(index):
    """One process per GPU: run on the CPUs NVML names as local to that GPU, so that the pinned host buffers of the end-to-end
    paths are first touched on the GPU's own NUMA node and eight ranks do not push their PCIe traffic through one socket's
    memory.  -> the number of CPUs bound to, or None (NVML absent, an empty mask, affinity not permitted: nothing changes)."""
    if os.environ.get("SCOPA_B200_BENCH_AFFINITY", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def headline_config(args, world, collective=None):
    """`config` of the JSON line: the same object for the CUDA arm and for --impl reference."""
    cfg = {"workload": "BASELINE.json configs[2]/[4]: MCCFR (the reference's estimator, mc_cfr.py:37-86), seed-42 deal, "
                       f"{args.trav} traversals per player per GPU per iteration against a frozen table, fp64 table",
           "traversals_per_step": 2 * args.trav * world,
           "parallelism": (f"dp{world}: traversals sharded by id, one exchange of the slot-aligned fp64 delta buffer per iteration"
                           if world > 1 else "single GPU"),
           "l2": "256 MiB flush between timed steps (the working set is on-chip anyway)", "philox_seed": args.seed}
    if collective is not None:
        cfg["collective"] = collective
    return cfg


def env_config(args, world):
    return {"workload": f"BASELINE.json configs[1]: {args.games} concurrent random-policy games per GPU, 8 plies each, "
                        "deals resident in HBM", "l2": "256 MiB flush between timed steps", "games_per_gpu": args.games}


class Cx:
    """What every section needs: rank/world, device, the L2 flush and the cross-rank reductions.  `flush_l2(sync=True)`
    issues a collective and may only be called from collective-symmetric code; everything else flushes locally."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.world, self.local = dist_env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- scopa_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa_cpus = bind_to_gpu_numa_node(self.local) if self.world > 1 else None
        if self.world > 1:
            # a collective mismatch must fail in minutes, not hang until the driver's limit
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=180))
        self.hbm_gbs, self.peak_src = load_peaks()
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        self.sync_word = torch.zeros(1, device=self.dev)
        self.K, self.W = args.steps, args.warmup

    def flush_l2(self, sync=False):
        """Evict L2 between timed steps.  sync=True (symmetric sections only) also lines the ranks up again with a
        stream-ordered all-reduce of one word, so a rank's timed step does not include another rank's flush."""
        self.flush_buf.fill_(1)
        if sync and self.world > 1:
            self.dist.all_reduce(self.sync_word)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def min_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN)

    def sum_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def events(self, n):
        ev = self.torch.cuda.Event
        return [(ev(enable_timing=True), ev(enable_timing=True)) for _ in range(n)]


# =========================================================================================== symmetric sections
def attach_peers_all_ranks(cx, sv):
    """Solver.attach_peers with every failure turned into a group decision: no rank is ever left alone in a collective.
    -> True when every rank mapped every peer's buffers."""
    import ctypes as C
    from scopa_b200 import _lib
    dist, torch = cx.dist, cx.torch
    handle, offs = (C.c_ubyte * 64)(), (C.c_uint64 * 3)()
    ok = 1.0
    try:
        with torch.cuda.device(cx.dev):
            _lib.check(sv.lib.ms_solver_ipc_export(sv.h, handle, offs))
    except Exception as e:
        print(f"[rank {cx.rank}] ipc export failed: {e}", file=sys.stderr)
        ok = 0.0
    if cx.min_over_ranks(ok) < 1.0:
        return False
    everyone = [None] * cx.world
    dist.all_gather_object(everyone, (bytes(handle), [int(o) for o in offs]))
    try:
        handles = b"".join(h for h, _ in everyone)
        flat = (C.c_uint64 * (3 * cx.world))(*[o for _, oo in everyone for o in oo])
        with torch.cuda.device(cx.dev):
            _lib.check(sv.lib.ms_solver_ipc_attach(sv.h, cx.rank, cx.world, handles, flat))
        sv._delta_t = None
        sv._peers = True
    except Exception as e:          # e.g. no peer access between the GPUs of this box
        print(f"[rank {cx.rank}] peer attach failed: {e}", file=sys.stderr)
        ok = 0.0
    return cx.min_over_ranks(ok) >= 1.0


def time_mccfr(cx, sv, exchange, B, seed, it0, warm, steps):
    """`warm` untimed + `steps` timed iterations of {batch kernel, exchange, apply}; -> dict of device times (ms) and
    counters, max over ranks for the step time.  `exchange(sv)` is one of the forms in section_mccfr; the string "fused"
    selects ms_mccfr_batch_peers (traversals + peer exchange + apply in ONE launch)."""
    torch = cx.torch
    from scopa_b200 import _lib
    rank, world = cx.rank, cx.world
    fused = exchange == "fused"

    def batch(s, first):
        if fused:
            s.mccfr_batch_peers(2, B, philox_seed=seed, first_trav=first)
        else:
            s.mccfr_batch(2, B, philox_seed=seed, first_trav=first)

    if fused:
        exchange = lambda s: None
    for i in range(warm):
        batch(sv, ((it0 + i) * world + rank) * B)
        exchange(sv)
    sv.counters(reset=True)
    ev, kev = cx.events(steps), cx.events(steps)
    cx.barrier()
    l0 = _lib.launch_count()
    wall0 = time.perf_counter()
    for i in range(steps):
        cx.flush_l2(sync=True)
        ev[i][0].record()
        kev[i][0].record()
        batch(sv, ((it0 + warm + i) * world + rank) * B)
        kev[i][1].record()
        exchange(sv)
        ev[i][1].record()
    cx.barrier()
    wall = time.perf_counter() - wall0
    launches = _lib.launch_count() - l0
    ms_total = cx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / steps
    cnt = sv.counters()
    return {"ms_total": ms_total, "ms_kernel": ms_kernel, "launches": int(launches), "wall": wall, "cnt": cnt,
            "updates_all": cx.sum_over_ranks(cnt["updates"]), "visits_all": cx.sum_over_ranks(cnt["visits"]),
            "edges_all": cx.sum_over_ranks(cnt["env_steps"]), "next_it": it0 + warm + steps}


def section_mccfr(cx, sampler):
    """Headline: batched MCCFR on the seed-42 deal (configs 3 / 5).  Collective-symmetric."""
    torch, dist, args = cx.torch, cx.dist, cx.args
    from scopa_b200 import _lib
    from scopa_b200.solver import Solver
    K, W, B, world, rank = cx.K, cx.W, args.trav, cx.world, cx.rank
    sv = Solver(seed=42, device=cx.dev)
    S = sv.n_slots
    delta = sv.delta_tensor()

    def ex_local(s):
        s.mccfr_apply()

    def ex_nccl(s):
        dist.all_reduce(delta)          # one all-reduce of the fp64 delta buffer per iteration (NVLink / NVSwitch)
        s.mccfr_apply()

    def ex_p2p(s):
        s.apply_peers()

    # the exchange of the headline number: the peer-memory exchange kernel (auto / p2p; verified on hardware by
    # tests/test_gpu_multigpu.py; the fastest of the forms at N = 2 / 4 / 8, profiles/README.md r02g), NCCL all-reduce +
    # apply when the peer mapping is unavailable or --collective nccl; every other form is timed beside it
    sv_p2p, p2p_ok, p2p_note = None, False, None
    if world > 1:
        sv_p2p = Solver(seed=42, device=cx.dev)
        p2p_ok = attach_peers_all_ranks(cx, sv_p2p)
        if not p2p_ok:
            p2p_note = "peer attach failed on at least one rank (stderr has the reason)"
            if args.collective in ("p2p", "p2p_fused"):
                raise SystemExit("--collective p2p requested but peer memory is unavailable")
    if world == 1:
        collective, exchange, hsv = "none", ex_local, sv
    elif args.collective == "p2p_fused" and p2p_ok:
        collective, exchange, hsv = "p2p_fused", "fused", sv_p2p
    elif args.collective in ("auto", "p2p") and p2p_ok:
        collective, exchange, hsv = "p2p", ex_p2p, sv_p2p
    else:
        collective, exchange, hsv = "nccl", ex_nccl, sv

    sampler.start()
    r = time_mccfr(cx, hsv, exchange, B, args.seed, 0, W, K)
    if collective.startswith("p2p") and cx.max_over_ranks(float(hsv.peer_error())):
        # a rank gave up on a peer (bounded wait): the tables are not those of a multi-GPU run -- redo with NCCL
        print(f"[rank {rank}] peer exchange reported an error word; falling back to NCCL", file=sys.stderr)
        collective, exchange, hsv, p2p_ok = "nccl", ex_nccl, sv, False
        p2p_note = "peer exchange timed out during the headline run (a rank waited 2 s for a peer)"
        r = time_mccfr(cx, hsv, exchange, B, args.seed, 0, W, K)
    value = r["updates_all"] / (r["ms_total"] * 1e-3)
    upd_per_launch = r["cnt"]["updates"] / K

    # e2e: the whole solver state crosses PCIe every step (table in from pinned host memory, table out)
    h_reg = torch.zeros((S, 4), dtype=torch.float64).pin_memory()
    h_str = torch.zeros((S, 4), dtype=torch.float64).pin_memory()
    reg0, str0, _ = hsv.export()
    h_reg.copy_(torch.from_numpy(reg0)); h_str.copy_(torch.from_numpy(str0))
    lib = _lib.load()
    e2e_steps = max(3, min(K, 10))
    cx.barrier()
    hsv.counters(reset=True)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        _lib.check(lib.ms_solver_import_table(hsv.h, h_reg.data_ptr(), h_str.data_ptr(), hsv._stream()))
        if exchange == "fused":
            hsv.mccfr_batch_peers(2, B, philox_seed=args.seed, first_trav=((r["next_it"] + i) * world + rank) * B)
        else:
            hsv.mccfr_batch(2, B, philox_seed=args.seed, first_trav=((r["next_it"] + i) * world + rank) * B)
            exchange(hsv)
        _lib.check(lib.ms_solver_export_table(hsv.h, None, None, None, h_reg.data_ptr(), h_str.data_ptr(), None,
                                              hsv._stream()))
    cx.barrier()
    e2e_s = cx.max_over_ranks(time.perf_counter() - t0)
    e2e_updates = cx.sum_over_ranks(hsv.counters()["updates"])
    e2e = {"value": e2e_updates / e2e_s, "unit": "infoset-node updates/s",
           "h2d_bytes_per_step": 2 * S * 4 * 8, "d2h_bytes_per_step": 2 * S * 4 * 8,
           "what": "ms_solver_import_table (pinned host) + mccfr batch + exchange + apply + ms_solver_export_table"}

    # N > 1: the same step with each exchange form, so the line names what limits the scaling curve
    exchange_ms = None
    if world > 1:
        exchange_ms = {"headline": collective}
        it = r["next_it"] + e2e_steps
        short = max(5, min(K, 10))
        for name, fn, s in (("no_exchange_floor", ex_local, sv), ("nccl", ex_nccl, sv), ("p2p", ex_p2p, sv_p2p),
                            ("p2p_fused", "fused", sv_p2p)):
            if name == collective:
                exchange_ms[name + "_ms_per_step"] = r["ms_total"] / K
                continue
            if name.startswith("p2p"):
                if not p2p_ok:
                    exchange_ms[name] = p2p_note
                    continue
            rr = time_mccfr(cx, s, fn, B, args.seed, it, 3, short)
            it = rr["next_it"]
            exchange_ms[name + "_ms_per_step"] = rr["ms_total"] / short
            if name.startswith("p2p"):
                err = s.peer_error()
                bad = cx.max_over_ranks(float(err))
                if bad:
                    exchange_ms[name] = f"peer exchange reported error word {int(bad)} (a rank timed out waiting for a peer)"
                    exchange_ms.pop(name + "_ms_per_step", None)
        exchange_ms["note"] = ("ms per iteration, max over ranks, same batch size; nccl = batch kernel + NCCL all-reduce + apply "
                               "kernel; p2p = batch kernel + ms_mccfr_apply_peers (peer-memory reads, rank-ordered sum, apply in "
                               "one kernel); p2p_fused = ms_mccfr_batch_peers, ONE launch: the last CTA to finish its traversals "
                               "does the exchange; no_exchange_floor applies only the LOCAL delta (not a valid multi-GPU "
                               "result): the gap to it is what the exchange costs")

    obj = {
        "metric": "mccfr_infoset_node_updates_per_sec", "value": value, "unit": "infoset-node updates/s",
        "ms_per_step": r["ms_total"] / K, "e2e": e2e, "gpu_launches": r["launches"],
        "node_visits_per_sec": r["visits_all"] / (r["ms_total"] * 1e-3),
        "tree_edges_per_sec_inside_mccfr": r["edges_all"] / (r["ms_total"] * 1e-3),
        "roofline": mccfr_roofline(cx, upd_per_launch, r["ms_kernel"]),
        "update_composition": {
            "per_traversal_pair": {"updates": UPDATES_PER_PAIR, "node_visits": VISITS_PER_PAIR,
                                   "updates_at_one_card_infosets": 120, "updates_with_regret_arithmetic": 52},
            "note": "counted as the reference counts them (mc_cfr.py:83-84 runs at every traverser node): 120 of the 172 updates "
                    "per traversal pair are one-card infosets whose regret delta is exactly 0 and whose update is one "
                    "strategy-count increment; the deterministic last two plies are played once and accounted twice, as "
                    "the reference's two recursive calls would visit them (node_visits_per_sec includes those visits)"},
        "config": headline_config(args, world, collective),
        "wall_s": r["wall"],
    }
    if exchange_ms is not None:
        obj["exchange"] = exchange_ms
    return obj, sv


def mccfr_roofline(cx, upd_per_launch, ms_kernel, capture_name="mccfr_headline"):
    """The headline kernel keeps the table, the tree and every delta in shared memory: DRAM traffic is a few KB per
    launch, so the bound is on chip.  `frac` = issue-slot utilisation = warp instructions per launch (from the committed
    ncu capture of THIS kernel source, scaled by the traversals of the launch: the estimator's recursion shape is
    data-independent) / (kernel time measured in this run x 148 SMs x 4 schedulers x the SM clock).  The HBM-equivalent
    figure SURVEY 8(d) prescribes is kept as a side note."""
    hbm_equiv = upd_per_launch * BYTES_PER_UPDATE_FP64 / (ms_kernel * 1e-3) / 1e9
    cap = load_capture(capture_name)
    roof = {"bound": "issue", "kernel": cap["kernel"] if cap else capture_name, "kernel_ms": ms_kernel,
            "hbm_equivalent": {"achieved_gbs": hbm_equiv, "peak_gbs": cx.hbm_gbs, "ratio": hbm_equiv / cx.hbm_gbs,
                               "peak_source": cx.peak_src,
                               "note": "203.7 algorithmic B/update x updates per launch / kernel time: what an HBM-resident "
                                       "table would have to move; NOT a fraction of a binding resource (the table is "
                                       "shared-memory resident)"}}
    if cap is None:
        roof.update({"achieved": None, "peak": None, "unit": "G warp-instructions/s", "frac": None, "traffic": None,
                     "note": "no committed capture of this kernel (profiles/captures.json)"})
        return roof
    cap_keys = ("file", "commit", "source_sha16", "stale", "issue_slots_active_pct", "smem_wavefronts_pct_of_peak", "alu_pipe_pct",
                "fp64_pipe_pct", "warp_inst_per_traversal_pair", "cas_stall_share_pct")
    if cap["stale"]:
        roof.update({"achieved": None, "peak": None, "unit": "G warp-instructions/s", "frac": None, "traffic": None,
                     "capture": {k: cap.get(k) for k in cap_keys},
                     "note": "the committed capture describes an OLDER version of the kernel sources (hash mismatch): no "
                             "instruction count is quoted for the running kernel; re-capture with profiles/prof_*.sh"})
        return roof
    pairs = upd_per_launch / UPDATES_PER_PAIR
    winst = cap["warp_inst_per_traversal_pair"] * pairs
    sm_hz = cap["sm_clock_mhz_assumed"] * 1e6
    peak = 148 * 4 * sm_hz
    ach = winst / (ms_kernel * 1e-3)
    roof.update({"achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-instructions/s", "frac": ach / peak,
                 "traffic": cap.get("dram_bytes_per_launch"),
                 "capture": {k: cap.get(k) for k in cap_keys},
                 "note": "issue-slot roofline: achieved = warp instructions per launch (ncu smsp__inst_executed.sum of the "
                         "committed capture per traversal pair x pairs in this launch; the recursion shape is data-independent) / "
                         "kernel time measured in this run; peak = 148 SMs x 4 schedulers x 1 warp instruction per cycle at "
                         "the maximum SM clock; traffic = DRAM bytes per launch in the capture (table and tree staging only)"})
    return roof


def section_env(cx):
    """Config 2: 1 M concurrent random-policy games per GPU.  Collective-symmetric."""
    torch, args = cx.torch, cx.args
    from scopa_b200 import _lib
    from scopa_b200.batch import BatchedMiniScopa, rollout_random_host
    K, W, G, rank, world, dev = cx.K, cx.W, args.games, cx.rank, cx.world, cx.dev
    seeds_np = np.arange(1 + rank * G, 1 + (rank + 1) * G, dtype=np.int64)
    b = BatchedMiniScopa(dev).reset(seeds_np)
    actions = torch.empty((G, 8), dtype=torch.uint8, device=dev)
    rewards = torch.empty((G, 2), dtype=torch.float32, device=dev)
    for i in range(W):
        b.rollout_random(philox_seed=args.seed, game_offset=rank * G, actions=actions, rewards=rewards)
    rev = cx.events(K)
    cx.barrier()
    l0 = _lib.launch_count()
    for i in range(K):
        cx.flush_l2(sync=True)
        rev[i][0].record()
        b.rollout_random(philox_seed=args.seed + i, game_offset=rank * G, actions=actions, rewards=rewards)
        rev[i][1].record()
    cx.barrier()
    launches = _lib.launch_count() - l0
    ms_total = cx.max_over_ranks(sum(a.elapsed_time(c) for a, c in rev))
    value = 8.0 * G * world * K / (ms_total * 1e-3)
    kernel_ms = sum(a.elapsed_time(c) for a, c in rev) / K
    roof_ach = 8.0 * G * BYTES_PER_ENV_STEP / (kernel_ms * 1e-3) / 1e9
    # e2e: seeds in pinned host memory -> deal + rollout on device -> actions + rewards in pinned host memory
    h_seeds = torch.from_numpy(seeds_np).pin_memory()
    h_act = torch.empty((G, 8), dtype=torch.uint8).pin_memory()
    h_rew = torch.empty((G, 2), dtype=torch.float32).pin_memory()
    rollout_random_host(h_seeds, args.seed, rank * G, h_act, h_rew)
    e2e_steps = max(3, min(K, 10))
    cx.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        rollout_random_host(h_seeds, args.seed + i, rank * G, h_act, h_rew)
    cx.barrier()
    e2e_s = cx.max_over_ranks(time.perf_counter() - t0)
    cap = load_capture("rollout_kernel")
    return {
        "metric": "env_steps_per_sec", "value": value, "unit": "env steps/s", "ms_per_step": ms_total / K,
        "e2e": {"value": 8.0 * G * world * e2e_steps / e2e_s, "unit": "env steps/s",
                "h2d_bytes_per_step": 8 * G, "d2h_bytes_per_step": 16 * G,
                "what": "ms_rollout_random_host: reset(seed) + 8 steps per game, host buffers in and out"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": roof_ach, "peak": cx.hbm_gbs, "unit": "GB/s", "frac": roof_ach / cx.hbm_gbs,
                     "traffic": (cap or {}).get("dram_bytes_per_launch"), "capture": cap,
                     "kernel": "rollout_kernel", "kernel_ms": kernel_ms, "peak_source": cx.peak_src,
                     "note": "against the step-granular 34 B/step figure (SURVEY 8(d)); the fused kernel itself moves "
                             "36 B/game (4.5 B/step) and is integer-issue bound (ALU pipe 85 % in the committed capture), "
                             "not HBM bound; env_step_api.step_kernel is the HBM-bound member of the family"},
        "config": env_config(args, world),
    }


# =========================================================================================== single-GPU reporting
def guarded(fn, *a, **kw):
    """An auxiliary section must not take the headline line down with it."""
    try:
        return fn(*a, **kw)
    except Exception as e:
        import traceback
        traceback.print_exc(file=sys.stderr)
        return {"error": repr(e)}


def section_restep(cx, sv):
    """the same estimator re-stepping the env at every node (mode 3): step(), capture resolution, legal list, infoset
    key and hash probe at every node instead of walking the enumerated tree -- kept as the comparison point"""
    torch, args, K = cx.torch, cx.args, cx.K
    Br = 148 * 768 * 3
    for i in range(2):
        sv.mccfr_batch(2, Br, philox_seed=args.seed, first_trav=i * Br, mode=3)
        sv.mccfr_apply()
    sv.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        sv.mccfr_batch(2, Br, philox_seed=args.seed, first_trav=(2 + i) * Br, mode=3)
        sv.mccfr_apply()
    e1.record()
    torch.cuda.synchronize()
    rc = sv.counters(reset=True)
    rms = e0.elapsed_time(e1)
    return {"metric": "mccfr_infoset_node_updates_per_sec", "unit": "infoset-node updates/s",
            "value": rc["updates"] / (rms * 1e-3), "ms_per_step": rms / K,
            "env_steps_per_sec_inside_mccfr": rc["env_steps"] / (rms * 1e-3), "kernel": "mccfr_batch_kernel",
            "config": {"workload": f"same estimator, mode 3 (env re-stepped at every node), {Br} traversals per player per step"}}


def section_tree_walk(cx):
    """the generic tree-walking kernel (mode 4: shared-memory frames, any root) beside the headline"""
    torch, args, K = cx.torch, cx.args, cx.K
    from scopa_b200.solver import Solver
    sv = Solver(seed=42, device=cx.dev)
    Bt = 148 * 1024 * 3
    for i in range(2):
        sv.mccfr_batch(2, Bt, philox_seed=args.seed, first_trav=i * Bt, mode=4)
        sv.mccfr_apply()
    sv.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        sv.mccfr_batch(2, Bt, philox_seed=args.seed, first_trav=(2 + i) * Bt, mode=4)
        sv.mccfr_apply()
    e1.record()
    torch.cuda.synchronize()
    rc = sv.counters(reset=True)
    rms = e0.elapsed_time(e1)
    return {"metric": "mccfr_infoset_node_updates_per_sec", "unit": "infoset-node updates/s",
            "value": rc["updates"] / (rms * 1e-3), "ms_per_step": rms / K, "kernel": "mccfr_tree_kernel",
            "config": {"workload": f"same estimator, mode 4 (round 1's headline kernel: DFS frames in shared memory), {Bt} "
                                   "traversals per player per step"}}


def section_es(cx):
    """textbook external sampling (opt-in estimator) + BASELINE.json configs[2] exploitability curves"""
    torch, args, K, W = cx.torch, cx.args, cx.K, cx.W
    from scopa_b200.solver import Solver
    es_sv = Solver(seed=42, device=cx.dev)
    Bes = 148 * 1024 * 2
    for i in range(W):
        es_sv.mccfr_batch(2, Bes, philox_seed=args.seed, first_trav=i * Bes, mode=1)
        es_sv.mccfr_apply()
    es_sv.counters(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        es_sv.mccfr_batch(2, Bes, philox_seed=args.seed, first_trav=(W + i) * Bes, mode=1)
        es_sv.mccfr_apply()
    e1.record()
    torch.cuda.synchronize()
    es_cnt = es_sv.counters()
    es_s = e0.elapsed_time(e1) * 1e-3
    return {"estimator": "external sampling (Lanctot et al. 2009), not in the reference", "kernel": "mccfr_es_tree_kernel",
            "traversals_per_sec": 2.0 * Bes * K / es_s, "regret_updates_per_sec": es_cnt["updates"] / es_s,
            "node_visits_per_sec": es_cnt["visits"] / es_s,
            "exploitability_after": {"traversals_per_player": (W + K) * Bes, "value": es_sv.exploitability(1)},
            "note": "the reference's own estimator plateaus near 0.49 exploitability on this deal"}


def section_schedules(cx):
    """BASELINE.json configs[2] ("10 M traversals, exploitability vs iteration") for the schedules the path offers:
    B = 1 is the reference's own schedule (in-place kernel: every update visible to the next node visit), larger B are
    frozen-sigma batches (the headline runs at B = --trav).  Same estimator, same deal, empty table at the start; reports
    exploitability after 10^4..10^7 traversals per player, wall time, and the time to reach exploitability 0.3 (if it
    does).  SURVEY H5: a batch of B traversals uses one strategy for all B, so the curve depends on B."""
    torch, args = cx.torch, cx.args
    from scopa_b200.solver import Solver
    targets = [t for t in (10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7) if t <= args.curve_max]
    out = {"estimator": "reference (mc_cfr.py:37-86)", "targets_traversals_per_player": targets, "schedules": {}}
    if not targets:
        return out
    for B in (1, 4096, 65536, args.trav):
        sv = Solver(seed=42, device=cx.dev)
        done, pts, t_03 = 0, [], None
        my_targets = targets
        if B == 1:
            # one device thread runs the reference's schedule: time 1000 iterations, keep the targets that fit ~12 s
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sv.mccfr_inplace(1000, philox_seed=args.seed, first_iter=0)
            torch.cuda.synchronize()
            per_it = (time.perf_counter() - t0) / 1000
            sv = Solver(seed=42, device=cx.dev)
            my_targets = [t for t in targets if t * per_it <= 12.0] or [min(targets[0], max(1000, int(12.0 / per_it)))]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for target in my_targets:
            if B == 1:
                sv.mccfr_inplace(target - done, philox_seed=args.seed, first_iter=done)
                done = target
            else:
                while done < target:
                    sv.mccfr_batch(2, B, philox_seed=args.seed, first_trav=done)
                    sv.mccfr_apply()
                    done += B
            ex = sv.exploitability(1)       # synchronises
            el = time.perf_counter() - t0
            pts.append([done, ex, el])
            if t_03 is None and ex <= 0.3:
                t_03 = el
        out["schedules"]["in_place_B1" if B == 1 else f"B{B}"] = {
            "traversals_per_player_vs_exploitability_vs_wall_s": pts, "wall_s_to_exploitability_0.3_at_a_checkpoint": t_03}
        if B == 1:
            out["schedules"]["in_place_B1"]["us_per_iteration"] = per_it * 1e6
            out["schedules"]["in_place_B1"]["note"] = ("one device thread; targets beyond ~12 s of run time are left out "
                                                       "(the reference itself takes 62-67 ms per iteration)")
        del sv
    # the textbook estimator at one batch size for reference
    sv = Solver(seed=42, device=cx.dev)
    done, pts = 0, []
    t0 = time.perf_counter()
    for target in targets:
        while done < target:
            sv.mccfr_batch(2, 4096, philox_seed=args.seed, first_trav=done, mode=1)
            sv.mccfr_apply()
            done += 4096
        pts.append([done, sv.exploitability(1), time.perf_counter() - t0])
    out["external_sampling_B4096"] = {"traversals_per_player_vs_exploitability_vs_wall_s": pts}
    out["note"] = ("exploitability by the device best-response sweep (restated OpenSpiel definition, parity unpinned); wall time "
                   "includes the best-response launches at the checkpoints; the reference at B = 1 plateaus near 0.49 "
                   "(tests/golden/policies_eval.json), frozen-sigma batches of the same estimator plateau lower")
    return out


def section_inplace(cx):
    """MCCFRTrainer.iteration() as the reference runs it (mc_cfr.py:88-92): one traversal per player, every update
    visible to the next visit.  (a) one run: the one-thread in-place kernel; (b) the reference's experiment protocol
    (run_mccfr_experiment.py:195-202: independent runs) as ONE launch: ms_mccfr_inplace_many, one warp per run."""
    torch, args = cx.torch, cx.args
    from scopa_b200.solver import Solver, mccfr_inplace_many
    sv = Solver(seed=42, device=cx.dev)
    sv.mccfr_inplace(50, philox_seed=1)
    torch.cuda.synchronize()
    iters = 1000
    t0 = time.perf_counter()
    sv.mccfr_inplace(iters, philox_seed=1, first_iter=50)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"single_run": {"iterations": iters, "us_per_iteration": dt / iters * 1e6, "iterations_per_sec": iters / dt,
                          "updates_per_sec": iters * UPDATES_PER_PAIR / dt, "kernel": "mccfr_inplace_tree_kernel",
                          "note": "the reference takes 62-67 ms per iteration on one CPU core (BASELINE.md section 2)"}}
    for runs, it_m in ((10, 500), (148 * 4, 500), (148 * 32, 100)):
        tabs = mccfr_inplace_many(sv, runs, 20, philox_seed0=100)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tabs = mccfr_inplace_many(sv, runs, it_m, philox_seed0=100)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[f"many_runs_{runs}"] = {"runs": runs, "iterations_per_run": it_m, "ms": dt * 1e3,
                                    "run_iterations_per_sec": runs * it_m / dt,
                                    "updates_per_sec": runs * it_m * UPDATES_PER_PAIR / dt,
                                    "kernel": "mccfr_inplace_many_kernel"}
        del tabs
    out["note"] = ("reference-semantics throughput: every run is bit-identical to a solo ms_mccfr_inplace run with the same "
                   "Philox seed (tests/test_gpu_solver.py); 10 runs x 500 iterations is the reference's published protocol")
    return out


def section_step_api(cx):
    """ms_step round-trips the 16-byte state through HBM: the one kernel family here whose real bound IS the HBM
    roofline (33 B/step + 8 B rewards + 1 B done = 42 B moved per step with all outputs requested)"""
    torch, args, K, W, dev = cx.torch, cx.args, cx.K, cx.W, cx.dev
    from scopa_b200.batch import BatchedMiniScopa
    NS = args.step_states
    bs = BatchedMiniScopa(dev).reset(np.arange(1, NS + 1, dtype=np.int64))
    _, ordered0, _ = bs.legal_actions()
    acts_u8 = ordered0[:, 0].contiguous()                  # first legal card of every game
    st_backup = bs.states.clone()
    rew_s = torch.empty((NS, 2), dtype=torch.float32, device=dev)
    done_t = torch.empty((NS,), dtype=torch.uint8, device=dev)
    for i in range(W):
        bs.step(acts_u8, rewards=rew_s, done=done_t)
    stev = cx.events(K)
    torch.cuda.synchronize()
    for i in range(K):
        bs.states.copy_(st_backup)
        cx.flush_l2()
        stev[i][0].record()
        bs.step(acts_u8, rewards=rew_s, done=done_t)
        stev[i][1].record()
    torch.cuda.synchronize()
    step_ms = sum(a.elapsed_time(c) for a, c in stev) / K
    gbs = 42.0 * NS / (step_ms * 1e-3) / 1e9
    cap = load_capture("step_kernel")            # dram__bytes of one launch (16 M states), from the committed capture of this source
    traffic = cap.get("dram_bytes_per_launch") if cap and not cap["stale"] and NS == 16_000_000 else None
    return {"kernel": "step_kernel", "env_steps_per_sec": NS / (step_ms * 1e-3), "kernel_ms": step_ms,
            "bytes_per_step": 42, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / cx.hbm_gbs,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": cx.hbm_gbs, "unit": "GB/s", "frac": gbs / cx.hbm_gbs,
                         "traffic": traffic, "algorithmic_bytes_per_launch": 42.0 * NS, "peak_source": cx.peak_src,
                         "capture": ({k: cap.get(k) for k in ("file", "commit", "source_sha16", "stale", "duration_us_under_ncu")} if cap else None)},
            "note": f"{NS} states ({16 * NS >> 20} MiB, larger than L2), one ply per launch, state + action in, "
                    "state + rewards + done out"}


def section_cfr(cx):
    torch, dev, rank = cx.torch, cx.dev, cx.rank
    from scopa_b200.solver import Solver, cfr_iterate_many
    cfr_sv = Solver(seed=42, device=dev)
    cfr_sv.cfr_iterate(5)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    cfr_sv.cfr_iterate(200)
    c1.record()
    torch.cuda.synchronize()
    obj = {"kernel": "cfr_kernel", "us_per_iteration": c0.elapsed_time(c1) * 1e3 / 200,
           "node_visits_per_sec": 2 * 2229 * 200 / (c0.elapsed_time(c1) * 1e-3),
           "note": "BASELINE.json configs[0]: vanilla CFR on the seed-42 deal, 200 iterations in one launch, float64, "
                   "bit-identical to the reference's tables; the reference takes 389 ms per iteration on one CPU core"}
    # throughput mode: 148 independent deals, one CTA (= one SM) per deal, one launch
    many = [Solver(seed=1000 + rank * 148 + i, device=dev) for i in range(148)]
    cfr_iterate_many(many, 3)
    torch.cuda.synchronize()
    reps = []
    for _ in range(3):          # one launch of a few ms: repeated, the spread is reported (3.5 - 6.5 ms seen across runs)
        c0.record()
        cfr_iterate_many(many, 100)
        c1.record()
        torch.cuda.synchronize()
        reps.append(c0.elapsed_time(c1))
    nodes = sum(m_.n_nodes for m_ in many)
    ms = float(np.median(reps))
    obj["many_deals"] = {"deals": 148, "iterations": 100, "ms": ms, "ms_repeats": reps,
                         "deal_iterations_per_sec": 148 * 100 / (ms * 1e-3),
                         "node_visits_per_sec": 2.0 * nodes * 100 / (ms * 1e-3),
                         "note": "ms_cfr_iterate_many: seeds 1000.., float64, each deal's tables identical to a solo run; median of 3 launches"}
    return obj


def section_atomics(cx, mccfr_value):
    """atomic roofline: measured on the box"""
    import ctypes
    from scopa_b200 import _lib
    peaks = (ctypes.c_double * 3)()
    _lib.check(_lib.load().ms_debug_atomic_peaks(peaks, _lib.stream_ptr()))
    pairs_per_s = mccfr_value / float(UPDATES_PER_PAIR)          # traversal pairs per second on this GPU
    # per traversal pair the headline kernel issues at most 66 shared-memory fp64 atomic adds (nl - 1 difference
    # accumulators per traverser node with more than one action: (3 + 5*2 + 20*1) per player; zero addends are skipped)
    # into lane-private columns, and 172 u32 increments (update counts)
    return {"measured_peaks_per_sec": {"smem_f64_atomic_add": peaks[0], "smem_u32_atomic_add": peaks[1],
                                       "global_red_f64_l2_resident": peaks[2]},
            "mccfr_smem_f64_atomics_per_sec_upper_bound": 66.0 * pairs_per_s, "mccfr_smem_u32_atomics_per_sec": 172.0 * pairs_per_s,
            "frac_of_smem_f64_peak": 66.0 * pairs_per_s / peaks[0], "frac_of_smem_u32_peak": 172.0 * pairs_per_s / peaks[1],
            "note": "microbenchmark: 148 CTAs x 768 threads, pseudo-random addresses over a 738x4 table; shared-memory "
                    "fp64 atomicAdd compiles to an ATOMS.CAST.SPIN.64 compare-and-swap loop, global fp64 to REDG.E.ADD.F64.  "
                    "The microbenchmark's random addresses collide inside a warp; the kernel's lane-private columns do not, "
                    "so its adds are cheaper than the benchmark's (profiles/README.md, r02)"}


def section_multideal(cx):
    """Multi-deal MCCFR: the regime SURVEY 8(d) names as the one where the memory system is the
    bound: one infoset table for 65 536 deals (5.1 M stored infosets, 0.66 GB of 128-byte lines, far beyond L2) in HBM."""
    import ctypes
    torch, args, K, dev = cx.torch, cx.args, cx.K, cx.dev
    from scopa_b200 import _lib, multideal
    lg = args.md_log2_capacity
    md = multideal.MultiDealSolver(np.arange(1, args.md_deals + 1), log2_capacity=lg, device=dev)
    nb = args.md_trav
    for i in range(12):                                  # fill the table: inserts are rare afterwards
        md.mccfr_batch(nb, philox_seed=args.seed, first_trav=i * nb)
        md.apply()
    md.counters(reset=True)
    mev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
    l0 = _lib.launch_count()
    mev[0].record()
    for i in range(K):
        md.mccfr_batch(nb, philox_seed=args.seed, first_trav=(12 + i) * nb)
        mev[2 * i + 1].record()
        md.apply()
        mev[2 * i + 2].record()
    torch.cuda.synchronize()
    md_launches = _lib.launch_count() - l0
    mc = md.counters()
    t_trav = sum(mev[2 * i].elapsed_time(mev[2 * i + 1]) for i in range(K)) / K
    t_app = sum(mev[2 * i + 1].elapsed_time(mev[2 * i + 2]) for i in range(K)) / K
    # ceiling: dependent 64-byte reads of random 128-byte lines over a buffer the size of the STORED infosets
    lines_lg = max(10, int(np.ceil(np.log2(max(mc["infosets"], 1)))))
    rp = (ctypes.c_double * 3)()
    table_bytes = md.table_bytes
    # the deal-blocked form on the same table: one deal per CTA visit, 3072 traversal pairs per visit staged on chip
    VIS, PPV = 148, 12288
    tb0 = time.perf_counter()
    md.mccfr_blocked(VIS, pairs_per_visit=PPV, philox_seed=args.seed, first_visit=0)     # first call builds the deal records
    md.apply()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - tb0
    for i in range(1, 4):
        md.mccfr_blocked(VIS, pairs_per_visit=PPV, philox_seed=args.seed, first_visit=i * VIS)
        md.apply()
    md.counters(reset=True)
    bev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
    bev[0].record()
    for i in range(K):
        md.mccfr_blocked(VIS, pairs_per_visit=PPV, philox_seed=args.seed, first_visit=(4 + i) * VIS)
        bev[2 * i + 1].record()
        md.apply()
        bev[2 * i + 2].record()
    torch.cuda.synchronize()
    bc = md.counters()
    b_ms = bev[0].elapsed_time(bev[2 * K]) / K
    bk_ms = sum(bev[2 * i].elapsed_time(bev[2 * i + 1]) for i in range(K)) / K
    n_info = int(bc["infosets"])
    # deal-blocked traffic: per visit the deal's stored infosets are read once and their deltas written once (a few tens of
    # KB per visit of thousands of traversals): like the one-deal kernel it is bound by issue slots, not by the table
    mdb_obj = {"metric": "mccfr_infoset_node_updates_per_sec", "unit": "infoset-node updates/s",
               "value": bc["updates"] / K / (b_ms * 1e-3), "ms_per_step": b_ms, "ms_traverse": bk_ms, "infosets": n_info,
               "first_call_s_incl_describing_all_deals": build_s, "kernel": "md_blocked_kernel",
               "roofline": mccfr_roofline(cx, bc["updates"] / K, bk_ms, "md_blocked_kernel"),
               "config": {"workload": f"same table and estimator, deal-blocked: {VIS} visits x {PPV} traversal pairs per step, one "
                                      "deal per CTA visit staged in shared memory (node records, strategies, lane-private "
                                      "accumulators: the one-deal solver's static walk), table read once and written once per visit"}}
    del md
    torch.cuda.empty_cache()
    _lib.check(_lib.load().ms_debug_random_access_peaks(lines_lg, rp, _lib.stream_ptr()))
    # per traversal pair (data-independent recursion shape of the estimator, 4+4-card deals): 163 lookups of
    # stored infosets (player 0 traversal: 26 own + 85 opponent nodes with >1 card; player 1: 26 + 26) and 52
    # update groups (regret-delta REDs + visit count on the line just read)
    touches = (163.0 + 52.0) * nb / (t_trav * 1e-3)
    cap = load_capture("md_mccfr_kernel")
    md_obj = {"metric": "mccfr_infoset_node_updates_per_sec", "unit": "infoset-node updates/s",
              "value": mc["updates"] / K / ((t_trav + t_app) * 1e-3),
              "stored_infoset_updates_per_sec": 52.0 * nb / ((t_trav + t_app) * 1e-3),
              "node_visits_per_sec": mc["visits"] / K / ((t_trav + t_app) * 1e-3),
              "ms_traverse": t_trav, "ms_apply": t_app, "gpu_launches": int(md_launches),
              "infosets": int(mc["infosets"]), "load_factor": mc["infosets"] / float(1 << lg), "table_bytes": int(table_bytes),
              "roofline": {"bound": "hbm", "kind": "random 128-byte line transactions", "achieved": touches / 1e9,
                           "peak": rp[0] / 1e9, "unit": "G lines/s", "frac": touches / rp[0],
                           "peak_source": f"ms_debug_random_access_peaks over 2^{lines_lg} lines (the stored infosets' footprint), "
                                          "148 x 768 threads, dependent 64-byte reads, measured in this run",
                           "independent_reads_peak": rp[1] / 1e9, "red_x4_lines_peak": rp[2] / 1e9,
                           "traffic": (cap or {}).get("dram_bytes_per_launch"), "capture": cap, "kernel": "md_mccfr_kernel",
                           "note": "line touches = 163 lookups + 52 update groups per traversal pair; hot (shallow) "
                                   "infosets hit in L2 (ncu: 75 % of sectors), so the DRAM-resident ceiling is not the "
                                   "binding one yet: the kernel is latency-bound (profiles/README.md section 6)"},
              "config": {"workload": f"MCCFR (reference estimator) over {args.md_deals} deals (seeds 1..), chance-sampled root, "
                                     f"{nb} traversal pairs per step, one fp64 infoset table of 2^{lg} x 128 B in HBM",
                         "note": "updates are reference-equivalent (172 per traversal pair); infosets with one card in hand "
                                 "(120 of the 172) are not stored -- their strategy is the constant [1.0]"}}
    return {"traversal_per_thread": md_obj, "deal_blocked": mdb_obj}


def section_multideal_sharded(cx):
    """Multi-deal MCCFR across the GPUs of the box (last sentence): ONE infoset table sharded over the ranks by a
    hash of the key; every rank runs its share of an iteration's visits with md_blocked_kernel, which gathers the frozen
    regrets from and sends its deltas to the owners' shards through peer memory (the "all-to-all of deltas" is inside the
    kernel); then a barrier, the apply step on the own shard, a barrier.  Collective-symmetric: every failure is turned
    into a group decision before the next collective."""
    torch, args, K, W = cx.torch, cx.args, cx.K, cx.W
    from scopa_b200 import multideal
    world, rank = cx.world, cx.rank
    owner_bits = int(np.ceil(np.log2(world)))
    lg = max(16, args.md_log2_capacity - owner_bits)          # the same total capacity as the single-GPU section
    VIS, PPV = 148 * 4, 3072                                   # per rank and iteration: 4 visits per SM
    seeds = np.arange(1, args.md_deals + 1)
    res = {"metric": "mccfr_infoset_node_updates_per_sec", "unit": "infoset-node updates/s", "kernel": "md_blocked_kernel",
           "config": {"workload": f"MCCFR (reference estimator) over {args.md_deals} deals, deal-blocked, {VIS} visits x {PPV} traversal "
                                  f"pairs per GPU per iteration; one infoset table sharded over {world} GPUs by a hash of the key "
                                  f"(2^{lg} x 128 B per shard), regrets gathered / deltas sent through peer memory inside the kernel",
                      "exchange": "peer memory (CUDA IPC over NVLink / NVSwitch): remote loads + fp64 RED.ADDs in md_blocked_kernel, "
                                  "two flag barriers per iteration (md_peer_barrier_kernel)"}}
    sh = pv = None
    err = None
    try:
        sh = multideal.MultiDealSolver(seeds, log2_capacity=lg, device=cx.dev)
        pv = multideal.MultiDealSolver(seeds, log2_capacity=min(args.md_log2_capacity, 25), device=cx.dev)   # same per-GPU work, private table
    except Exception as e:
        err = f"create: {e}"
    if cx.min_over_ranks(0.0 if err else 1.0) < 1.0:
        res["error"] = err or "another rank could not create its table"
        return res
    try:
        sh.attach_peers()
    except Exception as e:                                     # raised on every rank (attach_peers agrees on the outcome)
        res["error"] = f"attach_peers: {e}"
        return res

    def run(sv, sharded, it):
        if sharded:
            sv.iterate_blocked(VIS * world, PPV, philox_seed=args.seed, first_visit=VIS * world * it)
        else:
            sv.mccfr_blocked(VIS, PPV, philox_seed=args.seed, first_visit=VIS * it)
            sv.apply()

    out = {}
    try:
        for name, sv, sharded in (("sharded", sh, True), ("private", pv, False)):
            for i in range(max(W, 3)):
                run(sv, sharded, i)
            sv.counters(reset=True)
            torch.cuda.synchronize()
            ev = cx.events(K)
            if sharded:
                sv.barrier()
            for i in range(K):
                ev[i][0].record()
                run(sv, sharded, max(W, 3) + i)
                ev[i][1].record()
            torch.cuda.synchronize()
            out[name] = (sum(a.elapsed_time(b) for a, b in ev) / K, sv.counters())
        perr = sh.peer_error()
        if perr:
            err = f"peer barrier: rank {perr - 1} did not arrive"
    except Exception as e:
        err = f"timed loop: {e}"
    if cx.min_over_ranks(0.0 if err else 1.0) < 1.0:
        res["error"] = err or "another rank failed"
        return res
    ms_sh, ms_pv = cx.max_over_ranks(out["sharded"][0]), cx.max_over_ranks(out["private"][0])
    upd_all = cx.sum_over_ranks(out["sharded"][1]["updates"]) / K
    upd_pv = out["private"][1]["updates"] / K
    info_all = cx.sum_over_ranks(out["sharded"][1]["infosets"])
    res.update({"value": upd_all / (ms_sh * 1e-3), "ms_per_step": ms_sh, "infosets": int(info_all),
                "private_table_per_gpu": {"value": upd_pv / (ms_pv * 1e-3), "ms_per_step": ms_pv,
                                          "note": "the same per-GPU work on an unsharded table of this GPU alone (no exchange): "
                                                  "the floor the sharded iteration is compared with"},
                "efficiency_vs_private": (upd_all / (ms_sh * 1e-3)) / (world * upd_pv / (ms_pv * 1e-3))})
    del sh, pv
    torch.cuda.empty_cache()
    return res


def section_full(cx):
    """40-card Scopa rollouts"""
    torch, args, K, dev = cx.torch, cx.args, cx.K, cx.dev
    from scopa_b200 import _lib
    from scopa_b200 import full as fs
    FG = args.full_games
    fb = fs.BatchedFullScopa(dev).reset(np.arange(1, FG + 1, dtype=np.int64))
    for i in range(2):
        fb.rollout_random(philox_seed=args.seed, game_offset=i * FG)
    fev = cx.events(K)
    for i in range(K):
        cx.flush_l2()                     # local flush only: this section runs on one rank
        fev[i][0].record()
        fb.rollout_random(philox_seed=args.seed, game_offset=(2 + i) * FG)
        fev[i][1].record()
    torch.cuda.synchronize()
    f_ms = sum(a.elapsed_time(c) for a, c in fev) / K
    h_seeds = np.arange(1, FG + 1, dtype=np.int64)
    # e2e: seeds in pinned host memory -> deck + deal + 36-ply rollout on device -> actions + rewards in pinned host memory
    p_seeds = torch.from_numpy(h_seeds).pin_memory()
    p_act = torch.empty((FG, fs.PLIES), dtype=torch.uint8).pin_memory()
    p_rew = torch.empty((FG, 2), dtype=torch.float32).pin_memory()
    flib = _lib.load()
    _lib.check(flib.ms_full_rollout_random_host(p_seeds.data_ptr(), FG, args.seed, 0, p_act.data_ptr(), p_rew.data_ptr()))
    f_reps = max(3, min(K, 10))
    t0 = time.perf_counter()
    for i in range(f_reps):
        _lib.check(flib.ms_full_rollout_random_host(p_seeds.data_ptr(), FG, args.seed + i, 0, p_act.data_ptr(), p_rew.data_ptr()))
    f_e2e = (time.perf_counter() - t0) / f_reps
    cpu_full = None
    if not args.no_cpu:
        from oracle import ms_oracle as ora              # CPU leg: the checker timed as the baseline
        ncpu = os.cpu_count() or 1
        ora.full_rollout_random(h_seeds[:2000], args.seed)
        t0 = time.perf_counter()
        ora.full_rollout_random(h_seeds[:200_000], args.seed)
        dt = time.perf_counter() - t0
        cpu_full = {"value": 200_000 * fs.PLIES / dt, "unit": "env steps/s", "cores": ncpu, "kind": "port",
                    "sample": "200 000 games x 36 plies, OpenMP over all host threads"}
    ach = FG * fs.PLIES * 65.0 / (f_ms * 1e-3) / 1e9
    fcap = load_capture("full_rollout_kernel")
    return {"metric": "env_steps_per_sec", "unit": "env steps/s", "value": FG * fs.PLIES / (f_ms * 1e-3), "ms_per_step": f_ms,
            "e2e": {"value": FG * fs.PLIES / f_e2e, "unit": "env steps/s", "h2d_bytes_per_step": 8 * FG,
                    "d2h_bytes_per_step": (fs.PLIES + 8) * FG,
                    "what": "ms_full_rollout_random_host: FullDeck(seed) + deal + 36 plies per game, pinned host buffers in and out, "
                            "H2D / kernels / D2H pipelined over three streams"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": cx.hbm_gbs, "unit": "GB/s", "frac": ach / cx.hbm_gbs,
                         "traffic": (fcap.get("dram_bytes_per_launch") if fcap and not fcap["stale"] and FG == 1_000_000 else None),
                         "capture": ({k: fcap.get(k) for k in ("file", "commit", "source_sha16", "stale", "alu_pipe_pct",
                                                               "issue_slots_active_pct", "duration_us_under_ncu")} if fcap else None),
                         "kernel": "full_rollout_kernel", "peak_source": cx.peak_src,
                         "note": "against the step-granular figure for this game, 65 B/step (32 B state load + 1 B action "
                                 "+ 32 B state store); the fused kernel keeps the state in registers and moves "
                                 "108 B/game = 3 B/step, so it is integer-issue bound like the Miniscopa rollout"},
            "cpu_baseline": cpu_full,
            "config": {"workload": f"{FG} concurrent random-policy games of 40-card Scopa (FullScopaEnv), 36 plies each, "
                                   "deals resident in HBM", "l2": "256 MiB flush between timed steps"}}


def section_sdcfr(cx, sv):
    """SDCFR traversal (config 4)"""
    torch, args, K, dev = cx.torch, cx.args, cx.K, cx.dev
    from scopa_b200 import _lib
    from scopa_b200 import sdcfr as sd
    T = args.sd_trav
    torch.manual_seed(1234)
    blobs = [(torch.randn(sd.NET_FLOATS, device=dev) * 0.15).contiguous() for _ in range(2)]   # random-init nets
    trv = sd.Traverser(sv.root_words, sv.hand_order, device=dev)
    sd_obj = None
    for prec, pname in ((sd.TENSOR_CORE, "bf16_tcgen05"), (sd.FP32, "fp32_cuda_cores")):
        for i in range(2):
            trv.run(i & 1, blobs, T, philox_seed=args.seed, first_trav=0, precision=prec)
        sev = cx.events(K)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        for i in range(K):
            cx.flush_l2()
            sev[i][0].record()
            trv.run(0, blobs, T, philox_seed=args.seed, first_trav=i * T, precision=prec)
            trv.run(1, blobs, T, philox_seed=args.seed, first_trav=i * T, precision=prec)
            sev[i][1].record()
        torch.cuda.synchronize()
        sd_launches = _lib.launch_count() - l0
        sd_ms = sum(a.elapsed_time(c) for a, c in sev)
        inf = 187.0 * T * K                      # 105 + 82 advantage-net inferences per traversal pair
        o = {"inferences_per_sec": inf / (sd_ms * 1e-3), "traversals_per_sec": 2.0 * T * K / (sd_ms * 1e-3),
             "ms_per_step": sd_ms / K, "gpu_launches": int(sd_launches),
             "algorithmic_tflops": inf * 27136.0 / (sd_ms * 1e-3) / 1e12}
        if sd_obj is None:
            sd_obj = {"metric": "sdcfr_advantage_net_inferences_per_sec", "unit": "inferences/s",
                      "config": {"workload": f"BASELINE.json configs[3]: SDCFR external-sampling traversals, {T} per player per "
                                             "GPU per step, level-batched frontier inference, random-init nets 34-128-64-16",
                                 "note": "whole traversal timed (8 forward levels incl. env steps + sampling, 8 backward levels), "
                                         "not the MMA alone; 27136 algorithmic FLOP per inference (un-padded)"}}
        sd_obj[pname] = o
    sd_obj["value"] = sd_obj["bf16_tcgen05"]["inferences_per_sec"]
    try:       # tensor roofline of the SDCFR section: un-padded algorithmic FLOP of whole traversals over the sustained bf16 peak
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            _pk = json.load(f)
        tf_peak, tf_src = float(_pk.get("bf16_tflops_sustained") or _pk["bf16_tflops"]), "measured, sustained (MEASURED_PEAKS.json)"
    except Exception:
        tf_peak, tf_src = 2250.0, "fallback: nominal dense bf16 (B200_PROFILING.md)"
    sd_obj["roofline"] = {"bound": "tensor", "achieved": sd_obj["bf16_tcgen05"]["algorithmic_tflops"], "peak": tf_peak,
                          "unit": "TFLOP/s", "frac": sd_obj["bf16_tcgen05"]["algorithmic_tflops"] / tf_peak, "traffic": None,
                          "kernel": "sd_level_mlp_kernel<1>", "peak_source": tf_src,
                          "note": "numerator = 27 136 FLOP x inferences of WHOLE traversals (env steps, sampling, backward levels "
                                  "and sample emission included in the time), so this is a floor on the tensor-pipe share; "
                                  "K = 34 / N = 16 layers pad to MMA tiles; the kernel that holds the MMAs is sd_level_mlp_kernel<1> "
                                  "(its own tensor-pipe utilisation: `capture`, profiles/README.md R2.5)"}
    cap = load_capture("sd_level_mlp_kernel")
    if cap:
        keys = ("file", "commit", "source_sha16", "stale", "tensor_pipe_active_pct", "issue_slots_active_pct", "warps_active_pct",
                "duration_us_under_ncu", "registers_per_thread", "dram_bytes_per_launch", "note")
        sd_obj["roofline"]["capture"] = {k: cap.get(k) for k in keys}
        if not cap["stale"]:
            sd_obj["roofline"]["traffic"] = cap.get("dram_bytes_per_launch")
    sd_obj["train"] = guarded(bench_sd_train, dev, _lib)
    sd_obj["drop_in_train"] = guarded(bench_sd_dropin, dev)
    return sd_obj


def bench_sd_train(dev, _lib, epochs=10, calls=20, rows=100_000):
    """AdvantageNetwork.train (deep_cfr.py:77-110), `epochs` optimiser steps per call on a full replay buffer:
    the fused kernel (one launch per call) beside the PyTorch step it replaces (about 40 launches and one
    device->host read per epoch).  Wall clock around synchronised calls: that is what DeepCFR.train waits for."""
    import torch
    from scopa_b200.algorithms.deep_cfr.deep_cfr import AdvantageNetwork
    torch.manual_seed(7)
    out = {"config": {"workload": f"{epochs} optimiser steps per train() call, minibatch 128 drawn from a replay buffer of "
                                  f"{rows} rows, net 34-128-64-16 fp32, masked MSE + clip-norm 1.0 + Adam 5e-4",
                      "flop_per_step": 3 * 2 * 128 * 13568 - 2 * 128 * 34 * 128}}
    feat = (torch.rand((rows, 34), device=dev) < 0.25).float()
    mask = (torch.rand((rows, 16), device=dev) < 0.2).float()
    target = (torch.rand((rows, 16), device=dev) * 2 - 1) * mask
    def one(name):
        adv = AdvantageNetwork(34, 16, device=dev, optimizer=name)
        adv.buffer.add_batch(feat, target, mask)
        for _ in range(3):
            adv.train(batch_size=128, epochs=epochs)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        for _ in range(calls):
            loss = adv.train(batch_size=128, epochs=epochs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        o = {"us_per_step": 1e6 * dt / (calls * epochs), "ms_per_call": 1e3 * dt / calls, "last_loss": loss,
             "our_launches_per_call": (_lib.launch_count() - l0) / calls}
        if name != "torch":     # the kernel alone (CUDA events around the launches, minibatch rows drawn beforehand)
            idx = adv._sample_rows(128, epochs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(calls):
                adv._fused.step(adv.buffer.feat, adv.buffer.target, adv.buffer.mask, idx)
            e1.record()
            torch.cuda.synchronize()
            o["kernel_us_per_step"] = 1e3 * e0.elapsed_time(e1) / (calls * epochs)
        return o

    for name in ("fused", "torch", "fused-cluster"):
        out[name] = guarded(one, name)
    out["speedup_vs_torch_step"] = out["torch"]["us_per_step"] / out["fused"]["us_per_step"]
    out["note"] = ("sd_train_kernel: one CTA keeps the 13 776 weights, the minibatch and all activations in shared memory "
                   "for every epoch of the call; bit-identical to the host emulation of the same source "
                   "(tests/emu), which is pinned to torch on the CPU")
    return out


def bench_sd_dropin(dev, iterations=20):
    """DeepCFR.train through the drop-in API in the reference's own configuration (BASELINE.json configs[3]: one
    traversal per player per iteration, 10 optimiser epochs, evaluation against random every 5 iterations with 50
    episodes; deep_cfr.py:431-490), wall clock per iteration: the default build of the class (PyTorch optimiser,
    episode-loop evaluation) beside optimizer="fused" + device_eval=True."""
    import torch
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401  (registers "mini_scopa")
    from scopa_b200.algorithms.deep_cfr import DeepCFR
    game = pyspiel.load_game("mini_scopa")
    out = {"config": {"workload": f"DeepCFR(game).train(iterations={iterations}, advantage_epochs=10, eval_freq=5), 50 evaluation "
                                  "episodes, 1 traversal per player per iteration, fp32 inference"}}
    for name, kw in (("default", {}), ("fused_optimizer_device_eval", {"optimizer": "fused", "device_eval": True}),
                     ("fused_cluster_optimizer_device_eval", {"optimizer": "fused-cluster", "device_eval": True})):
        torch.manual_seed(3)
        np.random.seed(3)
        d = DeepCFR(game, 2, dev, seed=3, **kw)
        d.train(iterations=5, advantage_epochs=10, eval_freq=5, eval_episodes=50)        # warm-up (incl. one evaluation)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.train(iterations=iterations, advantage_epochs=10, eval_freq=5, eval_episodes=50)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"ms_per_iteration": 1e3 * dt / iterations, "iterations_per_sec": iterations / dt,
                     "last_eval_reward_vs_random": d.training_history["eval_rewards"][-1]}
    out["speedup"] = out["default"]["ms_per_iteration"] / out["fused_optimizer_device_eval"]["ms_per_iteration"]
    out["note"] = "the reference takes about 130 ms per iteration (traversal pair) on one CPU core (BASELINE.md section 2)"
    return out


# =========================================================================================== ours
def run_ours(args):
    cx = Cx(args)
    rank, world = cx.rank, cx.world
    sampler = ClockSampler(cx.local)
    mccfr_obj, sv = section_mccfr(cx, sampler)        # starts the sampler
    env_obj = section_env(cx)
    clocks = sampler.stop()
    extras = {}
    cpu_mccfr = cpu_env = None
    if world == 1:
        # single-GPU reporting: no collectives below this line
        if not args.no_extras:
            t_ex = time.perf_counter()

            def extra(name, fn, *a):
                # the default run must end within minutes whatever a section does: past the budget the rest is skipped
                if args.only and name != args.only:
                    return
                if time.perf_counter() - t_ex > args.extras_budget_s:
                    extras[name] = {"skipped": f"extras wall budget of {args.extras_budget_s:.0f} s used up"}
                    return
                t0 = time.perf_counter()
                extras[name] = guarded(fn, *a)
                if isinstance(extras[name], dict):
                    extras[name]["section_wall_s"] = time.perf_counter() - t0

            extra("mccfr_in_place", section_inplace, cx)
            extra("mccfr_schedules", section_schedules, cx)
            extra("mccfr_tree_walk", section_tree_walk, cx)
            extra("mccfr_restep", section_restep, cx, sv)
            extra("mccfr_external_sampling", section_es, cx)
            extra("env_step_api", section_step_api, cx)
            extra("cfr", section_cfr, cx)
            extra("atomics", section_atomics, cx, mccfr_obj["value"])
            if args.md_deals > 0:
                extra("mccfr_multi_deal", section_multideal, cx)
            if args.full_games > 0:
                extra("full_scopa", section_full, cx)
            extra("sdcfr", section_sdcfr, cx, sv)
        if not args.no_cpu:
            cpu = guarded(cpu_baselines, args, 8.0)
            if isinstance(cpu, dict) and "error" in cpu:
                cpu_mccfr = cpu_env = cpu
            else:
                cpu_mccfr, cpu_env = cpu
    if world > 1:
        if args.md_deals > 0 and not args.no_extras:
            extras["mccfr_multi_deal_sharded"] = section_multideal_sharded(cx)      # collective-symmetric
        cx.dist.barrier()
        cx.dist.destroy_process_group()
    if rank != 0:
        return None
    mccfr_obj["cpu_baseline"], env_obj["cpu_baseline"] = cpu_mccfr, cpu_env
    primary, secondary = (mccfr_obj, env_obj) if args.workload == "mccfr" else (env_obj, mccfr_obj)
    line = {
        "metric": primary["metric"], "value": primary["value"], "unit": primary["unit"], "n_gpus": world, "steps": cx.K,
        "warmup": cx.W, "ms_per_step": primary["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if primary is mccfr_obj else "u32", "data": "synthetic",
        "config": primary["config"], "e2e": primary["e2e"], "gpu_launches": primary["gpu_launches"],
        "roofline": primary["roofline"], "cpu_baseline": primary["cpu_baseline"], "clocks": clocks,
        ("env" if primary is mccfr_obj else "mccfr"): secondary,
        "collective": mccfr_obj["config"].get("collective"),
    }
    if world > 1:
        line["host_binding"] = ({"cpus_per_rank": cx.numa_cpus, "note": "each rank runs on the CPUs NVML reports as local to its GPU "
                                 "(pinned host buffers of the e2e paths are first touched on that NUMA node)"}
                                if cx.numa_cpus else {"cpus_per_rank": None, "note": "not bound (NVML affinity unavailable or disabled)"})
    for k in ("node_visits_per_sec", "update_composition", "exchange"):
        if primary is mccfr_obj and k in mccfr_obj:
            line[k] = mccfr_obj[k]
    line.update(extras)
    return line


# ------------------------------------------------------------------------------------------- CPU side
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def have_reference():
    return os.path.exists(os.path.join(REF_DIR, "src", "algorithms", "mc_cfr.refc"))


class ReferencePool:
    """`procs` persistent single-threaded Python processes, each with the UNMODIFIED reference imported (oracle/_ref:
    byte-compiled from /root/reference/src by oracle/make_ref.py, behind the import shims of oracle/stubs).  The
    reference has no multi-core path, so "all host cores" = one independent run per core.  Interpreter start-up, imports
    and one warm-up iteration happen before `run` is first timed."""

    def __init__(self, procs):
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "time_reference.py"), "serve"]
        self.ps = [subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                                    bufsize=1) for _ in range(procs)]
        for p in self.ps:
            line = p.stdout.readline()
            if not line or not json.loads(line).get("ready"):
                self.close()
                raise RuntimeError("oracle/time_reference.py did not start (oracle/_ref unusable with this interpreter?)")

    def run(self, what, n):
        for p in self.ps:
            p.stdin.write(f"{what} {n}\n")
            p.stdin.flush()
        return [json.loads(p.stdout.readline()) for p in self.ps]

    def close(self):
        for p in self.ps:
            try:
                p.stdin.write("quit\n"); p.stdin.flush(); p.stdin.close()
                p.wait(timeout=5)
            except Exception:
                p.kill()


def time_reference_python(what, n):
    pool = ReferencePool(1)
    try:
        return pool.run(what, n)
    finally:
        pool.close()


def cpu_baselines(args, sample_seconds, threads=None):
    """The CPU leg beside the GPU numbers (rank 0, N = 1): the C restatement (oracle/ms_oracle.c, "port") on all host
    cores, and -- when oracle/_ref is there -- the unmodified Python reference on one core."""
    from oracle import ms_oracle as ora
    ora.build()
    cores = threads or (os.cpu_count() or 1)
    # calibrate, then run a bounded sample
    t0 = time.perf_counter()
    u, v = ora.mccfr_bench(200, cores, args.seed)
    rate = u / max(time.perf_counter() - t0, 1e-6)
    per_thread = max(200, int(rate * sample_seconds / 172.0 / cores))
    t0 = time.perf_counter()
    u, v = ora.mccfr_bench(per_thread, cores, args.seed)
    dt = time.perf_counter() - t0
    cpu_mccfr = {"value": u / dt, "unit": "infoset-node updates/s", "cores": cores, "kind": "port",
                 "sample": f"{per_thread} traversal pairs x {cores} independent workers ({u} updates, {dt:.2f} s); "
                           "oracle/ms_oracle.c, frozen-sigma batches",
                 "node_visits_per_sec": v / dt}
    t0 = time.perf_counter()
    tab = ora.Table()
    tab.cfr_train(20)
    cpu_mccfr["vanilla_cfr_ms_per_iteration_1core"] = (time.perf_counter() - t0) * 1e3 / 20
    seeds = np.arange(1, 200_001, dtype=np.int64)
    t0 = time.perf_counter()
    ora.rollout_random(seeds, args.seed, nthreads=cores)
    rate = 8 * len(seeds) / max(time.perf_counter() - t0, 1e-6)
    n = int(min(max(rate * sample_seconds / 8, 200_000), 20_000_000))
    seeds = np.arange(1, n + 1, dtype=np.int64)
    t0 = time.perf_counter()
    ora.rollout_random(seeds, args.seed, nthreads=cores)
    dt = time.perf_counter() - t0
    cpu_env = {"value": 8 * n / dt, "unit": "env steps/s", "cores": cores, "kind": "port",
               "sample": f"{n} games (reset(seed) + 8 random-policy steps each, {dt:.2f} s); oracle/ms_oracle.c"}
    if have_reference():
        try:
            m = time_reference_python("mccfr", 100)[0]
            c = time_reference_python("cfr", 20)[0]
            e = time_reference_python("env", 3000)[0]
            cpu_mccfr["reference_python"] = {
                "value": m["updates_per_sec"], "unit": "infoset-node updates/s", "cores": 1, "kind": "reference",
                "sample": f"MCCFRTrainer.iteration() x {m['n']} (mc_cfr.py:88-92), {m['seconds']:.2f} s, one process",
                "ms_per_iteration": m["ms_per_iteration"], "node_visits_per_sec": m["visits_per_sec"],
                "vanilla_cfr": {"ms_per_iteration": c["ms_per_iteration"], "sample": f"CFRTrainer.train({c['n']}) (vanilla_cfr.py:105-120)"},
                "host_cores": os.cpu_count()}
            cpu_env["reference_python"] = {
                "value": e["steps_per_sec"], "unit": "env steps/s", "cores": 1, "kind": "reference",
                "sample": f"MiniScopaEnv.reset(seed=g) + 8 random legal steps x {e['n']} games (mini_scopa_game.py:131-167), "
                          f"{e['seconds']:.2f} s, one process", "host_cores": os.cpu_count()}
        except Exception as ex:
            cpu_mccfr["reference_python"] = {"error": repr(ex)}
    else:
        cpu_mccfr["reference_python"] = {"unavailable": "oracle/_ref is absent (run oracle/make_ref.py where /root/reference exists)"}
    return cpu_mccfr, cpu_env


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores, same metric / config.
    oracle/_ref (the unmodified Python reference, one single-threaded process per core) when present and not
    --ref-kind port; else the C restatement on all cores."""
    rank, world, local = dist_env()
    if rank != 0:
        return None
    K, W = args.steps, args.warmup
    cores = os.cpu_count() or 1
    use_ref = have_reference() and args.ref_kind != "port"
    from oracle import ms_oracle as ora
    ora.build()
    if use_ref:
        # one "step" = every core runs `ref_iters` iterations of the unmodified MCCFRTrainer / `ref_py_games` env games
        it, games = args.ref_iters, args.ref_py_games
        pool = ReferencePool(cores)
        try:
            for _ in range(min(W, 2)):
                pool.run("mccfr", max(1, it // 4))
            tot_u = tot_v = 0.0
            t0 = time.perf_counter()
            for i in range(K):
                for r in pool.run("mccfr", it):
                    tot_u += r["updates"]; tot_v += r["visits"]
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            tot_s = 0.0
            for i in range(K):
                for r in pool.run("env", games):
                    tot_s += r["steps"]
            dt_env = time.perf_counter() - t1
        finally:
            pool.close()
        kind = "reference"
        sample_m = (f"MCCFRTrainer.iteration() x {it} in each of {cores} independent single-threaded processes per step "
                    "(unmodified reference, oracle/_ref; interpreter start-up, imports and warm-up outside the timed region)")
        sample_e = f"{games} MiniScopaEnv games (reset + 8 random steps) in each of {cores} processes per step"
        mccfr = {"metric": "mccfr_infoset_node_updates_per_sec", "value": tot_u / dt, "unit": "infoset-node updates/s",
                 "ms_per_step": dt / K * 1e3, "node_visits_per_sec": tot_v / dt}
        env = {"metric": "env_steps_per_sec", "value": tot_s / dt_env, "unit": "env steps/s", "ms_per_step": dt_env / K * 1e3}
    else:
        per_thread, games = args.ref_trav, args.ref_games
        seeds = np.arange(1, games + 1, dtype=np.int64)
        for _ in range(W):
            ora.mccfr_bench(max(50, per_thread // 10), cores, args.seed)
        t0 = time.perf_counter()
        tot_u = tot_v = 0
        for i in range(K):
            u, v = ora.mccfr_bench(per_thread, cores, args.seed + i)
            tot_u += u; tot_v += v
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        for i in range(K):
            ora.rollout_random(seeds, args.seed + i, nthreads=cores)
        dt_env = time.perf_counter() - t1
        kind = "port"
        sample_m = f"{per_thread} traversal pairs x {cores} workers per step (oracle/ms_oracle.c, C restatement of the reference)"
        sample_e = f"{games} games per step (oracle/ms_oracle.c)"
        mccfr = {"metric": "mccfr_infoset_node_updates_per_sec", "value": tot_u / dt, "unit": "infoset-node updates/s",
                 "ms_per_step": dt / K * 1e3, "node_visits_per_sec": tot_v / dt}
        env = {"metric": "env_steps_per_sec", "value": 8.0 * games * K / dt_env, "unit": "env steps/s",
               "ms_per_step": dt_env / K * 1e3}
    primary, secondary = (mccfr, env) if args.workload == "mccfr" else (env, mccfr)
    sample = sample_m if primary is mccfr else sample_e
    port_line = None
    if use_ref:        # the C port beside it, for scale (it is what cpu_baseline.kind "port" reports in the CUDA arm)
        u, v = ora.mccfr_bench(200, cores, args.seed)
        t0 = time.perf_counter()
        u, v = ora.mccfr_bench(args.ref_trav, cores, args.seed)
        port_line = {"value": u / (time.perf_counter() - t0), "unit": "infoset-node updates/s", "cores": cores, "kind": "port",
                     "sample": f"{args.ref_trav} traversal pairs x {cores} workers, oracle/ms_oracle.c"}
    line = {
        "impl": "reference", "metric": primary["metric"], "value": primary["value"], "unit": primary["unit"],
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": primary["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if primary is mccfr else "u32", "data": "synthetic",
        "config": headline_config(args, world) if primary is mccfr else env_config(args, world),
        "cpu_baseline": {"value": primary["value"], "unit": primary["unit"], "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": primary["value"], "unit": primary["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        ("env" if primary is mccfr else "mccfr"): secondary,
        "c_port_all_cores": port_line,
        "note": "the reference is single-threaded pure Python with no multi-core path: all host cores = one independent "
                "process per core (updates summed); each step is a bounded sample of the workload in `config`",
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mccfr", choices=["mccfr", "rollout"])
    ap.add_argument("--trav", type=int, default=1818624,
                    help="traversals per player per GPU per step (default: 12 full waves of 148 CTAs x 1024 threads, about half a "
                         "millisecond of traversal per step)")
    ap.add_argument("--games", type=int, default=1_000_000, help="concurrent games per GPU")
    ap.add_argument("--sd-trav", type=int, default=65536, help="SDCFR traversals per player per GPU per step")
    ap.add_argument("--step-states", type=int, default=16_000_000, help="states in the step-granular API measurement")
    ap.add_argument("--full-games", type=int, default=1_000_000, help="concurrent 40-card Scopa games (0 = skip)")
    ap.add_argument("--md-deals", type=int, default=65536, help="deals in the multi-deal MCCFR section (0 = skip)")
    ap.add_argument("--md-log2-capacity", type=int, default=26, help="multi-deal table slots (128 B each)")
    ap.add_argument("--md-trav", type=int, default=340992, help="traversal pairs per multi-deal step")
    ap.add_argument("--curve-max", type=int, default=10 ** 7,
                    help="last point of the exploitability-vs-traversals curves (traversals per player; 0 = skip the curves)")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: only the two headline sections (and the CPU leg)")
    ap.add_argument("--only", default=None, help="N = 1: of the extra sections run only this one (profiling runs), e.g. sdcfr")
    ap.add_argument("--extras-budget-s", type=float, default=240.0,
                    help="N = 1: no further single-GPU reporting section is started after this many seconds of them")
    ap.add_argument("--collective", default="auto", choices=["auto", "p2p", "p2p_fused", "nccl"],
                    help="multi-GPU delta exchange of the headline number: the peer-memory exchange kernel after the traversal kernel "
                         "(p2p; auto = the same with an NCCL fallback when the peer mapping fails), the two fused into one launch "
                         "(p2p_fused), or NCCL all-reduce + apply (nccl); the other forms are timed beside it")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--ref-trav", type=int, default=1500)
    ap.add_argument("--ref-games", type=int, default=400_000)
    ap.add_argument("--ref-iters", type=int, default=8, help="--impl reference (python): MCCFR iterations per process per step")
    ap.add_argument("--ref-py-games", type=int, default=400, help="--impl reference (python): env games per process per step")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # stdout must carry exactly ONE line (the JSON): libraries print there too (NCCL writes its version
    # banner to stdout), so fd 1 points at stderr while the benchmark runs.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
