#!/usr/bin/env python
"""bench.py -- throughput of the Miniscopa hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mccfr|rollout]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

BASELINE.json's metric is double-barrelled ("MCCFR infoset-node updates/sec & env steps/sec"), so the one
JSON line carries both:
  * primary (metric/value/e2e/roofline/cpu_baseline): MCCFR infoset-node updates/s, config 3/5 of
    BASELINE.json -- a step = one batched MCCFR iteration on the seed-42 deal: `--trav` traversals per
    player per GPU against the frozen table (mccfr_tree_kernel: the estimator walking the deal's enumerated
    game tree; "mccfr_restep" beside it = the same estimator re-stepping the env at every node), one exchange
    of the slot-aligned delta buffer when N > 1, then table += delta (mccfr_apply_kernel).  Traversal ids are
    global (rank-offset), so the union of all ranks' work is independent of N ("weak" scaling: per-GPU
    work is fixed).
  * "env": env steps/s, config 2 of BASELINE.json -- 1 M concurrent random-policy games per GPU, a step =
    one fused rollout launch (8 plies per game) over deals already resident in HBM; its own e2e
    (seeds on the host -> actions + rewards on the host), roofline and CPU baseline.
`--workload rollout` swaps which of the two is reported as the primary metric.

--impl reference times the CPU restatement of the reference (oracle/, kind "port": the reference itself
is pure Python and /root/reference does not exist on the GPU box) on all host cores, same metric/config.
"""
import argparse
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
# DRAM bytes per launch of the dominant kernels, from the committed ncu --set full captures (profiles/README.md)
NCU_DRAM_BYTES_PER_LAUNCH = {"mccfr_batch_kernel": 117504, "rollout_kernel": 20052736,
                             "md_mccfr_kernel": 1111869952,   # 65 536 deals, 340 992 traversal pairs (md_r01f_raw.csv)
                             "mccfr_tree_kernel": 69376}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


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


# ------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from scopa_b200 import _lib
    from scopa_b200.batch import BatchedMiniScopa, rollout_random_host
    from scopa_b200.solver import Solver

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- scopa_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_gbs, peak_src = load_peaks()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    sync_word = torch.zeros(1, device=dev)

    def flush_l2():
        """Evict L2 between timed steps; with several ranks also line the ranks up again (stream-ordered
        all-reduce of one word), so that a rank's timed step does not include waiting for another rank's flush."""
        flush_buf.fill_(1)
        if world > 1:
            dist.all_reduce(sync_word)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    K, W = args.steps, args.warmup
    sampler = ClockSampler(local)

    # ------------------------------------------------------------------ MCCFR (configs 3 / 5)
    B = args.trav
    sv = Solver(seed=42, device=dev)
    S = sv.n_slots
    # exchange of the delta buffer between ranks: "p2p" = ms_mccfr_apply_peers (every rank reads every rank's
    # buffer over NVLink peer memory and applies the rank-ordered sum in one kernel), "nccl" = all-reduce + apply
    collective = "none"
    if world > 1:
        collective = "nccl"
        if args.collective in ("auto", "p2p"):
            ok = torch.ones(1, device=dev)
            try:
                sv.attach_peers()
            except Exception as e:          # e.g. no peer access between the GPUs of this box
                print(f"[rank {rank}] peer attach failed, using NCCL: {e}", file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() > 0:
                collective = "p2p"
            elif args.collective == "p2p":
                raise SystemExit("--collective p2p requested but peer memory is unavailable")
            else:
                sv = Solver(seed=42, device=dev)      # a clean, un-attached solver for the NCCL path
    delta = sv.delta_tensor() if collective != "p2p" else None

    def exchange_and_apply():
        if collective == "p2p":
            sv.apply_peers()
        else:
            if collective == "nccl":
                dist.all_reduce(delta)      # one all-reduce of 5*S float64 per iteration (NVLink / NVSwitch)
            sv.mccfr_apply()

    def mccfr_step(i):
        # global traversal ids: iteration i, rank r -> [ (i*world + r) * B, ... + B )
        sv.mccfr_batch(2, B, philox_seed=args.seed, first_trav=(i * world + rank) * B)
        exchange_and_apply()

    for i in range(W):
        mccfr_step(i)
    sv.counters(reset=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    launches0 = _lib.launch_count()
    sampler.start()
    wall0 = time.perf_counter()
    for i in range(K):
        flush_l2()
        ev[i][0].record()
        kev[i][0].record()
        sv.mccfr_batch(2, B, philox_seed=args.seed, first_trav=((W + i) * world + rank) * B)
        kev[i][1].record()
        exchange_and_apply()
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    mccfr_launches = _lib.launch_count() - launches0
    ms_total = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / K
    cnt = sv.counters()
    updates_all = sum_over_ranks(cnt["updates"])
    visits_all = sum_over_ranks(cnt["visits"])
    steps_all = sum_over_ranks(cnt["env_steps"])
    mccfr_value = updates_all / (ms_total * 1e-3)
    upd_per_launch = cnt["updates"] / K
    mccfr_roof_ach = upd_per_launch * BYTES_PER_UPDATE_FP64 / (ms_kernel * 1e-3) / 1e9

    # e2e: the whole solver state crosses PCIe every step (table in from pinned host memory, table out)
    h_reg = torch.zeros((S, 4), dtype=torch.float64).pin_memory()
    h_str = torch.zeros((S, 4), dtype=torch.float64).pin_memory()
    reg0, str0, _ = sv.export()
    h_reg.copy_(torch.from_numpy(reg0)); h_str.copy_(torch.from_numpy(str0))
    lib = _lib.load()
    e2e_steps = max(3, min(K, 10))
    barrier()
    sv.counters(reset=True)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        _lib.check(lib.ms_solver_import_table(sv.h, h_reg.data_ptr(), h_str.data_ptr(), sv._stream()))
        mccfr_step(W + K + i)
        _lib.check(lib.ms_solver_export_table(sv.h, None, None, None, h_reg.data_ptr(), h_str.data_ptr(), None,
                                              sv._stream()))
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_updates = sum_over_ranks(sv.counters()["updates"])
    mccfr_e2e = {"value": e2e_updates / e2e_s, "unit": "infoset-node updates/s",
                 "h2d_bytes_per_step": 2 * S * 4 * 8, "d2h_bytes_per_step": 2 * S * 4 * 8,
                 "what": "ms_solver_import_table (pinned host) + mccfr batch + all-reduce + apply + ms_solver_export_table"}

    # ------------------------------------------------------------------ the same estimator re-stepping the env (mode 3)
    # mccfr_batch_kernel: step(), capture resolution, legal list, infoset key and hash probe at every node instead of
    # walking the enumerated tree -- what the headline kernel was before; kept as the comparison point
    restep_obj = None
    if rank == 0:
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
        restep_obj = {"metric": "mccfr_infoset_node_updates_per_sec", "unit": "infoset-node updates/s",
                      "value": rc["updates"] / (rms * 1e-3), "ms_per_step": rms / K,
                      "env_steps_per_sec_inside_mccfr": rc["env_steps"] / (rms * 1e-3), "kernel": "mccfr_batch_kernel",
                      "config": {"workload": f"same estimator, mode 3 (env re-stepped at every node), {Br} traversals per player per step"}}

    # ------------------------------------------------------------------ textbook external sampling (opt-in estimator)
    es_obj = None
    if rank == 0:
        es_sv = Solver(seed=42, device=dev)
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
        es_obj = {"estimator": "external sampling (Lanctot et al. 2009), not in the reference", "kernel": "mccfr_es_tree_kernel",
                  "traversals_per_sec": 2.0 * Bes * K / es_s, "regret_updates_per_sec": es_cnt["updates"] / es_s,
                  "node_visits_per_sec": es_cnt["visits"] / es_s,
                  "exploitability_after": {"traversals_per_player": (W + K) * Bes, "value": es_sv.exploitability(1)},
                  "note": "the reference's own estimator plateaus near 0.49 exploitability on this deal"}
        del es_sv
        # BASELINE.json configs[2]: "10M traversals, exploitability vs iteration" -- both estimators from an empty table,
        # 4096 traversals per player per iteration (frozen-sigma batches), exploitability by the device best response
        curve = {}
        for mode_id, name, kind in ((1, "external_sampling", 1), (0, "reference_estimator", 1)):
            csv_ = Solver(seed=42, device=dev)
            done, pts, Bc = 0, [], 4096
            t0 = time.perf_counter()
            for target in [t for t in (10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7) if t <= args.curve_max]:
                while done < target:
                    csv_.mccfr_batch(2, Bc, philox_seed=args.seed, first_trav=done, mode=mode_id)
                    csv_.mccfr_apply()
                    done += Bc
                pts.append([done, csv_.exploitability(kind)])
            curve[name] = {"traversals_per_player_vs_exploitability": pts, "wall_s": time.perf_counter() - t0}
            del csv_
        es_obj["exploitability_vs_traversals"] = curve

    # ------------------------------------------------------------------ env rollouts (config 2)
    G = args.games
    seeds_np = np.arange(1 + rank * G, 1 + (rank + 1) * G, dtype=np.int64)
    b = BatchedMiniScopa(dev).reset(seeds_np)
    actions = torch.empty((G, 8), dtype=torch.uint8, device=dev)
    rewards = torch.empty((G, 2), dtype=torch.float32, device=dev)
    for i in range(W):
        b.rollout_random(philox_seed=args.seed, game_offset=rank * G, actions=actions, rewards=rewards)
    rev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    launches1 = _lib.launch_count()
    for i in range(K):
        flush_l2()
        rev[i][0].record()
        b.rollout_random(philox_seed=args.seed + i, game_offset=rank * G, actions=actions, rewards=rewards)
        rev[i][1].record()
    barrier()
    env_launches = _lib.launch_count() - launches1
    clocks = sampler.stop()
    env_ms_total = max_over_ranks(sum(a.elapsed_time(c) for a, c in rev))
    env_value = 8.0 * G * world * K / (env_ms_total * 1e-3)
    env_kernel_ms = sum(a.elapsed_time(c) for a, c in rev) / K
    env_roof_ach = 8.0 * G * BYTES_PER_ENV_STEP / (env_kernel_ms * 1e-3) / 1e9
    # e2e: seeds in pinned host memory -> deal + rollout on device -> actions + rewards in pinned host memory
    h_seeds = torch.from_numpy(seeds_np).pin_memory()
    h_act = torch.empty((G, 8), dtype=torch.uint8).pin_memory()
    h_rew = torch.empty((G, 2), dtype=torch.float32).pin_memory()
    rollout_random_host(h_seeds, args.seed, rank * G, h_act, h_rew)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        rollout_random_host(h_seeds, args.seed + i, rank * G, h_act, h_rew)
    barrier()
    env_e2e_s = max_over_ranks(time.perf_counter() - t0)
    env_e2e = {"value": 8.0 * G * world * e2e_steps / env_e2e_s, "unit": "env steps/s",
               "h2d_bytes_per_step": 8 * G, "d2h_bytes_per_step": 16 * G,
               "what": "ms_rollout_random_host: reset(seed) + 8 steps per game, host buffers in and out"}

    # ------------------------------------------------------------------ step-granular env API (HBM-bound kernels)
    # ms_step round-trips the 16-byte state through HBM: the one kernel family here whose real bound IS the HBM
    # roofline (33 B/step + 8 B rewards + 1 B done = 42 B moved per step with all outputs requested)
    NS = args.step_states
    bs = BatchedMiniScopa(dev).reset(np.arange(1, NS + 1, dtype=np.int64))
    _, ordered0, _ = bs.legal_actions()
    acts_u8 = ordered0[:, 0].contiguous()                  # first legal card of every game
    st_backup = bs.states.clone()
    rew_s = torch.empty((NS, 2), dtype=torch.float32, device=dev)
    done_t = torch.empty((NS,), dtype=torch.uint8, device=dev)
    for i in range(W):
        bs.step(acts_u8, rewards=rew_s, done=done_t)
    stev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for i in range(K):
        bs.states.copy_(st_backup)
        flush_l2()
        stev[i][0].record()
        bs.step(acts_u8, rewards=rew_s, done=done_t)
        stev[i][1].record()
    barrier()
    step_ms = sum(a.elapsed_time(c) for a, c in stev) / K
    step_obj = {"kernel": "step_kernel", "env_steps_per_sec": NS / (step_ms * 1e-3), "kernel_ms": step_ms,
                "bytes_per_step": 42, "achieved_gbs": 42.0 * NS / (step_ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": 42.0 * NS / (step_ms * 1e-3) / 1e9 / hbm_gbs,
                "note": f"{NS} states ({16 * NS >> 20} MiB, larger than L2), one ply per launch, state + action in, "
                        "state + rewards + done out"}
    del bs, st_backup, rew_s, done_t

    # ------------------------------------------------------------------ vanilla CFR (config 1)
    cfr_sv = Solver(seed=42, device=dev)
    cfr_sv.cfr_iterate(5)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    cfr_sv.cfr_iterate(200)
    c1.record()
    torch.cuda.synchronize()
    cfr_obj = {"kernel": "cfr_kernel", "us_per_iteration": c0.elapsed_time(c1) * 1e3 / 200,
               "node_visits_per_sec": 2 * 2229 * 200 / (c0.elapsed_time(c1) * 1e-3),
               "note": "BASELINE.json configs[0]: vanilla CFR on the seed-42 deal, 200 iterations in one launch, float64, "
                       "bit-identical to the reference's tables; the reference takes 389 ms per iteration on one CPU core"}

    # throughput mode: 148 independent deals, one CTA (= one SM) per deal, one launch
    from scopa_b200.solver import cfr_iterate_many
    many = [Solver(seed=1000 + rank * 148 + i, device=dev) for i in range(148)]
    cfr_iterate_many(many, 3)
    torch.cuda.synchronize()
    c0.record()
    cfr_iterate_many(many, 100)
    c1.record()
    torch.cuda.synchronize()
    nodes = sum(m_.n_nodes for m_ in many)
    cfr_obj["many_deals"] = {"deals": 148, "iterations": 100, "ms": c0.elapsed_time(c1),
                             "deal_iterations_per_sec": 148 * 100 / (c0.elapsed_time(c1) * 1e-3),
                             "node_visits_per_sec": 2.0 * nodes * 100 / (c0.elapsed_time(c1) * 1e-3),
                             "note": "ms_cfr_iterate_many: seeds 1000.., float64, each deal's tables identical to a solo run"}
    del many

    # ------------------------------------------------------------------ atomic roofline
    atom_obj = None
    if rank == 0:
        import ctypes
        peaks = (ctypes.c_double * 3)()
        _lib.check(_lib.load().ms_debug_atomic_peaks(peaks, _lib.stream_ptr()))
        pairs_per_s = mccfr_value / world / 172.0          # traversal pairs per second on this GPU
        # per traversal pair the batch kernel issues 118 shared-memory fp64 atomic adds (regret deltas of the
        # traverser nodes with more than one action: (1*4 + 5*3 + 20*2) per player) and 172 u32 adds (visit counts)
        atom_obj = {"measured_peaks_per_sec": {"smem_f64_atomic_add": peaks[0], "smem_u32_atomic_add": peaks[1],
                                               "global_red_f64_l2_resident": peaks[2]},
                    "mccfr_smem_f64_atomics_per_sec": 118.0 * pairs_per_s, "mccfr_smem_u32_atomics_per_sec": 172.0 * pairs_per_s,
                    "frac_of_smem_f64_peak": 118.0 * pairs_per_s / peaks[0], "frac_of_smem_u32_peak": 172.0 * pairs_per_s / peaks[1],
                    "note": "microbenchmark: 148 CTAs x 768 threads, pseudo-random addresses over a 738x4 table; shared-memory "
                            "fp64 atomicAdd compiles to an ATOMS.CAST.SPIN.64 compare-and-swap loop, global fp64 to REDG.E.ADD.F64"}

    # ------------------------------------------------------------------ multi-deal MCCFR
    # The regime SURVEY 8(d) names as the one where the memory system is the bound: one infoset table for 65 536
    # deals (5.1 M stored infosets, 0.66 GB of 128-byte lines, far beyond the 126 MB L2) in HBM.
    md_obj = mdb_obj = None
    if rank == 0 and args.md_deals > 0:
        import ctypes
        from scopa_b200 import multideal
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
        VIS, PPV = 148, 3072
        tb0 = time.perf_counter()
        md.mccfr_blocked(VIS, pairs_per_visit=PPV, philox_seed=args.seed, first_visit=0)     # first call builds the deal records
        md.apply()
        torch.cuda.synchronize()
        build_s = time.perf_counter() - tb0
        for i in range(1, 4):
            md.mccfr_blocked(VIS, pairs_per_visit=PPV, philox_seed=args.seed, first_visit=i * VIS)
            md.apply()
        md.counters(reset=True)
        bev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        bev[0].record()
        for i in range(K):
            md.mccfr_blocked(VIS, pairs_per_visit=PPV, philox_seed=args.seed, first_visit=(4 + i) * VIS)
            md.apply()
            bev[i + 1].record()
        torch.cuda.synchronize()
        bc = md.counters()
        b_ms = bev[0].elapsed_time(bev[K]) / K
        mdb_obj = {"metric": "mccfr_infoset_node_updates_per_sec", "unit": "infoset-node updates/s",
                   "value": bc["updates"] / K / (b_ms * 1e-3), "ms_per_step": b_ms, "infosets": int(bc["infosets"]),
                   "first_call_s_incl_describing_all_deals": build_s, "kernel": "md_blocked_kernel",
                   "config": {"workload": f"same table and estimator, deal-blocked: {VIS} visits x {PPV} traversal pairs per step, one "
                                          "deal per CTA visit staged in shared memory (tree, strategies, delta tables), table "
                                          "read once and written once per visit"}}
        del md
        torch.cuda.empty_cache()
        _lib.check(_lib.load().ms_debug_random_access_peaks(lines_lg, rp, _lib.stream_ptr()))
        # per traversal pair (data-independent recursion shape of the estimator, 4+4-card deals): 163 lookups of
        # stored infosets (player 0 traversal: 26 own + 85 opponent nodes with >1 card; player 1: 26 + 26) and 52
        # update groups (regret-delta REDs + visit count on the line just read)
        touches = (163.0 + 52.0) * nb / (t_trav * 1e-3)
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
                               "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get("md_mccfr_kernel"), "kernel": "md_mccfr_kernel",
                               "note": "line touches = 163 lookups + 52 update groups per traversal pair; hot (shallow) "
                                       "infosets hit in L2 (ncu: 75 % of sectors), so the DRAM-resident ceiling is not the "
                                       "binding one yet: the kernel is latency-bound (profiles/README.md section 6)"},
                  "config": {"workload": f"MCCFR (reference estimator) over {args.md_deals} deals (seeds 1..), chance-sampled root, "
                                         f"{nb} traversal pairs per step, one fp64 infoset table of 2^{lg} x 128 B in HBM",
                             "note": "updates are reference-equivalent (172 per traversal pair); infosets with one card in hand "
                                     "(120 of the 172) are not stored -- their strategy is the constant [1.0]"}}

    # ------------------------------------------------------------------ 40-card Scopa rollouts
    full_obj = None
    if rank == 0 and args.full_games > 0:
        from scopa_b200 import full as fs
        FG = args.full_games
        fb = fs.BatchedFullScopa(dev).reset(np.arange(1, FG + 1, dtype=np.int64))
        for i in range(2):
            fb.rollout_random(philox_seed=args.seed, game_offset=i * FG)
        fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush_l2()
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
        if world == 1 and not args.no_cpu:
            from oracle import ms_oracle as ora              # CPU leg: the checker timed as the baseline
            ncpu = os.cpu_count() or 1
            ora.full_rollout_random(h_seeds[:2000], args.seed)
            t0 = time.perf_counter()
            ora.full_rollout_random(h_seeds[:200_000], args.seed)
            dt = time.perf_counter() - t0
            cpu_full = {"value": 200_000 * fs.PLIES / dt, "unit": "env steps/s", "cores": ncpu, "kind": "port",
                        "sample": "200 000 games x 36 plies, OpenMP over all host threads"}
        full_obj = {"metric": "env_steps_per_sec", "unit": "env steps/s", "value": FG * fs.PLIES / (f_ms * 1e-3), "ms_per_step": f_ms,
                    "e2e": {"value": FG * fs.PLIES / f_e2e, "unit": "env steps/s", "h2d_bytes_per_step": 8 * FG,
                            "d2h_bytes_per_step": (fs.PLIES + 8) * FG,
                            "what": "ms_full_rollout_random_host: FullDeck(seed) + deal + 36 plies per game, pinned host buffers in and out, "
                                    "H2D / kernels / D2H pipelined over three streams"},
                    "roofline": {"bound": "hbm", "achieved": FG * fs.PLIES * 65.0 / (f_ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                                 "frac": FG * fs.PLIES * 65.0 / (f_ms * 1e-3) / 1e9 / hbm_gbs, "traffic": None,
                                 "kernel": "full_rollout_kernel", "peak_source": peak_src,
                                 "note": "against the step-granular figure for this game, 65 B/step (32 B state load + 1 B action "
                                         "+ 32 B state store); the fused kernel keeps the state in registers and moves "
                                         "108 B/game = 3 B/step, so it is integer-issue bound like the Miniscopa rollout"},
                    "cpu_baseline": cpu_full,
                    "config": {"workload": f"{FG} concurrent random-policy games of 40-card Scopa (FullScopaEnv), 36 plies each, "
                                           "deals resident in HBM", "l2": "256 MiB flush between timed steps"}}
        del fb

    # ------------------------------------------------------------------ SDCFR traversal (config 4)
    from scopa_b200 import sdcfr as sd
    T = args.sd_trav
    torch.manual_seed(1234)
    blobs = [(torch.randn(sd.NET_FLOATS, device=dev) * 0.15).contiguous() for _ in range(2)]   # random-init nets
    trv = sd.Traverser(sv.root_words, sv.hand_order, device=dev)
    sd_obj = None
    for prec, pname in ((sd.TENSOR_CORE, "bf16_tcgen05"), (sd.FP32, "fp32_cuda_cores")):
        for i in range(2):
            trv.run(i & 1, blobs, T, philox_seed=args.seed, first_trav=rank * T, precision=prec)
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        l0 = _lib.launch_count()
        for i in range(K):
            flush_l2()
            sev[i][0].record()
            trv.run(0, blobs, T, philox_seed=args.seed, first_trav=(i * world + rank) * T, precision=prec)
            trv.run(1, blobs, T, philox_seed=args.seed, first_trav=(i * world + rank) * T, precision=prec)
            sev[i][1].record()
        barrier()
        sd_launches = _lib.launch_count() - l0
        sd_ms = max_over_ranks(sum(a.elapsed_time(c) for a, c in sev))
        inf = 187.0 * T * world * K                      # 105 + 82 advantage-net inferences per traversal pair
        o = {"inferences_per_sec": inf / (sd_ms * 1e-3), "traversals_per_sec": 2.0 * T * world * K / (sd_ms * 1e-3),
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
                          "kernel": "sd_forward_kernel<1>", "peak_source": tf_src,
                          "note": "numerator = 27 136 FLOP x inferences of WHOLE traversals (env steps, sampling, backward levels "
                                  "and sample emission included in the time), so this is a floor on the tensor-pipe share; "
                                  "K = 34 / N = 16 layers pad to MMA tiles (profiles/README.md section 3)"}
    if rank == 0:
        try:
            sd_obj["train"] = bench_sd_train(dev, _lib)
        except Exception as e:      # an auxiliary section must not take the headline line down with it
            sd_obj["train"] = {"error": repr(e)}
        try:
            sd_obj["drop_in_train"] = bench_sd_dropin(dev)
        except Exception as e:
            sd_obj["drop_in_train"] = {"error": repr(e)}

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu_mccfr = cpu_env = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_mccfr, cpu_env = cpu_baselines(args, sample_seconds=8.0)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None

    mccfr_obj = {
        "metric": "mccfr_infoset_node_updates_per_sec", "value": mccfr_value, "unit": "infoset-node updates/s",
        "ms_per_step": ms_total / K, "e2e": mccfr_e2e, "gpu_launches": int(mccfr_launches),
        "node_visits_per_sec": visits_all / (ms_total * 1e-3), "tree_edges_per_sec_inside_mccfr": steps_all / (ms_total * 1e-3),
        "roofline": {"bound": "hbm", "achieved": mccfr_roof_ach, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": mccfr_roof_ach / hbm_gbs, "traffic": NCU_DRAM_BYTES_PER_LAUNCH["mccfr_tree_kernel"],
                     "traffic_source": "ncu --set full, profiles/mccfr_r01h_raw.csv: dram__bytes_read.sum + "
                                       "dram__bytes_write.sum per launch (table and tree staging only; independent of the batch size)",
                     "kernel": "mccfr_tree_kernel",
                     "kernel_ms": ms_kernel, "peak_source": peak_src,
                     "on_chip": {"issue_slots_active_pct": 76.9, "shared_memory_wavefronts_pct_of_peak": 65.4,
                                 "alu_pipe_pct": 53.6, "source": "ncu --set full, profiles/mccfr_r01h_raw.csv"},
                     "note": "HBM-EQUIVALENT figure SURVEY 8(d) prescribes (203.7 algorithmic B/update: what an HBM-resident table "
                             "would have to move); it exceeds 1 because the 53 KB table and the 9 KB tree of the one deal are "
                             "shared-memory resident -- this is NOT an HBM result.  The binding resources are on chip: "
                             "issue slots 77 % busy, shared-memory wavefronts at 65 % of peak (profiles/README.md section 1)"},
        "cpu_baseline": cpu_mccfr,
        "config": {"workload": "BASELINE.json configs[2]/[4]: MCCFR (reference estimator), seed-42 deal, "
                               f"{B} traversals per player per GPU per iteration, fp64 table", "traversals_per_step": 2 * B * world,
                   "parallelism": f"dp{world}: traversals sharded by id, one exchange of {5 * S} f64 per iteration via " +
                                  ("ms_mccfr_apply_peers (NVLink peer-memory reads + rank-ordered sum + apply in one kernel)"
                                   if collective == "p2p" else "NCCL all-reduce")
                   if world > 1 else "single GPU",
                   "l2": "256 MiB flush between timed steps (working set is on-chip anyway)", "philox_seed": args.seed},
    }
    env_obj = {
        "metric": "env_steps_per_sec", "value": env_value, "unit": "env steps/s", "ms_per_step": env_ms_total / K,
        "e2e": env_e2e, "gpu_launches": int(env_launches),
        "roofline": {"bound": "hbm", "achieved": env_roof_ach, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": env_roof_ach / hbm_gbs, "traffic": NCU_DRAM_BYTES_PER_LAUNCH["rollout_kernel"] * (G / 1e6),
                     "traffic_source": "ncu --set full, profiles/env_r01c_raw.csv: 20.05 MB read per 1 M games (writes "
                                       "stayed in L2 under ncu)",
                     "kernel": "rollout_kernel", "kernel_ms": env_kernel_ms,
                     "peak_source": peak_src,
                     "note": "against the step-granular 34 B/step figure (SURVEY 8(d)); the fused kernel itself moves "
                             "36 B/game (4.5 B/step) and is integer-issue bound, not HBM bound"},
        "cpu_baseline": cpu_env,
        "config": {"workload": f"BASELINE.json configs[1]: {G} concurrent random-policy games per GPU, 8 plies each, "
                               "deals resident in HBM", "l2": "256 MiB flush between timed steps", "games_per_gpu": G},
    }
    primary, secondary = (mccfr_obj, env_obj) if args.workload == "mccfr" else (env_obj, mccfr_obj)
    line = {
        "metric": primary["metric"], "value": primary["value"], "unit": primary["unit"], "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": primary["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if primary is mccfr_obj else "u32", "data": "synthetic",
        "config": primary["config"], "e2e": primary["e2e"], "gpu_launches": primary["gpu_launches"],
        "roofline": primary["roofline"], "cpu_baseline": primary["cpu_baseline"], "clocks": clocks,
        "wall_s_mccfr_region": wall,
        ("env" if primary is mccfr_obj else "mccfr"): secondary,
        "sdcfr": sd_obj,
        "env_step_api": step_obj,
        "cfr": cfr_obj,
        "mccfr_external_sampling": es_obj,
        "atomics": atom_obj,
        "mccfr_restep": restep_obj,
        "mccfr_multi_deal": md_obj,
        "mccfr_multi_deal_blocked": mdb_obj,
        "full_scopa": full_obj,
        "collective": collective,
    }
    if primary is mccfr_obj:
        line["node_visits_per_sec"] = mccfr_obj["node_visits_per_sec"]
    return line


# ------------------------------------------------------------------------------------------- CPU side
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
    # the cluster form has not run on a GPU yet (written after the round's GPU budget was spent): opt-in here
    variants = ("fused", "torch") + (("fused-cluster",) if os.environ.get("SCOPA_B200_BENCH_CLUSTER") == "1" else ())
    for name in variants:
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
        out[name] = {"us_per_step": 1e6 * dt / (calls * epochs), "ms_per_call": 1e3 * dt / calls, "last_loss": loss,
                     "our_launches_per_call": (_lib.launch_count() - l0) / calls}
        if name != "torch":     # the kernel alone (CUDA events around the launches, minibatch rows drawn beforehand)
            idx = adv._sample_rows(128, epochs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(calls):
                adv._fused.step(adv.buffer.feat, adv.buffer.target, adv.buffer.mask, idx)
            e1.record()
            torch.cuda.synchronize()
            out[name]["kernel_us_per_step"] = 1e3 * e0.elapsed_time(e1) / (calls * epochs)
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
    for name, kw in (("default", {}), ("fused_optimizer_device_eval", {"optimizer": "fused", "device_eval": True})):
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
    return out


def cpu_baselines(args, sample_seconds, threads=None):
    """Times the CPU restatement (oracle/, a C port of the reference's algorithm) on the host cores."""
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
    return cpu_mccfr, cpu_env


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return None
    K, W = args.steps, args.warmup
    from oracle import ms_oracle as ora
    ora.build()
    cores = os.cpu_count() or 1
    # one "step" = a bounded sample of the same workload: traversal pairs on every core / games on every core
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
    mccfr = {"metric": "mccfr_infoset_node_updates_per_sec", "value": tot_u / dt, "unit": "infoset-node updates/s",
             "ms_per_step": dt / K * 1e3, "node_visits_per_sec": tot_v / dt}
    env = {"metric": "env_steps_per_sec", "value": 8.0 * games * K / dt_env, "unit": "env steps/s",
           "ms_per_step": dt_env / K * 1e3}
    primary, secondary = (mccfr, env) if args.workload == "mccfr" else (env, mccfr)
    sample = (f"{per_thread} traversal pairs x {cores} workers per step" if primary is mccfr
              else f"{games} games per step")
    line = {
        "impl": "reference", "metric": primary["metric"], "value": primary["value"], "unit": primary["unit"],
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": primary["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if primary is mccfr else "u32", "data": "synthetic",
        "config": {"workload": "same metric/config as the CUDA arm, CPU restatement of the reference (oracle/ms_oracle.c) "
                               "on all host cores; each step is a bounded sample: " + sample},
        "cpu_baseline": {"value": primary["value"], "unit": primary["unit"], "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": primary["value"], "unit": primary["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        ("env" if primary is mccfr else "mccfr"): secondary,
        "note": "the reference itself is single-threaded pure Python (survey-measured 2.8 k updates/s, 106 k env steps/s "
                "on one core); this C port is about 80x faster per core and uses every core",
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mccfr", choices=["mccfr", "rollout"])
    ap.add_argument("--trav", type=int, default=454656,
                    help="traversals per player per GPU per step (default: 3 full waves of 148 CTAs x 1024 threads)")
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
    ap.add_argument("--collective", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU delta exchange: peer-memory kernel (p2p), NCCL all-reduce, or p2p with NCCL fallback")
    ap.add_argument("--ref-trav", type=int, default=1500)
    ap.add_argument("--ref-games", type=int, default=400_000)
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
