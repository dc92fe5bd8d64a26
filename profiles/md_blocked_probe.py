"""Throughput of the deal-blocked multi-deal MCCFR kernel: python profiles/md_blocked_probe.py [deals:log2cap:pairs_per_visit ...]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from scopa_b200 import multideal

torch.cuda.set_device(0)
cfgs = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(1024, 21, 3072), (65536, 26, 3072), (65536, 26, 1024), (65536, 26, 9216)]
for D, lc, P in cfgs:
    md = multideal.MultiDealSolver(np.arange(1, D + 1), log2_capacity=lc)
    t0 = time.perf_counter()
    md.mccfr_blocked(148, pairs_per_visit=P, philox_seed=1, first_visit=0)
    md.apply()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    for b in range(1, 4):
        md.mccfr_blocked(148, pairs_per_visit=P, philox_seed=1, first_visit=148 * b); md.apply()
    md.counters(reset=True)
    reps = 5
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps + 1)]
    ev[0].record()
    for b in range(reps):
        md.mccfr_blocked(148, pairs_per_visit=P, philox_seed=1, first_visit=148 * (4 + b))
        ev[2 * b + 1].record()
        md.apply()
        ev[2 * b + 2].record()
    torch.cuda.synchronize()
    t_trav = sum(ev[2 * b].elapsed_time(ev[2 * b + 1]) for b in range(reps)) / reps
    t_app = sum(ev[2 * b + 1].elapsed_time(ev[2 * b + 2]) for b in range(reps)) / reps
    c = md.counters()
    print(f"D={D} cap=2^{lc} pairs/visit={P}: first call (build + visit) {t_build*1e3:.1f} ms; infosets={c['infosets']} load={c['infosets']/md.capacity:.3f}; "
          f"148 visits {t_trav:.3f} ms + apply {t_app:.3f} ms -> {c['updates']/reps/(t_trav+t_app)/1e6:.2f} G upd/s", flush=True)
    del md
    torch.cuda.empty_cache()
