"""Throughput probe of the multi-deal MCCFR kernel (md_mccfr_kernel): python profiles/md_probe.py [D:log2cap ...]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from scopa_b200 import multideal

torch.cuda.set_device(0)
cfgs = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:] if ":" in a] or [(1, 12), (1024, 21), (16384, 25), (65536, 27)]
warm = 3
for a in sys.argv[1:]:
    if a.startswith("warm="):
        warm = int(a[5:])
for D, lc in cfgs:
    md = multideal.MultiDealSolver(np.arange(1, D + 1), log2_capacity=lc)
    n = 340992
    for b in range(warm):
        md.mccfr_batch(n, philox_seed=1, first_trav=b * n); md.apply()
    md.counters(reset=True)
    reps = 4
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps + 1)]
    ev[0].record()
    for b in range(reps):
        md.mccfr_batch(n, philox_seed=1, first_trav=(warm + b) * n)
        ev[2 * b + 1].record()
        md.apply()
        ev[2 * b + 2].record()
    torch.cuda.synchronize()
    t_trav = sum(ev[2 * b].elapsed_time(ev[2 * b + 1]) for b in range(reps)) / reps
    t_app = sum(ev[2 * b + 1].elapsed_time(ev[2 * b + 2]) for b in range(reps)) / reps
    c = md.counters()
    print(f"D={D} cap=2^{lc} ({md.table_bytes/2**30:.2f} GiB) infosets={c['infosets']} load={c['infosets']/md.capacity:.3f} "
          f"trav {t_trav:.3f} ms apply {t_app:.3f} ms -> {c['updates']/reps/t_trav/1e6:.2f} G upd/s (trav only), "
          f"{c['updates']/reps/(t_trav+t_app)/1e6:.2f} G incl apply; visits/upd {c['visits']/c['updates']:.2f}", flush=True)
    del md
    torch.cuda.empty_cache()
