"""time of one reference-semantics MCCFR iteration (in-place kernel, one device thread): python profiles/inplace_probe.py"""
import sys, time
sys.path.insert(0, '.')
import torch
from scopa_b200.solver import Solver
sv = Solver(seed=42)
sv.mccfr_inplace(50, philox_seed=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
sv.mccfr_inplace(2000, philox_seed=1, first_iter=50)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"in-place MCCFR: {dt / 2000 * 1e6:.1f} us per iteration ({2000 * 172 / dt / 1e6:.2f} M updates/s on one thread)")
