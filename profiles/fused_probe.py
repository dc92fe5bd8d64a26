"""Single-GPU probe of the fused traversal + exchange kernel: with world = 1 the "exchange" is a push into the own inbox, a
barrier with oneself and the apply -- so {ms_mccfr_batch, ms_mccfr_apply} (2 launches), {ms_mccfr_batch,
ms_mccfr_apply_peers} (2 launches) and ms_mccfr_batch_peers (1 launch) do the same work and can be timed side by side
without a second GPU.   python profiles/fused_probe.py [traversals per player]"""
import ctypes as C
import sys

sys.path.insert(0, ".")
import torch

from scopa_b200 import _lib
from scopa_b200.solver import Solver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1818624
lib = _lib.load()


def attach_self(sv):
    handle, offs = (C.c_ubyte * 64)(), (C.c_uint64 * 3)()
    _lib.check(lib.ms_solver_ipc_export(sv.h, handle, offs))
    flat = (C.c_uint64 * 3)(*offs)
    _lib.check(lib.ms_solver_ipc_attach(sv.h, 0, 1, bytes(handle), flat))


def timed(fn, iters=30):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(100 + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


a, b, c = Solver(seed=42), Solver(seed=42), Solver(seed=42)
attach_self(b)
attach_self(c)


def two(i):
    a.mccfr_batch(2, B, philox_seed=1, first_trav=i * B)
    a.mccfr_apply()


def peers(i):
    b.mccfr_batch(2, B, philox_seed=1, first_trav=i * B)
    b.apply_peers()


def fused(i):
    c.mccfr_batch_peers(2, B, philox_seed=1, first_trav=i * B)


def batch_only(i):
    a.mccfr_batch(2, B, philox_seed=1, first_trav=i * B)


for name, fn in (("batch only", batch_only), ("batch + apply", two), ("batch + apply_peers", peers), ("fused batch_peers", fused),
                 ("batch + apply", two), ("fused batch_peers", fused)):
    print(f"{name:24s} {timed(fn) * 1e3:9.1f} us per iteration (B = {B})", flush=True)
print("peer errors", b.peer_error(), c.peer_error())
ra, sa, _ = a.export(); rc, sc, _ = c.export()
import numpy as np
print("tables finite", np.isfinite(ra).all() and np.isfinite(rc).all())
