"""Multi-GPU check of the peer-memory exchange (run under torchrun, one process per GPU; tests/test_gpu_multigpu.py
launches it with 2 ranks when the box has 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_peers_check.py

For the same traversal ids, {ms_mccfr_batch, ms_mccfr_apply_peers} must produce the table that
{ms_mccfr_batch, NCCL all-reduce, ms_mccfr_apply} produces (fp64 summation order differs: 1e-9), every rank's
replica must hold the SAME bits on the peer path (rank-ordered sums), and it times both exchanges at a small and
a large batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scopa_b200.solver import Solver  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    a, b, f = Solver(seed=42, device=dev), Solver(seed=42, device=dev), Solver(seed=42, device=dev)
    a.attach_peers()
    f.attach_peers()
    B = 4096
    for it in range(6):
        first = (it * world + rank) * B
        a.mccfr_batch(2, B, philox_seed=5, first_trav=first)
        a.apply_peers()
        f.mccfr_batch_peers(2, B, philox_seed=5, first_trav=first)      # the fused form: one launch
        b.mccfr_batch(2, B, philox_seed=5, first_trav=first)
        dist.all_reduce(b.delta_tensor())
        b.mccfr_apply()
    torch.cuda.synchronize()
    assert a.peer_error() == 0 and f.peer_error() == 0
    rf, sf, tf = f.export()
    ra, sa, ta = a.export()
    np.testing.assert_allclose(rf, ra, rtol=1e-9, atol=1e-9)      # (CTAs flush their deltas in a different order)
    np.testing.assert_allclose(sf, sa, rtol=1e-9, atol=1e-9)
    assert np.array_equal(tf, ta)
    tt = torch.from_numpy(np.concatenate([rf.ravel(), sf.ravel()])).to(dev)          # fused path: replicas bit-identical too
    lo_, hi_ = tt.clone(), tt.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    assert torch.equal(lo_, hi_), "replicas diverged on the fused path"
    rb, sb, tb = b.export()
    np.testing.assert_allclose(ra, rb, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(sa, sb, rtol=1e-9, atol=1e-9)
    assert np.array_equal(ta, tb) and ta.sum() > 700        # which InfoNodes exist: same on both paths
    # ... and equal to ONE GPU running every rank's traversal ids (the result does not depend on the number of ranks)
    if rank == 0:
        one = Solver(seed=42, device=dev)
        for it in range(6):
            for r in range(world):
                one.mccfr_batch(2, B, philox_seed=5, first_trav=(it * world + r) * B)
            one.mccfr_apply()
        r1, s1, t1 = one.export()
        np.testing.assert_allclose(ra, r1, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(sa, s1, rtol=1e-9, atol=1e-9)
        assert np.array_equal(ta, t1)
    # replicas are bit-identical on the peer path (tables and touched flags)
    t = torch.from_numpy(np.concatenate([ra.ravel(), sa.ravel(), ta.astype(np.float64)])).to(dev)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "replicas diverged"

    def timed(fn, batch, iters=30):
        for i in range(5):
            fn(batch, 1000 + i)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(batch, 2000 + i)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_peers(batch, i):
        a.mccfr_batch(2, batch, philox_seed=6, first_trav=(i * world + rank) * batch)
        a.apply_peers()

    def step_fused(batch, i):
        f.mccfr_batch_peers(2, batch, philox_seed=6, first_trav=(i * world + rank) * batch)

    def step_nccl(batch, i):
        b.mccfr_batch(2, batch, philox_seed=6, first_trav=(i * world + rank) * batch)
        dist.all_reduce(b.delta_tensor())
        b.mccfr_apply()

    def step_none(batch, i):       # no exchange at all: the floor
        b.mccfr_batch(2, batch, philox_seed=6, first_trav=(i * world + rank) * batch)
        b.mccfr_apply()

    out = {}
    for batch in ((768, 113664) if os.environ.get("PEERS_CHECK_TIMING", "1") == "1" else ()):
        out[batch] = {"peers_ms": timed(step_peers, batch), "fused_ms": timed(step_fused, batch), "nccl_ms": timed(step_nccl, batch), "no_exchange_ms": timed(step_none, batch)}
    if rank == 0:
        print("PEERS_CHECK_OK world=%d" % world, out, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
