"""Multi-GPU check of the SHARDED multi-deal infoset table (run under torchrun, one process per GPU;
tests/test_gpu_multigpu.py launches it with 2 ranks when the box has 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_md_check.py

Every rank owns the shard of the table its hash assigns to it ("shard by hash(key) % G") and runs its
share of every iteration's visits with md_blocked_kernel, which gathers regrets from and sends deltas to the owners'
shards through peer memory.  The union of the shards must equal the table ONE GPU builds when it runs all the visits
(same infoset set, 1e-9: the order of the fp64 additions differs), also after the per-traversal kernel has worked on
the same sharded table, and no barrier may have timed out.  With MD_CHECK_TIMING=1 it also times an iteration at a
large batch against the same per-GPU work on an unsharded table."""
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scopa_b200.multideal import MultiDealSolver  # noqa: E402
from scopa_b200.sharding import shard_bounds  # noqa: E402


def close(reg, oreg, strat, ostrat):
    wild = (np.abs(reg) > 1e9).any(1) | (np.abs(oreg) > 1e9).any(1)      # ill-conditioned importance weights
    assert wild.mean() < 0.01
    np.testing.assert_allclose(reg[~wild], oreg[~wild], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(strat[~wild], ostrat[~wild], rtol=1e-9, atol=1e-9)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    seeds = np.arange(1, 201, dtype=np.int64)
    sh = MultiDealSolver(seeds, log2_capacity=18, device=dev)
    sh.attach_peers()
    V, P = 24, 512
    for it in range(4):
        sh.iterate_blocked(V, P, philox_seed=7, first_visit=V * it)
    lo, n = shard_bounds(3000, rank, world)                              # the per-traversal kernel on the sharded table
    sh.mccfr_batch(n, philox_seed=3, first_trav=10 ** 6 + lo)
    sh.barrier(); sh.apply(); sh.barrier()
    torch.cuda.synchronize()
    assert sh.peer_error() == 0
    cnt = sh.counters()
    tot = torch.tensor([cnt["updates"], cnt["visits"], cnt["infosets"]], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    mine = len(sh.export_shard()[0])
    keys, reg, strat = sh.export()                                       # all shards, gathered
    assert len(np.unique(keys)) == len(keys), "an infoset lives in two shards"
    assert int(tot[2]) == len(keys)
    assert mine > 0.7 * len(keys) / world, (mine, len(keys))             # the owner hash spreads the infosets
    if rank == 0:
        one = MultiDealSolver(seeds, log2_capacity=19, device=dev)
        for it in range(4):
            one.mccfr_blocked(V, P, philox_seed=7, first_visit=V * it)
            one.apply()
        one.mccfr_batch(3000, philox_seed=3, first_trav=10 ** 6)
        one.apply()
        c1 = one.counters()
        k1, r1, s1 = one.export()
        assert np.array_equal(keys, k1), "infoset sets differ"
        assert (int(tot[0]), int(tot[1]), int(tot[2])) == (c1["updates"], c1["visits"], c1["infosets"])
        close(reg, r1, strat, s1)
        assert np.abs(reg).sum() > 0
        lreg, lstrat, found = sh.lookup(keys[::101])                    # any rank reads any shard
        assert bool(found.all()) and np.array_equal(lreg.cpu().numpy(), reg[::101]) and np.array_equal(lstrat.cpu().numpy(), strat[::101])
    dist.barrier()
    if os.environ.get("MD_CHECK_TIMING", "0") == "1":
        D, Vr, Pr = 16384, 148 * 4, 3072                                 # per rank: 4 visits per SM
        big = np.arange(1, D + 1, dtype=np.int64)
        a = MultiDealSolver(big, log2_capacity=23, device=dev)
        a.attach_peers()
        b = MultiDealSolver(big, log2_capacity=24, device=dev)           # the same per-GPU work on a private table
        out = {}
        for name, sv in (("sharded", a), ("private", b)):
            for w in range(3):
                if name == "sharded":
                    sv.iterate_blocked(Vr * world, Pr, philox_seed=1, first_visit=Vr * world * w)
                else:
                    sv.mccfr_blocked(Vr, Pr, philox_seed=1, first_visit=Vr * w); sv.apply()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K = 10
            for w in range(K):
                if name == "sharded":
                    sv.iterate_blocked(Vr * world, Pr, philox_seed=1, first_visit=Vr * world * (3 + w))
                else:
                    sv.mccfr_blocked(Vr, Pr, philox_seed=1, first_visit=Vr * (3 + w)); sv.apply()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name] = float(t)
        assert a.peer_error() == 0
        if rank == 0:
            upd = 172 * Vr * Pr
            print(f"MD_CHECK_TIMING world={world} deals={D} visits_per_rank={Vr} pairs={Pr} "
                  f"sharded_ms={out['sharded']:.3f} private_ms={out['private']:.3f} "
                  f"sharded_G_updates_per_s={upd * world / out['sharded'] / 1e6:.1f} "
                  f"private_per_gpu_G_updates_per_s={upd / out['private'] / 1e6:.1f}", flush=True)
        dist.barrier()
        del a, b
    if rank == 0:
        print(f"MD_CHECK_OK world={world} infosets={len(keys)}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
