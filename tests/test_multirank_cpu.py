"""World-size-2 test of the multi-GPU host logic on CPU (gloo): traversal ids are sharded with
scopa_b200.sharding.shard_bounds, every rank produces the delta of ITS share (here with the CPU
oracle standing in for the kernel, since there is no GPU in this container), one all-reduce(sum),
and the result equals the single-rank delta of the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _delta_for(lo, n, philox_seed, warm_iters=3):
    from oracle import ms_oracle as ora
    t = ora.Table()
    t.mccfr_populate()
    t.mccfr_iterate(warm_iters, ora.Rng(1, 5))        # same non-trivial starting table on every rank
    _, r0, s0, _, _ = t.arrays()
    out = []
    for p in (0, 1):
        t.set_arrays(r0, s0)
        t.mccfr_batch(p, philox_seed, lo, n)
        _, r1, s1, _, _ = t.arrays()
        out.append((r1 - r0, s1 - s0))
    return out[0][0] + out[1][0], out[0][1] + out[1][1]


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scopa_b200.sharding import allreduce_delta, shard_bounds
    lo, n = shard_bounds(total, rank, world)
    dreg, dstr = _delta_for(lo, n, 99)
    buf = torch.from_numpy(np.concatenate([dreg.ravel(), dstr.ravel()]))
    allreduce_delta(buf)
    if rank == 0:
        q.put(buf.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_exactly():
    from scopa_b200.sharding import shard_bounds
    for total in (0, 1, 7, 64, 1000, 262144):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            ids = [i for lo, n in spans for i in range(lo, lo + n)] if total <= 1000 else None
            assert sum(n for _, n in spans) == total
            if ids is not None:
                assert ids == list(range(total))
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_two_ranks_allreduce_equals_single_rank():
    total, world = 300, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dreg, dstr = _delta_for(0, total, 99)
    want = np.concatenate([dreg.ravel(), dstr.ravel()])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
