"""World-size-2 test of the multi-GPU host logic on CPU (gloo): traversal ids are sharded with
scopa_b200.sharding.shard_bounds, every rank produces the delta of ITS share (here with the CPU
oracle standing in for the kernel, since there is no GPU in this container), one all-reduce(sum),
and the result equals the single-rank delta of the whole batch.  A second test does the same with the PRODUCT's
kernels in place of the oracle: every rank runs mccfr_tree_kernel (the headline kernel) on its share through the CTA
emulator of tests/emu, the delta buffers are summed with allreduce_delta, every rank applies the sum with
mccfr_apply_kernel, and the tables equal a single-rank run of the whole batch."""
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


def _emu_solver():
    """the product's solver kernels on the CTA emulator (tests/emu/ms_solver_host.cpp), seed-42 deal, warmed table"""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import emu_build
    from test_solver_host import HostSolver
    vp = C.c_void_p
    L = C.CDLL(emu_build.build_solver_host())
    L.host_solver_build.argtypes = [vp, C.c_uint32, vp, vp, vp]
    L.host_solver_export.argtypes = [vp] * 5
    L.host_solver_tree.argtypes = [vp] * 6
    L.host_mccfr_inplace_tree.argtypes = [C.c_longlong, C.c_ulonglong, C.c_ulonglong]
    L.host_mccfr_batch.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_ulonglong, C.c_ulonglong]
    L.host_solver_delta.argtypes = [vp]
    L.host_solver_set_delta.argtypes = [vp]
    L.host_solver_counters.argtypes = [vp, vp, C.c_int]
    sv = HostSolver(L, 42)
    assert L.host_mccfr_inplace_tree(3, 5, 0) == 0        # same non-trivial starting table on every rank
    return L, sv


def _emu_worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scopa_b200.sharding import allreduce_delta, shard_bounds
    L, sv = _emu_solver()
    lo, n = shard_bounds(total, rank, world)
    assert L.host_mccfr_batch(0, 2, n, 99, lo) == 0       # both players, traversal ids [lo, lo + n)
    buf = torch.zeros(6 * sv.n_slots, dtype=torch.float64)
    L.host_solver_delta(buf.data_ptr())
    allreduce_delta(buf)
    L.host_solver_set_delta(buf.data_ptr())
    assert L.host_mccfr_apply() == 0
    tab = sv.table()
    cnt = np.zeros(3, np.uint64)
    touched = np.zeros(sv.n_slots, np.uint8)
    L.host_solver_counters(cnt.ctypes.data, touched.ctypes.data, 0)
    q.put((rank, tab["regret"], tab["strategy"], buf.numpy()[4 * sv.n_slots:5 * sv.n_slots].copy(), touched))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_of_the_product_kernel_equal_single_rank():
    total, world = 2500, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_emu_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, (reg, strat, cnt, tch)) for r, reg, strat, cnt, tch in (q.get(timeout=180) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas stay identical: every rank applied the same summed delta
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    assert np.array_equal(got[0][3], got[1][3])         # ... including WHICH infosets exist (touched travels with the delta)
    L, sv = _emu_solver()
    assert L.host_mccfr_batch(0, 2, total, 99, 0) == 0
    whole = np.zeros(6 * sv.n_slots)
    L.host_solver_delta(whole.ctypes.data)
    assert L.host_mccfr_apply() == 0
    tab = sv.table()
    S = sv.n_slots
    assert np.array_equal(got[0][2], whole[4 * S:5 * S])                  # update counts: exact integers
    cnt = np.zeros(3, np.uint64)
    touched = np.zeros(S, np.uint8)
    L.host_solver_counters(cnt.ctypes.data, touched.ctypes.data, 0)
    assert np.array_equal(got[0][3], touched) and touched.sum() > 0.9 * S   # same infosets as the single-rank run
    np.testing.assert_allclose(got[0][0], tab["regret"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got[0][1], tab["strategy"], rtol=1e-9, atol=1e-9)
