"""Multi-GPU tests proper (collected with -m gpu; skipped on a box with fewer than 2 GPUs): 2 ranks under torchrun.

  * tests/multigpu_peers_check.py: {ms_mccfr_batch, ms_mccfr_apply_peers} (the peer-memory exchange kernel) against
    {ms_mccfr_batch, NCCL all-reduce, ms_mccfr_apply} and against one GPU running every rank's traversal ids; replicas
    bit-identical; touched flags agree; no peer error;
  * tests/multigpu_md_check.py: the multi-deal infoset table sharded over the ranks by a hash of the key (md_blocked_kernel
    gathers regrets from / sends deltas to the owners' shards through peer memory): union of the shards == the table one
    GPU builds from all the visits;
  * bench.py with the driver's exact N = 2 command line (default flags): exits 0 with one JSON line.
"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

two_gpus = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(n, script, *args, timeout=600, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), script, *args]
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=e)


@two_gpus
def test_peer_exchange_equals_nccl_and_single_gpu():
    res = _torchrun(2, os.path.join("tests", "multigpu_peers_check.py"), env={"PEERS_CHECK_TIMING": "0"})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "PEERS_CHECK_OK world=2" in res.stdout


@two_gpus
def test_sharded_multideal_table_equals_single_gpu():
    res = _torchrun(2, os.path.join("tests", "multigpu_md_check.py"), env={"MD_CHECK_TIMING": "0"})
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MD_CHECK_OK world=2" in res.stdout


@two_gpus
def test_bench_default_command_line_on_two_gpus():
    res = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "5", "--warmup", "3", timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["collective"] in ("nccl", "p2p")
    assert d["exchange"]["nccl_ms_per_step"] > 0 and d["exchange"]["p2p_fused_ms_per_step"] > 0
