"""The CPU arm's reference: oracle/_ref (the unmodified reference byte-compiled by oracle/make_ref.py) imports behind
oracle/stubs and runs its own loops (oracle/time_reference.py).  Skipped where oracle/_ref has not been built."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REFC = os.path.join(ROOT, "oracle", "_ref", "src", "algorithms", "mc_cfr.refc")


@pytest.mark.skipif(not os.path.exists(REFC), reason="oracle/_ref not built (python oracle/make_ref.py where /root/reference exists)")
def test_time_reference_runs_the_byte_compiled_reference():
    exe = [sys.executable, os.path.join(ROOT, "oracle", "time_reference.py")]
    p = subprocess.run(exe + ["serve"], input="mccfr 2\nenv 50\ncfr 1\nquit\n", capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.loads(ln) for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert lines[0]["ready"] and lines[0]["module_file"].endswith(os.path.join("oracle", "_ref", "src", "algorithms", "mc_cfr.refc"))
    m, e, c = lines[1], lines[2], lines[3]
    assert m["updates"] == 172 * 2 and m["visits"] == 703 * 2 and 1.0 < m["ms_per_iteration"] < 5000.0
    assert e["steps"] == 8 * 50 and e["steps_per_sec"] > 1000
    assert c["ms_per_iteration"] > 1.0


def test_no_reference_source_is_tracked():
    """oracle/_ref is a build output: git-ignored, and it holds bytecode only."""
    with open(os.path.join(ROOT, ".gitignore")) as f:
        assert "oracle/_ref/" in f.read().split()
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(ref):
        for dirpath, _, files in os.walk(ref):
            assert not [f for f in files if f.endswith(".py")], dirpath
