"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol the header declares
(and nothing the binding does not know), and the host-side codecs round-trip.  No compute calls."""
import os
import re
import subprocess

import numpy as np

from conftest import ROOT, load_golden_json
from scopa_b200 import _build, _lib, codec


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "scopa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    _build.build_library()
    lib = _lib.load()
    assert lib.ms_abi_version() == 1
    declared = _header_symbols()
    assert declared == _lib.exported_symbols()
    out = subprocess.run(["nm", "-D", "--defined-only", _build.LIB], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\bT (ms_[a-z0-9_]+)\b", out)))
    assert exported == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_library_is_sm100a_only_and_has_no_torch_types():
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    hdr = open(os.path.join(ROOT, "include", "scopa_b200.h")).read()
    assert "torch" not in hdr.lower() and "at::" not in hdr


def test_compute_entry_points_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    seeds = np.arange(4, dtype=np.int64)
    st = np.zeros((4, 4), dtype=np.uint32)
    ho = np.zeros(4, dtype=np.uint32)
    rc = lib.ms_deal_from_seeds_host(seeds.ctypes.data, 4, st.ctypes.data, ho.ctypes.data)
    assert rc == -1 and b"failed" in lib.ms_last_error()      # MS_ERR_CUDA: no device, no CPU fallback


def test_codec_round_trips_reference_strings():
    nodes = load_golden_json("env_tree_seed42.json.gz")["nodes"]
    ho = codec.pack_nibbles([7, 9, 5, 6, 14, 10, 12, 8])        # seed-42 deal
    seen = set()
    for nd in nodes:
        if nd["term"]:
            continue
        for p, k in ((0, "info0"), (1, "info1")):
            s = nd[k]
            if s in seen:
                continue
            seen.add(s)
            key = codec.string_to_key(s)
            assert codec.key_to_string(key, ho) == s
            f = codec.key_fields(key)
            assert f["player"] == p and f["table"] == nd["table"]
            assert codec.hand_in_order(f["hand_mask"], ho, p) == nd["hands"][p]
    assert codec.key_to_string(codec.TERMINAL_KEY, ho) == "TERMINAL"


def test_pack_unpack_state():
    w = codec.pack_state([0x00A1, 0x4400], [3, 9, 12], [0x0006, 0x0810], [1, 2], 5, 1, False, 16)
    u = codec.unpack_state(w)
    assert u == {"hand_mask": [0x00A1, 0x4400], "table": [3, 9, 12], "cap_mask": [0x0006, 0x0810], "scopas": [1, 2],
                 "step_count": 5, "cur": 1, "terminal": False, "max_steps": 16}


def test_pyspiel_compat_registry():
    from scopa_b200 import pyspiel_compat as pyspiel
    from scopa_b200.envs import openspiel_mini_scopa  # noqa: F401  (registers the game; no GPU work at import)
    assert pyspiel.PlayerId.TERMINAL == -4
    assert "load_game" in dir(pyspiel)


def test_argument_validation_needs_no_gpu():
    """Bad arguments are refused before any CUDA call (MS_ERR_ARG = -2), with a message."""
    import ctypes as C
    lib = _lib.load()
    buf = np.zeros(64, dtype=np.uint8)
    p = buf.ctypes.data
    assert lib.ms_step(None, None, None, None, 4, None) == -2 and b"ms_step" in lib.ms_last_error()
    assert lib.ms_step(p, p, None, None, -1, None) == -2
    assert lib.ms_legal_actions(p, p, 2, None, None, None, None, 1, None) == -2          # player must be -1, 0 or 1
    assert lib.ms_rollout_random(None, None, 8, 0, 0, None, None, None, None) == -2
    assert lib.ms_solver_create(None, 0, None) == -2
    assert lib.ms_mlp_forward(None, 0, None, None, None, None, 5, None) == -2
    assert lib.ms_mlp_forward(p, 7, p, p, None, None, 5, None) == -2                      # unknown precision
    assert lib.ms_sdcfr_traverse(None, 0, 0, None, None, 0, 1, 0, 0, None, 0, None, None, None, None, None) == -2
    assert lib.ms_team_step(None, None, None, None, 3, None) == -2
    assert lib.ms_cfr_iterate_many(None, 3, 1, None) == -2
    assert lib.ms_sdcfr_infer_states(None, 5, 0, None, 0, None, None) == -2 and b"ms_sdcfr_infer_states" in lib.ms_last_error()
    assert lib.ms_sdcfr_infer_states(p, 5, 2, p, 0, p, None) == -2                        # player to move must be 0 or 1
    assert lib.ms_sdcfr_infer_states(p, 5, 0, p, 3, p, None) == -2                        # unknown precision
    assert lib.ms_md_create(None, 4, 12, None, C.byref(C.c_void_p())) == -2
    assert lib.ms_md_ipc_export(None, p) == -2 and lib.ms_md_ipc_attach(None, 0, 1, p) == -2
    assert lib.ms_md_peer_barrier(None, None) == -2 and lib.ms_md_peer_error(None, None, None) == -2
    assert lib.ms_md_mccfr_blocked(None, 2, 0, 1, 256, 0, None) == -2
    # empty batches succeed without touching the device
    assert lib.ms_sdcfr_infer_states(None, 0, 0, None, 0, None, None) == 0
    assert lib.ms_step(None, None, None, None, 0, None) == 0
    assert lib.ms_rollout_random_host(None, 0, 0, 0, None, None) == 0
    assert lib.ms_sdcfr_samples_per_traversal(0) == 41 and lib.ms_sdcfr_samples_per_traversal(1) == 41
    assert lib.ms_sdcfr_workspace_bytes(1000) > 1000 * 129 * 16


def test_host_pipeline_stage_schedule():
    """The stage schedule of the *_host rollout entry points (ms_rollout_random_host, ms_full_rollout_random_host): the stages
    tile [0, n) exactly, all but the last are multiples of 128 games (128-byte aligned slices of every array), big calls
    start and end with a quarter-size stage, small calls are plain chunks.  Host logic: no device needed."""
    lib = _lib.load(build_if_missing=True)

    def stages(n):
        out, lo = [], 0
        while lo < n:
            m = lib.ms_debug_host_stage_size(lo, n)
            assert 0 < m <= n - lo
            out.append(m)
            lo += m
        return out

    assert lib.ms_debug_set_host_chunk(0) == 262144
    try:
        for n in (1, 127, 65537, 262144, 300000, 393216, 393217, 1_000_000, 4_000_000, 16_777_216 + 5):
            st = stages(n)
            assert sum(st) == n and all(m % 128 == 0 for m in st[:-1]) and max(st) <= 262144
            if n > 262144 + 2 * 65536:
                assert st[0] == 65536 and st[-1] <= 65536 + 127 and st.count(262144) >= len(st) - 4
            else:
                assert st == [262144] * (n // 262144) + ([n % 262144] if n % 262144 else [])
        assert stages(1_000_000) == [65536, 262144, 262144, 262144, 82560, 65472]
        assert lib.ms_debug_host_stage_size(-1, 10) == 0 and lib.ms_debug_host_stage_size(10, 10) == 0
        assert lib.ms_debug_set_host_chunk(300) == 384              # rounded up to 128 games
        st = stages(2000)
        assert sum(st) == 2000 and st[0] == 128 and all(m % 128 == 0 for m in st[:-1])
    finally:
        assert lib.ms_debug_set_host_chunk(0) == 262144
