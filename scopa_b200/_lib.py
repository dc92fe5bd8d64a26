"""ctypes binding of libscopa_b200.so (the C ABI declared in include/scopa_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and if that is not
possible, or a call fails (e.g. no CUDA device), a RuntimeError is raised.
"""
import ctypes as C
import os

from . import _build

_LIB = None

u8p, u16p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64))
vp, i64, u64, i32, dbl = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_double


class MsError(RuntimeError):
    pass


_SIGS = {
    # name: (argtypes, restype)
    "ms_abi_version": ([], C.c_int),
    "ms_last_error": ([], C.c_char_p),
    "ms_launch_count": ([], u64),
    "ms_deal_from_seeds": ([vp, i64, vp, vp, vp], C.c_int),
    "ms_deck_from_seeds": ([vp, i64, vp, vp], C.c_int),
    "ms_debug_deal_slow_path": ([vp, i64, vp, vp, vp], C.c_int),
    "ms_debug_atomic_peaks": ([C.POINTER(dbl), vp], C.c_int),
    "ms_debug_set_host_chunk": ([i64], i64),
    "ms_debug_host_stage_size": ([i64, i64], i64),
    "ms_step": ([vp, vp, vp, vp, i64, vp], C.c_int),
    "ms_legal_actions": ([vp, vp, C.c_int, vp, vp, vp, vp, i64, vp], C.c_int),
    "ms_capture": ([vp, vp, vp, i64, vp], C.c_int),
    "ms_infoset_keys": ([vp, C.c_int, vp, i64, vp], C.c_int),
    "ms_rollout_random": ([vp, vp, i64, u64, u64, vp, vp, vp, vp], C.c_int),
    "ms_deal_from_seeds_host": ([vp, i64, vp, vp], C.c_int),
    "ms_step_host": ([vp, vp, vp, vp, i64], C.c_int),
    "ms_legal_actions_host": ([vp, vp, C.c_int, vp, vp, vp, vp, i64], C.c_int),
    "ms_infoset_keys_host": ([vp, C.c_int, vp, i64], C.c_int),
    "ms_rollout_random_host": ([vp, i64, u64, u64, vp, vp], C.c_int),
    "ms_solver_create": ([vp, C.c_uint32, C.POINTER(vp)], C.c_int),
    "ms_solver_destroy": ([vp], None),
    "ms_solver_reset": ([vp, vp], C.c_int),
    "ms_solver_counts": ([vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], C.c_int),
    "ms_solver_export_tree": ([vp, vp, vp, vp, vp, vp, vp], C.c_int),
    "ms_solver_export_table": ([vp, vp, vp, vp, vp, vp, vp, vp], C.c_int),
    "ms_solver_import_table": ([vp, vp, vp, vp], C.c_int),
    "ms_solver_device_ptrs": ([vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)], C.c_int),
    "ms_cfr_iterate": ([vp, i32, vp], C.c_int),
    "ms_cfr_iterate_many": ([C.POINTER(vp), i32, i32, vp], C.c_int),
    "ms_cfr_traverse": ([vp, i32, dbl, dbl, C.POINTER(dbl), vp], C.c_int),
    "ms_mccfr_inplace": ([vp, i64, u64, u64, vp], C.c_int),
    "ms_mccfr_batch": ([vp, i32, i64, u64, u64, vp], C.c_int),
    "ms_mccfr_apply": ([vp, vp], C.c_int),
    "ms_mccfr_batch_mode": ([vp, i32, i32, i64, u64, u64, vp], C.c_int),
    "ms_solver_ipc_export": ([vp, vp, C.POINTER(u64)], C.c_int),
    "ms_solver_ipc_attach": ([vp, i32, i32, vp, C.POINTER(u64)], C.c_int),
    "ms_mccfr_apply_peers": ([vp, vp], C.c_int),
    "ms_mccfr_batch_peers": ([vp, i32, i64, u64, u64, vp], C.c_int),
    "ms_solver_peer_error": ([vp, C.POINTER(C.c_uint32), vp], C.c_int),
    "ms_mccfr_inplace_many": ([vp, i32, i64, u64, u64, vp, vp, vp, vp], C.c_int),
    "ms_solver_counters": ([vp, C.POINTER(u64), C.c_int, vp], C.c_int),
    "ms_best_response": ([vp, i32, C.POINTER(dbl), vp], C.c_int),
    "ms_solver_policy": ([vp, i32, vp, vp], C.c_int),
    "ms_eval_policies": ([vp, vp, vp, i64, u64, u64, vp, vp, vp], C.c_int),
    "ms_team_deal_from_seeds": ([vp, i64, vp, vp, vp], C.c_int),
    "ms_team_step": ([vp, vp, vp, vp, i64, vp], C.c_int),
    "ms_team_rollout_random": ([vp, vp, i64, u64, u64, vp, vp, vp, vp], C.c_int),
    "ms_team_deal_from_seeds_host": ([vp, i64, vp, vp], C.c_int),
    "ms_team_step_host": ([vp, vp, vp, vp, i64], C.c_int),
    "ms_sdcfr_samples_per_traversal": ([C.c_int], C.c_int),
    "ms_sdcfr_workspace_bytes": ([i64], C.c_size_t),
    "ms_mlp_forward": ([vp, C.c_int, vp, vp, vp, vp, i64, vp], C.c_int),
    "ms_sdcfr_infer_states": ([vp, i64, C.c_int, vp, C.c_int, vp, vp], C.c_int),
    "ms_sdcfr_traverse": ([vp, C.c_uint32, C.c_int, vp, vp, C.c_int, i64, u64, u64, vp, C.c_size_t, vp, vp, vp, vp, vp],
                          C.c_int),
    "ms_sdcfr_train_workspace_bytes": ([], C.c_size_t),
    "ms_sdcfr_train": ([vp, vp, vp, i64, vp, vp, vp, i64, vp, i32, i32, dbl, dbl, dbl, dbl, dbl, vp, vp, C.c_size_t, vp],
                       C.c_int),
    "ms_sdcfr_train_cluster": ([vp, vp, vp, i64, vp, vp, vp, i64, vp, i32, i32, dbl, dbl, dbl, dbl, dbl, vp, vp, C.c_size_t, vp],
                               C.c_int),
    "ms_sdcfr_sample_rows": ([vp, i32, i32, i64, u64, u64, vp], C.c_int),
    "ms_sdcfr_average_policy_workspace_bytes": ([i32, i64], C.c_size_t),
    "ms_sdcfr_average_policy": ([vp, vp, i32, vp, vp, i64, vp, vp, C.c_size_t, vp], C.c_int),
    "ms_full_deal_from_seeds": ([vp, i64, vp, vp, vp], C.c_int),
    "ms_full_deck_from_seeds": ([vp, i64, vp, C.c_int, vp], C.c_int),
    "ms_full_step": ([vp, vp, vp, vp, vp, i64, vp], C.c_int),
    "ms_full_legal_actions": ([vp, vp, C.c_int, vp, vp, i64, vp], C.c_int),
    "ms_full_rollout_random": ([vp, vp, i64, u64, u64, vp, vp, vp, vp], C.c_int),
    "ms_full_table_overflow": ([C.POINTER(C.c_int), vp], C.c_int),
    "ms_full_deal_from_seeds_host": ([vp, i64, vp, vp], C.c_int),
    "ms_full_step_host": ([vp, vp, vp, vp, vp, i64], C.c_int),
    "ms_full_rollout_random_host": ([vp, i64, u64, u64, vp, vp], C.c_int),
    "ms_full_evaluate_host": ([vp, vp, vp, i64], C.c_int),
    "ms_md_create": ([vp, i64, i32, vp, C.POINTER(vp)], C.c_int),
    "ms_md_destroy": ([vp], None),
    "ms_md_reset": ([vp, vp], C.c_int),
    "ms_md_info": ([vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)], C.c_int),
    "ms_md_mccfr_batch": ([vp, i32, i64, u64, u64, vp], C.c_int),
    "ms_md_mccfr_blocked": ([vp, i32, i64, i64, i32, u64, vp], C.c_int),
    "ms_md_apply": ([vp, vp], C.c_int),
    "ms_md_counters": ([vp, C.POINTER(u64), C.c_int, vp], C.c_int),
    "ms_md_export": ([vp, vp, vp, vp, i64, C.POINTER(i64), vp], C.c_int),
    "ms_md_lookup": ([vp, vp, i64, vp, vp, vp, vp], C.c_int),
    "ms_md_ipc_export": ([vp, vp], C.c_int),
    "ms_md_ipc_attach": ([vp, i32, i32, vp], C.c_int),
    "ms_md_peer_barrier": ([vp, vp], C.c_int),
    "ms_md_peer_error": ([vp, C.POINTER(C.c_uint32), vp], C.c_int),
    "ms_debug_random_access_peaks": ([i32, C.POINTER(dbl), vp], C.c_int),
}


def exported_symbols():
    """Every symbol include/scopa_b200.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGS)


def load(build_if_missing=True):
    """Load (building first if needed) libscopa_b200.so; no compute happens here."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("SCOPA_B200_LIB") or _build.LIB      # (override: kernel-variant experiments under profiles/)
    if path == _build.LIB and build_if_missing and _build.stale():
        try:
            _build.build_library()
        except Exception as e:  # on the GPU box nvcc exists too; if not, a prebuilt .so must be there
            if not os.path.exists(path):
                raise MsError(f"libscopa_b200.so is missing and could not be built: {e}") from e
    if not os.path.exists(path):
        raise MsError("libscopa_b200.so is missing (run `python __graft_entry__.py` or scopa_b200/_build.py); "
                      "scopa_b200 has no CPU fallback")
    lib = C.CDLL(path)
    for name, (argtypes, restype) in _SIGS.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = restype
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().ms_last_error().decode(errors="replace")
        raise MsError(f"libscopa_b200 error {rc}: {msg}")


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (default: torch's current stream) as an integer."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return stream.cuda_stream


def launch_count():
    return int(load().ms_launch_count())
