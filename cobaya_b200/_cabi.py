"""
ctypes binding of libcobaya_b200.so (the C ABI of include/cobaya_b200.h).

There is deliberately no fallback: if the CUDA library is missing or cannot be
loaded, importing the engine fails loudly (BASELINE.json north_star: "no CPU
fallback").
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# COBAYA_B200_LIB: another build of the same library (kernel experiments); never a fallback
LIB_PATH = os.environ.get("COBAYA_B200_LIB") or os.path.join(_HERE, "lib", "libcobaya_b200.so")

EXPORTS = [
    "cb2_abi_version", "cb2_last_error", "cb2_create", "cb2_destroy", "cb2_set_prior",
    "cb2_set_prior_shapes",
    "cb2_clear_likelihoods", "cb2_add_gaussian_mixture", "cb2_add_rosenbrock", "cb2_add_constant",
    "cb2_set_blocking", "cb2_set_proposal", "cb2_set_options", "cb2_set_state",
    "cb2_get_state", "cb2_snapshot_size", "cb2_export_state", "cb2_import_state",
    "cb2_load_rows", "cb2_logpost", "cb2_advance", "cb2_sync", "cb2_summary",
    "cb2_moments", "cb2_bounds", "cb2_copy_rows", "cb2_row_width", "cb2_n_derived", "cb2_debug_basis",
    "cb2_launch_count", "cb2_timer_start", "cb2_timer_stop", "cb2_last_step_kernel",
    "cb2_set_kernel_policy", "cb2_set_profiling", "cb2_kernel_times", "cb2_debug_message",
    "cb2_copy_rows_bulk", "cb2_drain_start", "cb2_drain_wait", "cb2_drain_reset",
    "cb2_host_alloc", "cb2_host_free", "cb2_grow_rows", "cb2_mem_info", "cb2_load_rows_bulk", "cb2_window_counts", "cb2_debug_counters",
    "cb2_checkpoint_device", "cb2_checkpoint_cov", "cb2_adopt_proposal", "cb2_get_proposal",
    "cb2_add_external_likelihood", "cb2_check_external_source", "cb2_measure_speeds",
    "cb2_add_external_prior", "cb2_check_external_fused", "cb2_ext_route_counts",
]


class EngineLibraryError(ImportError):
    pass


_lib = None


def load():
    """Load the shared library and declare prototypes (once)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineLibraryError(
            f"{LIB_PATH} not found. Build it with `python -m cobaya_b200.build` "
            "(needs nvcc); the engine has no CPU fallback."
        )
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise EngineLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    missing = [s for s in EXPORTS if not hasattr(L, s)]
    if missing:
        raise EngineLibraryError(f"{LIB_PATH} lacks symbols {missing}")
    vp, i32, i64, u64, u32, dbl = (C.c_void_p, C.c_int32, C.c_int64, C.c_uint64,
                                   C.c_uint32, C.c_double)
    L.cb2_abi_version.restype = C.c_int
    L.cb2_last_error.restype = C.c_char_p
    L.cb2_last_error.argtypes = [vp]
    L.cb2_create.argtypes = [C.c_int, i64, i32, u64, u64, C.POINTER(vp)]
    L.cb2_destroy.argtypes = [vp]
    L.cb2_set_prior.argtypes = [vp] + [vp] * 6 + [dbl]
    L.cb2_set_prior_shapes.argtypes = [vp, vp, vp, vp]
    L.cb2_clear_likelihoods.argtypes = [vp]
    L.cb2_snapshot_size.restype = i64
    L.cb2_snapshot_size.argtypes = [vp]
    L.cb2_export_state.argtypes = [vp, vp, i64]
    L.cb2_import_state.argtypes = [vp, vp, i64]
    L.cb2_load_rows.argtypes = [vp, i64, i64, vp]
    L.cb2_add_gaussian_mixture.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, i32]
    L.cb2_add_rosenbrock.argtypes = [vp, i32, vp, dbl]
    L.cb2_add_constant.argtypes = [vp, dbl]
    L.cb2_add_external_likelihood.argtypes = [vp, i32, vp, C.c_char_p, C.c_char_p]
    L.cb2_add_external_prior.argtypes = [vp, i32, vp, C.c_char_p, C.c_char_p]
    L.cb2_check_external_source.argtypes = [C.c_char_p, C.c_char_p, i32, C.c_char_p, i64]
    L.cb2_check_external_fused.argtypes = [C.c_char_p, C.c_char_p, i32, C.c_char_p, i64]
    L.cb2_ext_route_counts.argtypes = [vp, vp]
    L.cb2_set_blocking.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32]
    L.cb2_set_proposal.argtypes = [vp, vp, dbl]
    L.cb2_set_options.argtypes = [vp, dbl, i64, i64, i32, i64]
    L.cb2_set_state.argtypes = [vp, vp]
    L.cb2_get_state.argtypes = [vp] + [vp] * 6
    L.cb2_logpost.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.cb2_advance.argtypes = [vp, i64]
    L.cb2_sync.argtypes = [vp]
    L.cb2_summary.argtypes = [vp, vp]
    L.cb2_moments.argtypes = [vp, i32, i32, vp, vp, vp]
    L.cb2_bounds.argtypes = [vp, i32, i32, dbl, vp, vp, vp]
    L.cb2_measure_speeds.argtypes = [vp, vp, i64, i32, vp]
    L.cb2_checkpoint_device.argtypes = [vp, vp, vp]
    L.cb2_checkpoint_cov.argtypes = [vp, vp]
    L.cb2_adopt_proposal.argtypes = [vp]
    L.cb2_get_proposal.argtypes = [vp, vp]
    L.cb2_copy_rows.restype = i64
    L.cb2_copy_rows.argtypes = [vp, i64, i64, i64, vp]
    L.cb2_copy_rows_bulk.restype = i64
    L.cb2_copy_rows_bulk.argtypes = [vp, i64, i64, vp, vp, vp, i64]
    L.cb2_drain_start.restype = i64
    L.cb2_drain_start.argtypes = [vp, vp, i64, vp]
    L.cb2_drain_wait.argtypes = [vp]
    L.cb2_drain_reset.argtypes = [vp, vp]
    L.cb2_host_alloc.restype = vp
    L.cb2_host_alloc.argtypes = [i64]
    L.cb2_host_free.argtypes = [vp]
    L.cb2_grow_rows.argtypes = [vp, i64]
    L.cb2_mem_info.argtypes = [vp, vp, vp, vp]
    L.cb2_window_counts.argtypes = [vp, vp, i32]
    L.cb2_debug_counters.argtypes = [vp, vp, i32]
    L.cb2_load_rows_bulk.argtypes = [vp, i64, i64, vp, vp]
    L.cb2_row_width.restype = i32
    L.cb2_row_width.argtypes = [vp]
    L.cb2_n_derived.restype = i32
    L.cb2_n_derived.argtypes = [vp]
    L.cb2_debug_basis.argtypes = [vp, i64, i32, u32, vp]
    L.cb2_launch_count.restype = i64
    L.cb2_launch_count.argtypes = [vp]
    L.cb2_timer_start.argtypes = [vp]
    L.cb2_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.cb2_last_step_kernel.argtypes = [vp]
    L.cb2_set_kernel_policy.argtypes = [vp, i32]
    L.cb2_debug_message.restype = C.c_char_p
    L.cb2_debug_message.argtypes = [vp]
    L.cb2_set_profiling.argtypes = [vp, i32]
    L.cb2_kernel_times.argtypes = [vp, vp, vp, i32]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("cb2_abi_version",):
            fn.restype = C.c_int
    _lib = L
    return L


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
