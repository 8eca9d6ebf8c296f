"""ctypes binding of libbackend_pim.so (include/pygim_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises - there is no
Python/torch fallback for the aggregation path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB_PATH = os.path.join(_HERE, "libbackend_pim.so")

# pygim_dtype_t / pygim_format_t / pygim_mem_t
INT8, INT16, INT32, INT64, FLT32, DBL64 = range(6)
CSR, COO = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1

# every symbol include/pygim_b200.h declares (tests/test_abi.py checks header == this list == nm -D)
SYMBOLS = [
    "pygim_last_error", "pygim_abi_version", "pygim_dpu_init_ranks", "pygim_dpu_init_dpus", "pygim_dpu_release",
    "pygim_device_info", "pygim_spmm_to_device_group", "pygim_spmm_free_group", "pygim_plan_set_option",
    "pygim_plan_stats", "pygim_spmm_run_group_host", "pygim_spmm_run_many_host", "pygim_spmm_run_group_device", "pygim_spmm_device", "pygim_spmm_device_peers",
    "pygim_last_timers", "pygim_last_launches", "pygim_partition_rows_by_nnz", "pygim_partition_rows_even",
    "pygim_plan_layout", "pygim_plan_set_row_map", "pygim_spmm_device_ex", "pygim_wait_flags", "pygim_quantize", "pygim_plan_set_hot_tiles",
]


class Epilogue(C.Structure):
    """pygim_epilogue_t (include/pygim_b200.h)."""
    _fields_ = [("scale", C.c_void_p), ("residual", C.c_void_p), ("ld_residual", C.c_int64),
                ("residual_coeff", C.c_float), ("C_peers", C.POINTER(C.c_void_p)), ("n_peers", C.c_int),
                ("C_multicast", C.c_void_p), ("row_offset", C.c_int64), ("row_peer_mask", C.c_void_p),
                ("flag_peers", C.POINTER(C.c_void_p)), ("my_rank", C.c_int), ("epoch", C.c_int32)]

_lib: Optional[C.CDLL] = None
_lib_path: Optional[str] = None


class PygimError(RuntimeError):
    pass


def _declare(lib: C.CDLL) -> None:
    vp, i64, i32, ci = C.c_void_p, C.c_int64, C.c_int32, C.c_int
    P = C.POINTER
    lib.pygim_last_error.restype = C.c_char_p
    lib.pygim_last_error.argtypes = []
    lib.pygim_abi_version.restype = ci
    lib.pygim_dpu_init_ranks.argtypes = [i64, i64, ci, P(i32)]
    lib.pygim_dpu_init_dpus.argtypes = [i64, ci]
    lib.pygim_dpu_release.argtypes = []
    lib.pygim_device_info.argtypes = [P(ci), P(ci), P(i64), P(i64), P(i64), P(ci), P(ci)]
    lib.pygim_spmm_to_device_group.argtypes = [ci, ci, ci, P(vp), P(vp), P(vp), P(i64), P(i64), P(i64), ci, P(i64),
                                               i64, ci, P(C.c_uint64)]
    lib.pygim_spmm_free_group.argtypes = [C.c_uint64]
    lib.pygim_plan_set_option.argtypes = [C.c_uint64, C.c_char_p, i64]
    lib.pygim_plan_stats.argtypes = [C.c_uint64, ci, P(i64)]
    lib.pygim_spmm_run_group_host.argtypes = [C.c_uint64, ci, P(vp), P(i64), vp, i64]
    lib.pygim_spmm_run_many_host.argtypes = [ci, P(C.c_uint64), P(vp), P(i64), P(vp), P(i64)]
    lib.pygim_spmm_run_group_device.argtypes = [C.c_uint64, ci, P(vp), P(i64), vp, i64, vp]
    lib.pygim_spmm_device.argtypes = [C.c_uint64, vp, i64, vp, i64, vp]
    lib.pygim_spmm_device_peers.argtypes = [C.c_uint64, vp, i64, P(vp), ci, vp, i64, i64, vp]
    lib.pygim_last_timers.argtypes = [C.c_uint64, P(C.c_double)]
    lib.pygim_last_launches.argtypes = [C.c_uint64, P(i64)]
    lib.pygim_partition_rows_by_nnz.argtypes = [vp, i64, ci, P(i64)]
    lib.pygim_partition_rows_even.argtypes = [i64, ci, P(i64)]
    lib.pygim_plan_layout.argtypes = [C.c_uint64, ci, P(i64)]
    lib.pygim_plan_set_row_map.argtypes = [C.c_uint64, vp, i64, ci]
    lib.pygim_spmm_device_ex.argtypes = [C.c_uint64, vp, i64, vp, i64, P(Epilogue), vp]
    lib.pygim_wait_flags.argtypes = [vp, ci, i32, vp]
    lib.pygim_plan_set_hot_tiles.argtypes = [C.c_uint64, i64, vp, ci, vp, vp, ci]
    lib.pygim_quantize.argtypes = [vp, i64, i64, i64, ci, vp, i64, vp, vp]
    for name in SYMBOLS:
        if name != "pygim_last_error":
            getattr(lib, name).restype = ci


def load(path: Optional[str] = None) -> C.CDLL:
    """dlopen the backend (the role of torch.ops.load_library(args.lib_path), spmm_test.py:111)."""
    global _lib, _lib_path
    path = path or os.environ.get("PYGIM_LIB_PATH")   # tuning variants: python -m pygim_b200.build -D... -o ...
    path = os.path.abspath(path) if path else DEFAULT_LIB_PATH
    if _lib is not None and _lib_path == path:
        return _lib
    if not os.path.exists(path):
        raise PygimError(
            "pygim_b200: %s not found. Build it with `python -m pygim_b200.build` (needs nvcc); "
            "there is no CPU fallback for the aggregation path." % path)
    lib = C.CDLL(path)
    missing = [s for s in SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise PygimError("pygim_b200: %s does not export %s" % (path, ", ".join(missing)))
    _declare(lib)
    _lib, _lib_path = lib, path
    return lib


def lib() -> C.CDLL:
    return _lib if _lib is not None else load()


def loaded_path() -> Optional[str]:
    return _lib_path


def check(status: int) -> None:
    if status != 0:
        msg = lib().pygim_last_error()
        raise PygimError("pygim_b200 [status %d]: %s" % (status, msg.decode() if msg else "unknown error"))
