"""TEST INFRASTRUCTURE ONLY - Python face of the CPU oracle.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product package ``pygim_b200`` never does (tests/test_no_oracle_in_product.py
enforces it).

Two families of entry points:

* ``spmm_*``  - our restatement (``oracle/spmm_oracle.c`` -> ``liboracle.so``) of the arithmetic the
  reference defines (file:line anchors are in the C file's header).
* ``ref_*``   - the reference's *own* scalar host oracles, compiled in place from
  ``/root/reference`` by ``oracle/build_ref.sh`` into ``oracle/_ref/``; used to pin the
  restatement (tests/test_oracle_vs_reference.py) and to generate ``tests/golden``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# dtype table: backend_pim/spmm_default/support/common.h:39-60
SUFFIX = {np.dtype(np.int8): "i8", np.dtype(np.int16): "i16", np.dtype(np.int32): "i32",
          np.dtype(np.int64): "i64", np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
REF_DEFINE = {np.dtype(np.int8): "INT8", np.dtype(np.int16): "INT16", np.dtype(np.int32): "INT32",
              np.dtype(np.int64): "INT64", np.dtype(np.float32): "FLT32", np.dtype(np.float64): "DBL64"}

_lib = None
_ref_libs = {}


def build(native: bool = False, force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    target = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(target) or \
            os.path.getmtime(target) < os.path.getmtime(os.path.join(_HERE, "spmm_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir(os.environ.get("PYGIM_REFERENCE_ROOT", "/root/reference")):
        if force or not os.path.isdir(os.path.join(_HERE, "_ref")) or len(os.listdir(os.path.join(_HERE, "_ref"))) < 19:
            subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return target


def build_native() -> Optional[str]:
    """Best-effort -march=native build for the timed CPU baseline (bench.py); the portable
    liboracle.so stays the checker."""
    out = os.path.join(_HERE, "liboracle_native.so")
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    cmd = [cc, "-std=gnu11", "-O3", "-march=native", "-fPIC", "-fopenmp", "-ffp-contract=off",
           "-fvisibility=hidden", "-shared", "-o", out, os.path.join(_HERE, "spmm_oracle.c")]
    try:
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return out
    except Exception:
        return None


def lib(path: Optional[str] = None) -> C.CDLL:
    global _lib
    if path is not None:
        return C.CDLL(path)
    if _lib is None:
        target = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(target):
            build()
        _lib = C.CDLL(target)
    return _lib


def max_threads() -> int:
    """Host threads available to this process.  Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently serialise the CPU baseline."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _chk(a: np.ndarray, dt=None) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if dt is not None:
        a = a.astype(dt, copy=False)
    return a


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a).astype(np.int32, copy=False))


# ----------------------------------------------------------------------------- restatement
def spmm_coo(rowind, colind, val, x, nrows: int) -> np.ndarray:
    """spmm_default/spmm_mul_coo.c:40-51."""
    x = _chk(x)
    val = _chk(val, x.dtype)
    rowind, colind = _i32(rowind), _i32(colind)
    H = x.shape[1]
    y = np.zeros((nrows, H), dtype=x.dtype)
    f = getattr(lib(), "oracle_spmm_coo_" + SUFFIX[x.dtype])
    f(_p(y), C.c_int64(val.shape[0]), _p(rowind), _p(colind), _p(val), _p(x), C.c_int64(H))
    return y


def spmm_csr(rowptr, colind, values, x, ncols: Optional[int] = None) -> np.ndarray:
    """spmm_grande/spmm_mul_csr.c:119-136 (values used; x row stride = x.shape[1] >= ncols)."""
    x = _chk(x)
    values = _chk(values, x.dtype)
    rowptr, colind = _i32(rowptr), _i32(colind)
    ncols_pad = x.shape[1]
    ncols = ncols_pad if ncols is None else ncols
    nrows = rowptr.shape[0] - 1
    y = np.zeros((nrows, ncols), dtype=x.dtype)
    f = getattr(lib(), "oracle_spmm_csr_" + SUFFIX[x.dtype])
    f(_p(y), C.c_int64(nrows), _p(rowptr), _p(colind), _p(values), _p(x), C.c_int64(ncols), C.c_int64(ncols_pad))
    return y


def spmm_csr_ones(rowptr, colind, x) -> np.ndarray:
    """spmm_default/spmm_mul_csr.c:100-113 (values ignored)."""
    x = _chk(x)
    rowptr, colind = _i32(rowptr), _i32(colind)
    nrows = rowptr.shape[0] - 1
    y = np.zeros((nrows, x.shape[1]), dtype=x.dtype)
    f = getattr(lib(), "oracle_spmm_csr_ones_" + SUFFIX[x.dtype])
    f(_p(y), C.c_int64(nrows), _p(rowptr), _p(colind), _p(x), C.c_int64(x.shape[1]))
    return y


def spmm_group(fmt: str, parts: Sequence[dict], B_parts: Sequence[np.ndarray]) -> np.ndarray:
    """spmm_default/ops.hpp:42-62 / :97-118.  ``parts``: dicts with nrows, ncols and either
    (rowptr, colind, values) [CSR] or (rowind, colind, values) [COO]; ``B_parts``: the
    ``dense_split`` pieces, each [sum(ncols_i) x h_j] contiguous."""
    dt = np.dtype(B_parts[0].dtype)
    B_parts = [_chk(b) for b in B_parts]
    n_sp, n_ds = len(parts), len(B_parts)
    key = "rowptr" if fmt == "CSR" else "rowind"
    rowidx = [_i32(p[key]) for p in parts]
    colind = [_i32(p["colind"]) for p in parts]
    values = [_chk(p["values"], dt) for p in parts]
    nrows = np.array([p["nrows"] for p in parts], dtype=np.int64)
    ncols = np.array([p["ncols"] for p in parts], dtype=np.int64)
    nnz = np.array([v.shape[0] for v in values], dtype=np.int64)
    h = np.array([b.shape[1] for b in B_parts], dtype=np.int64)
    total_cols = int(h.sum())
    y = np.zeros((int(nrows[0]), total_cols), dtype=dt)
    arr = lambda xs: (C.c_void_p * len(xs))(*[x.ctypes.data for x in xs])
    f = getattr(lib(), "oracle_spmm_group_" + SUFFIX[dt])
    f(_p(y), C.c_int(0 if fmt == "CSR" else 1), C.c_int(n_sp), _p(nrows), _p(ncols), _p(nnz),
      arr(rowidx), arr(colind), arr(values), C.c_int(n_ds), arr(B_parts), _p(h), C.c_int64(total_cols))
    return y


def spmm_csr_rowpar(rowptr, colind, values, x, nthreads: int = 0, out: Optional[np.ndarray] = None,
                    accumulate: bool = False, clib: Optional[C.CDLL] = None) -> np.ndarray:
    """Row-parallel CSR SpMM (the `--version=cpu` algorithm class); values=None => implicit ones."""
    x = _chk(x)
    rowptr, colind = _i32(rowptr), _i32(colind)
    if values is not None:
        values = _chk(values, x.dtype)
    nrows, H = rowptr.shape[0] - 1, x.shape[1]
    y = np.zeros((nrows, H), dtype=x.dtype) if out is None else out
    f = getattr(clib or lib(), "oracle_spmm_csr_rowpar_" + SUFFIX[x.dtype])
    f(_p(y), C.c_int64(H), C.c_int64(nrows), _p(rowptr), _p(colind), _p(values), _p(x), C.c_int64(H),
      C.c_int64(H), C.c_int(1 if accumulate else 0), C.c_int(nthreads if nthreads > 0 else max_threads()))
    return y


def spmm_csr_f32_exact(rowptr, colind, values, x, nthreads: int = 0):
    """(f64-accumulated result, sum_e |a_e x_e|) for the float tolerance test."""
    x = _chk(x, np.float32)
    rowptr, colind = _i32(rowptr), _i32(colind)
    if values is not None:
        values = _chk(values, np.float32)
    nrows, H = rowptr.shape[0] - 1, x.shape[1]
    y = np.empty((nrows, H), dtype=np.float64)
    mag = np.empty((nrows, H), dtype=np.float64)
    lib().oracle_spmm_csr_f32_exact(_p(y), _p(mag), C.c_int64(nrows), _p(rowptr), _p(colind), _p(values), _p(x),
                                    C.c_int64(H), C.c_int64(H), C.c_int(nthreads if nthreads > 0 else max_threads()))
    return y, mag


# ----------------------------------------------------------------------------- reference objects
def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_default_FLT32.so"))


def ref_lib(variant: str, dtype) -> C.CDLL:
    key = (variant, np.dtype(dtype))
    if key not in _ref_libs:
        path = os.path.join(_HERE, "_ref", "libref_%s_%s.so" % (variant, REF_DEFINE[np.dtype(dtype)]))
        _ref_libs[key] = C.CDLL(path)
    return _ref_libs[key]


class _RefCOO(C.Structure):  # backend_pim/spmm_default/support/matrix.h:10-19
    _fields_ = [("nrows", C.c_uint32), ("ncols", C.c_uint32), ("nnz", C.c_uint32), ("rows", C.c_void_p),
                ("rowind", C.c_void_p), ("colind", C.c_void_p), ("val", C.c_void_p), ("nnz_size", C.c_uint32)]


class _RefCSR(C.Structure):  # backend_pim/spmm_default/support/matrix.h:23-33
    _fields_ = [("nrows", C.c_uint32), ("ncols", C.c_uint32), ("nnz", C.c_uint32), ("rowptr", C.c_void_p),
                ("colind", C.c_void_p), ("values", C.c_void_p), ("rowptr_size", C.c_uint32),
                ("colind_size", C.c_uint32), ("values_size", C.c_uint32)]


def ref_spmm_host_coo(rowind, colind, val, x, nrows: int, variant: str = "default") -> np.ndarray:
    """The reference's spmm_host_coo (spmm_default/spmm_mul_coo.c:40-51) or, for variant='spmv',
    spmm_host (spmv_sparseP/spmv_mul_coo.c:92-103), executed from oracle/_ref."""
    x = _chk(x)
    val = _chk(val, x.dtype)
    rowind, colind = _i32(rowind), _i32(colind)
    y = np.zeros((nrows, x.shape[1]), dtype=x.dtype)
    A = _RefCOO(nrows, x.shape[0], val.shape[0], None, rowind.ctypes.data, colind.ctypes.data, val.ctypes.data,
                val.shape[0])
    l = ref_lib(variant, x.dtype)
    f = l.spmm_host if variant == "spmv" else l.spmm_host_coo
    f(_p(y), C.byref(A), _p(x), C.c_uint32(x.shape[1]))
    return y


def ref_spmm_host_csr(rowptr, colind, values, x, variant: str = "grande", ncols: Optional[int] = None) -> np.ndarray:
    """grande: spmm_grande/spmm_mul_csr.c:119-136 (values used, padded x stride);
    default: spmm_default/spmm_mul_csr.c:100-113 (values ignored)."""
    x = _chk(x)
    values = _chk(values, x.dtype)
    rowptr, colind = _i32(rowptr), _i32(colind)
    nrows = rowptr.shape[0] - 1
    ncols_pad = x.shape[1]
    ncols = ncols_pad if ncols is None else ncols
    y = np.zeros((nrows, ncols), dtype=x.dtype)
    A = _RefCSR(nrows, x.shape[0], values.shape[0], rowptr.ctypes.data, colind.ctypes.data, values.ctypes.data,
                rowptr.shape[0], colind.shape[0], values.shape[0])
    l = ref_lib(variant, x.dtype)
    if variant == "grande":
        l.spmm_host_csr(_p(y), C.byref(A), _p(x), C.c_uint32(ncols), C.c_uint32(ncols_pad))
    else:
        assert ncols == ncols_pad
        l.spmm_host_csr(_p(y), C.byref(A), _p(x), C.c_uint32(ncols))
    return y


def ref_add_2d(A: np.ndarray, B: np.ndarray, off_x: int, off_y: int) -> np.ndarray:
    """Reference add_2D (spmm_default/spmm_mul_csr.c:77-86): A[off_x+i, off_y+j] += B[i, j] in place."""
    assert A.dtype == B.dtype and A.flags.c_contiguous and B.flags.c_contiguous
    ref_lib("default", A.dtype).add_2D(_p(A), _p(B), C.c_uint32(A.shape[1]), C.c_uint32(B.shape[1]),
                                       C.c_uint32(off_x), C.c_uint32(off_y), C.c_uint32(B.shape[0]),
                                       C.c_uint32(B.shape[1]))
    return A


def ref_partition_rows(rowptr, nparts: int, policy: str = "row") -> np.ndarray:
    """The reference's partition_by_row_csr / partition_by_nnz_csr (spmm_default/support/partition.c:14-44, 51-99)."""
    rowptr = _i32(rowptr)
    nrows = rowptr.shape[0] - 1
    A = _RefCSR(nrows, 0, int(rowptr[-1]), rowptr.ctypes.data, None, None, rowptr.shape[0], 0, 0)
    out = np.zeros(nparts + 2, dtype=np.uint32)
    l = C.CDLL(os.path.join(_HERE, "_ref", "libref_partition.so"))
    f = l.partition_by_row_csr if policy == "row" else l.partition_by_nnz_csr
    f(C.byref(A), _p(out), C.c_int(nparts))
    return out[: nparts + 1].astype(np.int64)
