"""The ``pim_ops`` operator set, backed by libbackend_pim.so through ctypes.

In the reference these are TORCH_LIBRARY custom ops registered by whichever per-configuration
plugin ``torch.ops.load_library(args.lib_path)`` loads (spmm_default/pytorch_api.cpp:372-389,
spmm_grande/pytorch_api.cpp, spmm_multigroup/pytorch_api.cpp, spmv_sparseP/pytorch_api.cpp).  Here
one library serves every dtype and format; the ops keep their names, argument meaning and return
values, are callable as plain functions from this module, and are also registered under
``torch.ops.pim_ops`` (``register_torch_ops``) because the reference's drivers call
``torch.ops.pim_ops.dpu_init_ranks / dpu_init_dpus / dpu_release`` directly
(spmm_test.py:112-118,136; inference.py:135-141,171).

Tensors may live on the host or on the GPU:
* sparse index/value tensors on the host are uploaded once (the reference's copy_sparse_*);
  CUDA tensors are borrowed without a copy and kept alive by the handle registry;
* a dense operand on the host goes through the host entry point (H2D, kernels, D2H - the result
  is a host tensor, as in the reference where everything is `device = 'cpu'`, spmm_test.py:16);
  a CUDA operand runs asynchronously on the current torch stream and returns a CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

from .. import _lib

# TORCH_TYPES of backend_pim/spmm.py:141 <-> pygim_dtype_t
DTYPE_CODE = {torch.int8: _lib.INT8, torch.int16: _lib.INT16, torch.int32: _lib.INT32, torch.int64: _lib.INT64,
              torch.float32: _lib.FLT32, torch.float64: _lib.DBL64}


@dataclass
class _GroupMeta:
    fmt: int
    dtype: torch.dtype
    total_rows: int
    total_cols: int
    h_size: int
    dense_cols: List[int]
    device: torch.device
    keep_alive: list = field(default_factory=list)
    # grande: per sparse part, the widths its B row block is sliced into (Tensor[] dense_cols)
    grande_cols: Optional[List[List[int]]] = None
    part_ncols: Optional[List[int]] = None
    variant: str = "spmm"   # spmm | grande | spmv
    # the column tiles the LIBRARY plans (spmv: one `groups`-wide tile although the op is handed `groups` vectors)
    lib_dense_cols: Optional[List[int]] = None


_GROUPS: Dict[int, _GroupMeta] = {}
_STATE = {"initialised": False, "nr_ranks": 0, "device": None}


def _cuda_device_index() -> int:
    if not torch.cuda.is_available():
        return -1   # the library reports PYGIM_ERR_NO_DEVICE itself
    return torch.cuda.current_device()


# ------------------------------------------------------------------------------------ bring-up
def dpu_init_ranks(nr_ranks: int, groups_per_rank: int = 1) -> List[int]:
    """spmm_default/pytorch_api.cpp:154-156; returns the per-rank unit counts like the grande
    build does (spmm_grande/pytorch_api.cpp:157-181) - callers of the default build ignore it."""
    lib = _lib.lib()
    out = (C.c_int32 * max(int(nr_ranks), 1))()
    _lib.check(lib.pygim_dpu_init_ranks(int(nr_ranks), int(groups_per_rank), _cuda_device_index(), out))
    _STATE.update(initialised=True, nr_ranks=int(nr_ranks), device=_cuda_device_index())
    return [int(v) for v in out[: int(nr_ranks)]]


def dpu_init_dpus(nr_dpus: int) -> None:
    """spmm_default/pytorch_api.cpp:158-160."""
    _lib.check(_lib.lib().pygim_dpu_init_dpus(int(nr_dpus), _cuda_device_index()))
    _STATE.update(initialised=True, nr_ranks=1, device=_cuda_device_index())


def dpu_release() -> None:
    """spmm_default/pytorch_api.cpp:162-164.  Plans still alive are freed too (the reference leaks them); handles
    are never reused, so a front-end object that outlives the release just holds an unknown handle."""
    for h in list(_GROUPS):
        spmm_free_group(h)
    _lib.check(_lib.lib().pygim_dpu_release())
    _STATE.update(initialised=False, nr_ranks=0)


def device_info() -> dict:
    lib = _lib.lib()
    dev, sm, maj, mnr = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    l2, pl2, hbm = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(lib.pygim_device_info(C.byref(dev), C.byref(sm), C.byref(l2), C.byref(pl2), C.byref(hbm),
                                     C.byref(maj), C.byref(mnr)))
    return {"device": dev.value, "sm_count": sm.value, "l2_bytes": l2.value, "persisting_l2_max_bytes": pl2.value,
            "hbm_bytes": hbm.value, "cc": (maj.value, mnr.value)}


# ------------------------------------------------------------------------------------ plans
def _ptr_array(tensors: Sequence[torch.Tensor]):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _i64_array(values: Sequence[int]):
    return (C.c_int64 * len(values))(*[int(v) for v in values])


def _to_device_group(fmt: int, row_indices, col_indices, values, nrows, ncols, dense_cols, h_size,
                     variant: str = "spmm", grande_cols=None) -> int:
    n_sp = len(values)
    assert n_sp == len(row_indices) == len(col_indices) == len(nrows) == len(ncols)
    dtype = values[0].dtype
    if dtype not in DTYPE_CODE:
        raise _lib.PygimError("unsupported data type %s" % dtype)
    on_gpu = values[0].is_cuda
    rows = [t.to(torch.int32).contiguous() for t in row_indices]
    cols = [t.to(torch.int32).contiguous() for t in col_indices]
    vals = [t.contiguous() for t in values]
    for t in rows + cols + vals:
        if t.is_cuda != on_gpu:
            raise _lib.PygimError("sparse index/value tensors must all live on the same device")
    if any(v.dtype != dtype for v in vals):
        raise _lib.PygimError("sparse parts must share one value dtype")
    handle = C.c_uint64(0)
    lib = _lib.lib()
    if on_gpu:
        torch.cuda.current_stream(vals[0].device).synchronize()   # borrowed arrays must be complete
    # spmv: the `groups` single-column vectors of one call are ONE groups-wide tile for the library (one launch,
    # A streamed once) - the per-vector list only describes the op's argument list
    lib_cols = [int(h_size)] if variant == "spmv" else [int(c) for c in dense_cols]
    _lib.check(lib.pygim_spmm_to_device_group(
        fmt, DTYPE_CODE[dtype], n_sp, _ptr_array(rows), _ptr_array(cols), _ptr_array(vals), _i64_array(nrows),
        _i64_array(ncols), _i64_array([v.numel() for v in vals]), len(lib_cols), _i64_array(lib_cols),
        int(h_size), _lib.MEM_DEVICE if on_gpu else _lib.MEM_HOST, C.byref(handle)))
    dev = vals[0].device if on_gpu else torch.device("cuda", max(_cuda_device_index(), 0))
    _GROUPS[handle.value] = _GroupMeta(
        fmt=fmt, dtype=dtype, total_rows=int(nrows[0]), total_cols=int(sum(ncols)), h_size=int(h_size),
        dense_cols=[int(c) for c in dense_cols], device=dev, keep_alive=(rows + cols + vals) if on_gpu else [],
        grande_cols=grande_cols, part_ncols=[int(c) for c in ncols], variant=variant, lib_dense_cols=lib_cols)
    return int(handle.value)


def spmm_csr_to_device_group(row_indices, col_indices, values, nrows, ncols, dense_cols, h_size) -> int:
    """spmm_default/pytorch_api.cpp:204-243.  With ``dense_cols`` a list of int32 tensors (one per
    sparse part) this is the grande build's op (spmm_grande/pytorch_api.cpp:221-264)."""
    if len(dense_cols) and isinstance(dense_cols[0], torch.Tensor):
        gcols = [[int(v) for v in t.tolist()] for t in dense_cols]
        for g in gcols:
            if sum(g) != int(h_size):
                raise _lib.PygimError("grande: per-rank column widths must sum to hidden_size")
        # on the GPU the per-DPU column slices are only a tiling: one full-width tile per sparse part
        return _to_device_group(_lib.CSR, row_indices, col_indices, values, nrows, ncols, [int(h_size)], h_size,
                                variant="grande", grande_cols=gcols)
    return _to_device_group(_lib.CSR, row_indices, col_indices, values, nrows, ncols, dense_cols, h_size)


def spmm_coo_to_device_group(row_indices, col_indices, values, nrows, ncols, dense_cols, h_size) -> int:
    """spmm_default/pytorch_api.cpp:286-329."""
    return _to_device_group(_lib.COO, row_indices, col_indices, values, nrows, ncols, dense_cols, h_size)


def spmv_coo_to_device_group(row_indices, col_indices, values, nrows, ncols, dense_cols, h_size,
                             ranks_per_spmv: int = 1) -> int:
    """spmv_sparseP/pytorch_api.cpp:184-229: `h_size` single-column dense parts per call."""
    return _to_device_group(_lib.COO, row_indices, col_indices, values, nrows, ncols, dense_cols, h_size,
                            variant="spmv")


def spmm_free_group(handle: int) -> None:
    """spmm_default/pytorch_api.cpp:198-201."""
    if handle in _GROUPS:
        del _GROUPS[handle]
        _lib.check(_lib.lib().pygim_spmm_free_group(int(handle)))


def plan_set_option(handle: int, key: str, value: int) -> None:
    _lib.check(_lib.lib().pygim_plan_set_option(int(handle), key.encode(), int(value)))


def plan_layout(handle: int, part: int = 0) -> dict:
    out = (C.c_int64 * 6)()
    _lib.check(_lib.lib().pygim_plan_layout(int(handle), int(part), out))
    keys = ["items", "supertickets", "coo_runs_as_csr", "coo_sorted", "row_map", "unit_values"]
    return dict(zip(keys, [int(v) for v in out]))


def plan_set_row_map(handle: int, row_map: Optional[torch.Tensor]) -> None:
    """Plan row r is row row_map[r] of the result (pygim_plan_set_row_map); None clears the map."""
    if row_map is None:
        _lib.check(_lib.lib().pygim_plan_set_row_map(int(handle), None, 0, _lib.MEM_HOST))
        return
    rm = row_map.to(torch.int32).contiguous()
    _lib.check(_lib.lib().pygim_plan_set_row_map(int(handle), C.c_void_p(rm.data_ptr()), rm.numel(),
                                                 _lib.MEM_DEVICE if rm.is_cuda else _lib.MEM_HOST))


def plan_set_hot_tiles(handle: int, super_rows: torch.Tensor, hot_cols: torch.Tensor, hot_cnt: torch.Tensor) -> None:
    """pygim_plan_set_hot_tiles: supertickets at `super_rows` (int32 [S+1]), their tile columns `hot_cols`
    (int32 [S x K], -1 = unused) and the per-row hot counts (int32 [nrows]); all on one device."""
    sr, hc, hn = (t.to(torch.int32).contiguous() for t in (super_rows, hot_cols, hot_cnt))
    if not (sr.is_cuda == hc.is_cuda == hn.is_cuda):
        raise _lib.PygimError("hot-tile arrays must live on one device")
    if sr.is_cuda:
        torch.cuda.current_stream(sr.device).synchronize()
    _lib.check(_lib.lib().pygim_plan_set_hot_tiles(int(handle), sr.numel() - 1, C.c_void_p(sr.data_ptr()), int(hc.size(1)),
                                                   C.c_void_p(hc.data_ptr()), C.c_void_p(hn.data_ptr()),
                                                   _lib.MEM_DEVICE if sr.is_cuda else _lib.MEM_HOST))


def plan_stats(handle: int, part: int = 0) -> dict:
    out = (C.c_int64 * 8)()
    _lib.check(_lib.lib().pygim_plan_stats(int(handle), int(part), out))
    keys = ["nrows", "ncols", "nnz", "max_row_nnz", "long_rows", "segments", "seg_len", "empty_rows"]
    return dict(zip(keys, [int(v) for v in out]))


def last_timers(handle: int) -> dict:
    out = (C.c_double * 5)()
    _lib.check(_lib.lib().pygim_last_timers(int(handle), out))
    keys = ["load_sparse_time", "load_dense_time", "kernel_time", "retrieve_result_time", "alignment_time"]
    return dict(zip(keys, [float(v) for v in out]))


def last_launches(handle: int) -> int:
    out = C.c_int64(0)
    _lib.check(_lib.lib().pygim_last_launches(int(handle), C.byref(out)))
    return int(out.value)


# ------------------------------------------------------------------------------------ run
def _meta(handle: int) -> _GroupMeta:
    try:
        return _GROUPS[int(handle)]
    except KeyError:
        raise _lib.PygimError("unknown or freed plan handle %r" % (handle,))


def _run_group(handle: int, B_parts: Sequence[torch.Tensor]) -> torch.Tensor:
    m = _meta(handle)
    if len(B_parts) != len(m.dense_cols):   # assert at pytorch_api.cpp:252
        raise _lib.PygimError("expected %d dense parts, got %d" % (len(m.dense_cols), len(B_parts)))
    if m.lib_dense_cols is not None and m.lib_dense_cols != m.dense_cols:
        # spmv: `groups` vectors [N x 1] -> one [N x groups] operand, one launch
        for b in B_parts:
            if b.dim() != 2 or b.size(1) != 1 or b.size(0) != m.total_cols:
                raise _lib.PygimError("dense part has shape %s, expected (%d, 1)" % (tuple(b.shape), m.total_cols))
        return spmm_run_dense(handle, torch.cat(list(B_parts), dim=1))
    B_parts = [b if b.stride(-1) == 1 or b.numel() == 0 else b.contiguous() for b in B_parts]
    for b, w in zip(B_parts, m.dense_cols):
        if b.dtype != m.dtype:
            raise _lib.PygimError("dense part dtype %s does not match the plan's %s" % (b.dtype, m.dtype))
        if b.dim() != 2 or b.size(1) != w or b.size(0) != m.total_cols:   # asserts at pytorch_api.cpp:264-266
            raise _lib.PygimError("dense part has shape %s, expected (%d, %d)" % (tuple(b.shape), m.total_cols, w))
    lib = _lib.lib()
    ldb = _i64_array([b.stride(0) if b.size(0) > 1 else max(b.size(1), 1) for b in B_parts])
    on_gpu = B_parts[0].is_cuda
    if on_gpu:
        dev = B_parts[0].device
        out = torch.empty((m.total_rows, m.h_size), dtype=m.dtype, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.pygim_spmm_run_group_device(int(handle), len(B_parts), _ptr_array(B_parts), ldb,
                                                       out.data_ptr(), m.h_size, C.c_void_p(stream)))
    else:
        out = torch.empty((m.total_rows, m.h_size), dtype=m.dtype)
        _lib.check(lib.pygim_spmm_run_group_host(int(handle), len(B_parts), _ptr_array(B_parts), ldb,
                                                 out.data_ptr(), m.h_size))
    return out


def spmm_csr_run_group(handle: int, B_parts: Sequence[torch.Tensor]) -> torch.Tensor:
    """spmm_default/pytorch_api.cpp:248-280; for a grande plan B_parts is the flat list of padded
    per-unit column slices that grande.py:83-107 builds (spmm_grande/pytorch_api.cpp:269-321)."""
    m = _meta(handle)
    if m.variant == "grande":
        return _run_group(handle, [_grande_reassemble(m, B_parts)])
    return _run_group(handle, B_parts)


def spmm_coo_run_group(handle: int, B_parts: Sequence[torch.Tensor]) -> torch.Tensor:
    """spmm_default/pytorch_api.cpp:332-367."""
    return _run_group(handle, B_parts)


def spmv_coo_run_group(handle: int, B_parts: Sequence[torch.Tensor]) -> torch.Tensor:
    """spmv_sparseP/pytorch_api.cpp:231-266: B_parts are `groups` vectors [N_pad x 1]; the result is
    [N_pad x groups].  The vectors are computed together as ONE `groups`-column launch (the plan holds a single
    groups-wide tile, so A is streamed once per call, not once per vector)."""
    return _run_group(handle, B_parts)


def _check_out(out: torch.Tensor, m: _GroupMeta, like: torch.Tensor) -> torch.Tensor:
    if out.dtype != m.dtype or tuple(out.shape) != (m.total_rows, m.h_size) or out.is_cuda != like.is_cuda or \
            (out.numel() and out.stride(1) != 1):
        raise _lib.PygimError("out= must be a row-major (%d, %d) %s tensor on the operand's device"
                              % (m.total_rows, m.h_size, m.dtype))
    return out


def spmm_run_dense(handle: int, B: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fast path used by SparseTensorCOO.mul: the whole [sum ncols x h_size] operand in one piece, so
    the dense_split copies of spmm.py:9-13 are not needed - dense parts become column tiles.
    `out` (optional) is a preallocated result, e.g. a pinned host tensor for full-rate D2H."""
    m = _meta(handle)
    if B.dtype != m.dtype:
        raise _lib.PygimError("dense operand dtype %s does not match the plan's %s" % (B.dtype, m.dtype))
    if B.dim() != 2 or B.size(1) != m.h_size or B.size(0) != m.total_cols:
        raise _lib.PygimError("dense operand has shape %s, expected (%d, %d)" % (tuple(B.shape), m.total_cols, m.h_size))
    if B.stride(1) != 1 and B.numel():
        B = B.contiguous()
    lib = _lib.lib()
    ldb = B.stride(0) if B.size(0) > 1 else max(B.size(1), 1)
    if B.is_cuda:
        out = torch.empty((m.total_rows, m.h_size), dtype=m.dtype, device=B.device) if out is None \
            else _check_out(out, m, B)
        ldc = out.stride(0) if out.size(0) > 1 else max(m.h_size, 1)
        with torch.cuda.device(B.device):
            stream = torch.cuda.current_stream(B.device).cuda_stream
            _lib.check(lib.pygim_spmm_device(int(handle), B.data_ptr(), ldb, out.data_ptr(), ldc,
                                             C.c_void_p(stream)))
        return out
    # host operand: present the column tiles as views of B (no copies) to the host entry point
    parts, col = [], 0
    for w in (m.lib_dense_cols or m.dense_cols):
        parts.append(B[:, col:col + w])
        col += w
    out = torch.empty((m.total_rows, m.h_size), dtype=m.dtype) if out is None else _check_out(out, m, B)
    ldc = out.stride(0) if out.size(0) > 1 else max(m.h_size, 1)
    _lib.check(lib.pygim_spmm_run_group_host(int(handle), len(parts), _ptr_array(parts),
                                             _i64_array([ldb] * len(parts)), out.data_ptr(), ldc))
    return out


def _dense_operand(m: _GroupMeta, B: torch.Tensor) -> torch.Tensor:
    if not B.is_cuda:
        raise _lib.PygimError("this entry point needs a CUDA operand")
    if B.dtype != m.dtype or B.dim() != 2 or B.size(1) != m.h_size or B.size(0) != m.total_cols:
        raise _lib.PygimError("dense operand has shape %s/%s, expected (%d, %d) %s"
                              % (tuple(B.shape), B.dtype, m.total_cols, m.h_size, m.dtype))
    if B.stride(1) != 1 and B.numel():
        B = B.contiguous()
    return B


def spmm_run_dense_ex(handle: int, B: torch.Tensor, out: Optional[torch.Tensor] = None,
                      scale: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                      residual_coeff: float = 1.0, peer_ptrs: Sequence[int] = (), multicast_ptr: int = 0,
                      ldc: Optional[int] = None, row_offset: int = 0, peer_mask: Optional[torch.Tensor] = None,
                      flag_ptrs: Sequence[int] = (), my_rank: int = 0, epoch: int = 0) -> Optional[torch.Tensor]:
    """pygim_spmm_device_ex: the SpMM with the rest of a conv layer fused into its row store.
    * `scale` (0-dim float32 CUDA tensor): float32 result = float(sum) * scale  (symmetric_dequantize);
    * `residual` (float32 [rows x H]) and `residual_coeff`: result += coeff * residual  (GIN's (1+eps) x);
    * `peer_ptrs` / `multicast_ptr` / `row_offset` / `peer_mask`: fused all-gather of a row-sharded run;
    * `flag_ptrs` / `my_rank` / `epoch`: in-kernel arrival flags (see wait_flags).
    Returns the result tensor (None for a peer run, whose result lives in the peers' buffers)."""
    m = _meta(handle)
    B = _dense_operand(m, B)
    float_out = scale is not None or residual is not None
    out_dtype = torch.float32 if float_out else m.dtype
    epi = _lib.Epilogue()
    keep = []
    if scale is not None:
        if not scale.is_cuda or scale.dtype != torch.float32 or scale.numel() != 1:
            raise _lib.PygimError("scale must be a one-element float32 CUDA tensor")
        epi.scale = scale.data_ptr()
    if residual is not None:
        if not residual.is_cuda or residual.dtype != torch.float32 or residual.dim() != 2 or \
                residual.size(1) != m.h_size or residual.stride(1) != 1:
            raise _lib.PygimError("residual must be a row-major float32 CUDA matrix with %d columns" % m.h_size)
        epi.residual = residual.data_ptr()
        epi.ld_residual = residual.stride(0) if residual.size(0) > 1 else max(m.h_size, 1)
        epi.residual_coeff = float(residual_coeff)
    n_peers = len(peer_ptrs)
    if n_peers:
        ptrs = (C.c_void_p * n_peers)(*[int(p) for p in peer_ptrs])
        keep.append(ptrs)
        epi.C_peers = ptrs
        epi.n_peers = n_peers
        epi.C_multicast = int(multicast_ptr) or None
        epi.row_offset = int(row_offset)
        if peer_mask is not None:
            if not peer_mask.is_cuda or peer_mask.dtype != torch.uint8 or peer_mask.numel() != m.total_rows:
                raise _lib.PygimError("peer_mask must be a uint8 CUDA vector with one entry per plan row")
            epi.row_peer_mask = peer_mask.data_ptr()
        if len(flag_ptrs):
            fl = (C.c_void_p * n_peers)(*[int(p) for p in flag_ptrs])
            keep.append(fl)
            epi.flag_peers = fl
            epi.my_rank = int(my_rank)
            epi.epoch = int(epoch)
        c_ptr, c_ld = None, int(ldc if ldc is not None else m.h_size)
    else:
        if out is None:
            out = torch.empty((m.total_rows, m.h_size), dtype=out_dtype, device=B.device)
        elif out.dtype != out_dtype or tuple(out.shape) != (m.total_rows, m.h_size) or not out.is_cuda or \
                (out.numel() and out.stride(1) != 1):
            raise _lib.PygimError("out= must be a row-major (%d, %d) %s CUDA tensor" % (m.total_rows, m.h_size, out_dtype))
        c_ptr, c_ld = out.data_ptr(), (out.stride(0) if out.size(0) > 1 else max(m.h_size, 1))
    ldb = B.stride(0) if B.size(0) > 1 else max(B.size(1), 1)
    with torch.cuda.device(B.device):
        stream = torch.cuda.current_stream(B.device).cuda_stream
        _lib.check(_lib.lib().pygim_spmm_device_ex(int(handle), B.data_ptr(), ldb, c_ptr, c_ld, C.byref(epi),
                                                   C.c_void_p(stream)))
    return out if not n_peers else None


def spmm_run_dense_peers(handle: int, B: torch.Tensor, peer_ptrs: Sequence[int], multicast_ptr: int, ldc: int,
                         row_offset: int, **kw) -> None:
    """Row-sharded multi-GPU run with the all-gather fused into the kernel epilogue: this rank's rows are stored at
    `row_offset` of every peer's result matrix (`peer_ptrs`: NVLink-mapped base pointers, own buffer included;
    `multicast_ptr`: NVSwitch multicast mapping or 0).  Asynchronous on the current stream; the caller owns the
    cross-rank synchronisation (a barrier, or the arrival flags of spmm_run_dense_ex + wait_flags)."""
    spmm_run_dense_ex(handle, B, peer_ptrs=peer_ptrs, multicast_ptr=multicast_ptr, ldc=ldc, row_offset=row_offset, **kw)


def wait_flags(flags: torch.Tensor, epoch: int) -> None:
    """Enqueue a wait on the current stream until every entry of the int32 CUDA vector `flags` has reached `epoch`."""
    with torch.cuda.device(flags.device):
        stream = torch.cuda.current_stream(flags.device).cuda_stream
        _lib.check(_lib.lib().pygim_wait_flags(C.c_void_p(flags.data_ptr()), flags.numel(), int(epoch),
                                               C.c_void_p(stream)))


def quantize(x: torch.Tensor, dtype: torch.dtype):
    """symmetric_quantize (models/quantize.py:20-38) in two kernels; returns (scale [0-dim float32], x_q).
    Bit-identical to the torch expression for float32 input."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
        raise _lib.PygimError("quantize needs a 2-D float32 CUDA tensor")
    if dtype not in (torch.int8, torch.int16, torch.int32, torch.float32):
        raise _lib.PygimError("quantize supports INT8 / INT16 / INT32 / FLT32, got %s" % dtype)
    if x.stride(1) != 1 and x.numel():
        x = x.contiguous()
    xq = torch.empty(x.shape, dtype=dtype, device=x.device)
    scale = torch.empty((), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib().pygim_quantize(C.c_void_p(x.data_ptr()), x.size(0), x.size(1),
                                             x.stride(0) if x.size(0) > 1 else max(x.size(1), 1), DTYPE_CODE[dtype],
                                             C.c_void_p(xq.data_ptr()), max(x.size(1), 1), C.c_void_p(scale.data_ptr()),
                                             C.c_void_p(stream)))
    return scale, xq


def spmm_run_dense_many(handles: Sequence[int], Bs: Sequence[torch.Tensor], outs: Sequence[torch.Tensor]) -> None:
    """Several host-operand SpMMs (one per hidden size of spmm_test.py's sweep, or per layer) as ONE software
    pipeline (pygim_spmm_run_many_host): uploads run tile by tile in one stream, smallest operand first so the
    kernels start after a fraction of a millisecond; every finished tile downloads while the next computes.
    `Bs` / `outs` are host tensors (pinned for full PCIe rate); results are complete when the call returns.  The
    per-call host entry point (spmm_run_dense with a host operand) exposes the first upload and the last download of
    EVERY call; this exposes one tile's upload and a fraction of one tile's download per batch."""
    assert len(handles) == len(Bs) == len(outs)
    if not handles:
        return
    order = sorted(range(len(handles)), key=lambda k: Bs[k].numel())
    hs, bs, cs, ldbs, ldcs, keep = [], [], [], [], [], []
    for k in order:
        m, B, out = _meta(handles[k]), Bs[k], outs[k]
        if B.is_cuda or out.is_cuda:
            raise _lib.PygimError("spmm_run_dense_many takes host operands")
        if B.dtype != m.dtype or B.dim() != 2 or B.size(1) != m.h_size or B.size(0) != m.total_cols:
            raise _lib.PygimError("dense operand %d has shape %s/%s, expected (%d, %d) %s"
                                  % (k, tuple(B.shape), B.dtype, m.total_cols, m.h_size, m.dtype))
        if B.stride(1) != 1 and B.numel():
            B = B.contiguous()
        keep.append(B)
        out = _check_out(out, m, B)
        hs.append(int(handles[k]))
        bs.append(B.data_ptr())
        cs.append(out.data_ptr())
        ldbs.append(B.stride(0) if B.size(0) > 1 else max(B.size(1), 1))
        ldcs.append(out.stride(0) if out.size(0) > 1 else max(m.h_size, 1))
    n = len(hs)
    _lib.check(_lib.lib().pygim_spmm_run_many_host(n, (C.c_uint64 * n)(*hs), (C.c_void_p * n)(*bs), _i64_array(ldbs),
                                                   (C.c_void_p * n)(*cs), _i64_array(ldcs)))


def _grande_reassemble(m: _GroupMeta, B_parts: Sequence[torch.Tensor]) -> torch.Tensor:
    """Undo grande.dense_split (grande.py:12-23): for sparse part i the next len(cols_i) entries are
    its row block's column slices, each `pad` columns wide (the tail beyond the slice's real width is
    overlap/padding)."""
    blocks, k = [], 0
    for cols, nrow in zip(m.grande_cols, m.part_ncols):
        pieces = []
        for w in cols:
            pieces.append(B_parts[k][:, :w])
            k += 1
        blk = torch.cat(pieces, dim=1) if len(pieces) > 1 else pieces[0]
        assert blk.size(0) == nrow
        blocks.append(blk)
    if k != len(B_parts):
        raise _lib.PygimError("grande: expected %d column slices, got %d" % (k, len(B_parts)))
    return (torch.cat(blocks, dim=0) if len(blocks) > 1 else blocks[0]).contiguous()


# ------------------------------------------------------------------------------------ partitioners
def partition_rows_by_nnz(rowptr: torch.Tensor, nparts: int) -> List[int]:
    rp = rowptr.to(torch.int32).cpu().contiguous()
    out = (C.c_int64 * (nparts + 1))()
    _lib.check(_lib.lib().pygim_partition_rows_by_nnz(C.c_void_p(rp.data_ptr()), rp.numel() - 1, nparts, out))
    return [int(v) for v in out]


def partition_rows_even(nrows: int, nparts: int) -> List[int]:
    out = (C.c_int64 * (nparts + 1))()
    _lib.check(_lib.lib().pygim_partition_rows_even(int(nrows), nparts, out))
    return [int(v) for v in out]


# ------------------------------------------------------------------------------------ torch.ops.pim_ops
_TORCH_LIB = None


def register_torch_ops() -> None:
    """Expose the ops as torch.ops.pim_ops.* (schemas of spmm_default/pytorch_api.cpp:372-389 plus the
    grande / spmv variants).  Idempotent; a second registration attempt by another loader is ignored."""
    global _TORCH_LIB
    if _TORCH_LIB is not None:
        return
    try:
        lib = torch.library.Library("pim_ops", "DEF")
    except Exception:   # namespace already defined in this process (e.g. a real PyGim plugin)
        return
    defs = [
        ("dpu_init_ranks(int nr_ranks, int groups_per_rank=1) -> int[]", dpu_init_ranks),
        ("dpu_init_dpus(int nr_dpus) -> ()", dpu_init_dpus),
        ("dpu_release() -> ()", dpu_release),
        ("spmm_free_group(int sp_group_ptr) -> ()", spmm_free_group),
        ("spmm_csr_to_device_group(Tensor[] row_indices, Tensor[] col_indices, Tensor[] values, int[] nrows, "
         "int[] ncols, int[] dense_cols, int h_size) -> int", spmm_csr_to_device_group),
        ("spmm_csr_to_device_group.grande(Tensor[] row_indices, Tensor[] col_indices, Tensor[] values, int[] nrows, "
         "int[] ncols, Tensor[] dense_cols, int h_size) -> int", spmm_csr_to_device_group),
        ("spmm_coo_to_device_group(Tensor[] row_indices, Tensor[] col_indices, Tensor[] values, int[] nrows, "
         "int[] ncols, int[] dense_cols, int h_size) -> int", spmm_coo_to_device_group),
        ("spmv_coo_to_device_group(Tensor[] row_indices, Tensor[] col_indices, Tensor[] values, int[] nrows, "
         "int[] ncols, int[] dense_cols, int h_size, int ranks_per_spmv=1) -> int", spmv_coo_to_device_group),
        ("spmm_csr_run_group(int sp_group_ptr, Tensor[] B_parts) -> Tensor", spmm_csr_run_group),
        ("spmm_coo_run_group(int sp_group_ptr, Tensor[] B_parts) -> Tensor", spmm_coo_run_group),
        ("spmv_coo_run_group(int sp_group_ptr, Tensor[] B_parts) -> Tensor", spmv_coo_run_group),
    ]
    for schema, fn in defs:
        lib.define(schema)
        name = schema.split("(")[0]
        lib.impl(name, fn, "CompositeExplicitAutograd")
    _TORCH_LIB = lib
