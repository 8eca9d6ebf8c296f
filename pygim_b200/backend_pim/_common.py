"""Shared machinery of the three backend front-ends (spmm / grande / spmv).

Each front-end of the reference is a self-contained copy of the same class
(backend_pim/spmm.py:15-136, grande.py:25-121, spmv.py:21-109).  Here the common behaviour lives
once: splitting the adjacency by columns, turning each part into int32 CSR or coalesced COO index
arrays of the requested value dtype, and holding the plan handle.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import pim_ops

TORCH_TYPES = {"INT64": torch.int64, "INT32": torch.int32, "INT16": torch.int16, "INT8": torch.int8,
               "FLT32": torch.float32, "DBL64": torch.float64}


class _Csr:
    """The three arrays of one CSR part, with the accessor names of torch.sparse_csr_tensor."""

    def __init__(self, rowptr, col, value, shape):
        self._rowptr, self._col, self._value, self._shape = rowptr, col, value, tuple(shape)

    def crow_indices(self):
        return self._rowptr

    def col_indices(self):
        return self._col

    def values(self):
        return self._value

    def size(self, dim=None):
        return self._shape if dim is None else self._shape[dim]


class _Coo:
    """Row-major sorted, duplicate-free COO part (what `.coalesce()` yields, spmm.py:40-42)."""

    def __init__(self, row, col, value, shape):
        self._row, self._col, self._value, self._shape = row, col, value, tuple(shape)

    def indices(self):
        return torch.stack([self._row, self._col], dim=0)

    def row_indices(self):
        return self._row

    def col_indices(self):
        return self._col

    def values(self):
        return self._value

    def size(self, dim=None):
        return self._shape if dim is None else self._shape[dim]


def edge_values(item, dtype: torch.dtype) -> torch.Tensor:
    """Missing edge values become ones, present ones are cast with .type(dtype) (spmm.py:36-39,48-51)."""
    value = item.coo()[2]
    if value is None:
        return torch.ones(item.nnz(), dtype=dtype, device=item.device())
    return value.type(dtype)


def coalesce(row: torch.Tensor, col: torch.Tensor, value: torch.Tensor, ncols: int
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Sort row-major and sum duplicate (row, col) entries in the value dtype (wrapping for ints),
    i.e. what coalescing a sparse COO tensor does, without building a sparse tensor."""
    if row.numel() == 0:
        return row, col, value
    key = row.to(torch.int64) * max(int(ncols), 1) + col.to(torch.int64)
    if bool((key[1:] > key[:-1]).all()):
        return row, col, value           # already sorted and unique
    key, perm = torch.sort(key, stable=True)
    value = value[perm]
    uniq, inverse = torch.unique_consecutive(key, return_inverse=True)
    if uniq.numel() != key.numel():
        summed = torch.zeros(uniq.numel(), dtype=value.dtype, device=value.device)
        summed.index_add_(0, inverse, value)
        value = summed
    return uniq // max(int(ncols), 1), uniq % max(int(ncols), 1), value


def split_widths(total: int, nparts: int) -> List[int]:
    """ceil(total/nparts) each, remainder in the last part (spmm.py:60-72, :129-133)."""
    width = (total + nparts - 1) // nparts
    out = [width] * nparts
    if nparts * width != total:
        out[nparts - 1] = total - (nparts - 1) * width
    return out


class SparseTensorBase:
    """State shared by the front-ends; attribute names follow the reference so callers that read
    `.dtype`, `.raw`, `.parts`, `.format`, `.hidden_size`, `.dense_parts`, `.sp_info_ptr` keep working
    (models/pyg_gcn_conv.py:131 reads `.dtype`)."""

    def __init__(self, coo, dtype=torch.int32, format=""):
        self.raw = coo
        self.dtype = dtype
        self.sp_info_ptr: Optional[int] = None
        self.result = None
        self.parts = [self.raw]
        self.dense_parts = 0
        self.csr: list = []
        self.coo: list = []
        self.hidden_size = 0
        self.nparts = 1
        self.format = format
        # row reordering (pygim_b200/reorder.py): `raw` then holds the row-permuted adjacency and row_perm[r] is the
        # original row of its row r; the plan scatters on store, so mul() returns rows in the original order
        self.row_perm = None
        self.plan_options = {}
        self.hot_plan = None       # reorder.hot_cold_plan(...): `raw` is then in hot/cold form (CSR only)

    def _plan_created(self):
        """Attach what belongs to the plan besides the sparse arrays: the row map and the tuned options."""
        if self.row_perm is not None:
            pim_ops.plan_set_row_map(self.sp_info_ptr, self.row_perm)
        for key, value in self.plan_options.items():
            pim_ops.plan_set_option(self.sp_info_ptr, key, value)
        if self.hot_plan is not None:
            assert self.format == "CSR" and len(self.parts) == 1, "hot/cold plans are CSR, sp_parts == 1"
            pim_ops.plan_set_option(self.sp_info_ptr, "seg_len", self.hot_plan["seg_len"])
            pim_ops.plan_set_hot_tiles(self.sp_info_ptr, self.hot_plan["super_rows"], self.hot_plan["hot_cols"],
                                       self.hot_plan["hot_cnt"])

    # -- column split of the adjacency (sparse parts; partial products are summed)
    def col_split(self, nparts=4):
        assert nparts > 0
        width = (self.raw.size(1) + nparts - 1) // nparts
        if nparts != len(self.parts):
            assert len(self.parts) == 1
            pieces = [self.raw[:, k * width:(k + 1) * width] for k in range(nparts - 1)]
            pieces.append(self.raw[:, (nparts - 1) * width:])
            self.parts = pieces
            self.csr, self.coo = [], []
        return self.parts

    def row_split(self, nparts=4):
        assert False   # spmm.py:124-125

    # -- per-part index arrays
    def build_csr(self):
        self.csr = []
        for item in self.parts:
            rowptr, col, _ = item.csr()
            self.csr.append(_Csr(rowptr.int().contiguous(), col.int().contiguous(),
                                 edge_values(item, self.dtype).contiguous(), item.sizes()[:2]))

    def build_coo(self, pad_to: int = 1):
        self.coo = []
        for item in self.parts:
            row, col, _ = item.coo()
            n, m = item.size(0), item.size(1)
            if pad_to > 1 and n % pad_to != 0:     # spmv.py:45-51: BOTH dims grow by the row padding
                pad = pad_to - n % pad_to
                n, m = n + pad, m + pad
            row, col, value = coalesce(row, col, edge_values(item, self.dtype), m)
            self.coo.append(_Coo(row.int().contiguous(), col.int().contiguous(), value.contiguous(), (n, m)))

    _FUSED_DTYPES = (torch.int8, torch.int16, torch.int32, torch.float32)

    def mul_fused(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None, residual_coeff: float = 1.0):
        """symmetric_quantize -> A @ x_q -> symmetric_dequantize (+ residual_coeff * residual) as the conv layers
        do around `mul` (models/pyg_gcn_conv.py:130-137, pyg_gin_conv.py:80-101), with the elementwise passes fused
        into the kernels: one absmax + one quantise kernel produce x_q, the SpMM's row store de-quantises and adds
        the residual.  Bit-identical to the unfused torch expression.  Returns None when the fusion does not apply
        (host operand, several sparse parts, 64-bit dtypes) - the caller then composes it from `mul`."""
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and len(self.parts) == 1 and
                self.dtype in self._FUSED_DTYPES and self.sp_info_ptr is not None):
            return None
        if self.format == "COO" and not pim_ops.plan_layout(self.sp_info_ptr)["coo_runs_as_csr"]:
            return None
        assert self.hidden_size == x.size(1)
        scale, x_q = pim_ops.quantize(x, self.dtype)
        return pim_ops.spmm_run_dense_ex(self.sp_info_ptr, x_q, scale=scale, residual=residual,
                                         residual_coeff=residual_coeff)

    def free(self):
        """Release the device plan (the reference leaks it: spmm_free_group is never called, spmv.py:37-41)."""
        if self.sp_info_ptr is not None:
            pim_ops.spmm_free_group(self.sp_info_ptr)
            self.sp_info_ptr = None

    def __copy__(self):
        # a shallow copy shares the index arrays but must not share (and later double-free) the plan handle
        clone = self.__class__.__new__(self.__class__)
        clone.__dict__.update(self.__dict__)
        clone.sp_info_ptr = None
        return clone

    def __del__(self):
        try:
            self.free()
        except Exception:      # interpreter shutdown, backend already released
            pass
