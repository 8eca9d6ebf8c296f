"""`--version=spmv`: SpMM executed as batches of `groups` single-column SpMVs (SparseP COO).

Follows backend_pim/spmv.py of the reference (SparseTensorCOO :21-109, prepare_pim_spmv :113-117,
pim_spmv :119-120): COO only, matrix dims padded to a multiple of 64/bits, B cut into
hidden/groups batches of `groups` columns, each batch handed to the SpMV op as `groups` vectors,
padded rows cropped, batches concatenated.  The batch of vectors runs as ONE `groups`-column launch
(the library plans a single groups-wide tile; the per-vector width list only describes the op's arguments).
"""
from __future__ import annotations

import torch

from . import pim_ops
from ._common import TORCH_TYPES, SparseTensorBase, split_widths  # noqa: F401


def dense_split(B, nparts, dim=1):
    if nparts == 1:
        return [B.contiguous()]
    return [piece.contiguous() for piece in torch.chunk(B, nparts, dim)]


def _bits(dtype: torch.dtype) -> int:
    # the reference uses torch.iinfo (spmv.py:46), which rejects FLT32/DBL64; element size is what is meant
    return torch.empty((), dtype=dtype).element_size() * 8


class SparseTensorCOO(SparseTensorBase):
    def __init__(self, coo, dtype=torch.int32, groups=32):
        super().__init__(coo.int(), dtype=dtype, format="")
        self.groups = groups

    def build_coo(self):
        super().build_coo(pad_to=max(64 // _bits(self.dtype), 1))

    def to_pim_group_coo(self, hidden_size, rank_pre_spmv=1):
        # `hidden_size` here is the number of vectors per call (= groups): one dense part per vector
        self.format = "COO"
        self.hidden_size = hidden_size
        self.dense_parts = hidden_size
        self.max_B_parts_ncols = 1.0
        if len(self.coo) != len(self.parts):
            self.build_coo()
        self.row_indices = [p.row_indices() for p in self.coo]
        self.col_indices = [p.col_indices() for p in self.coo]
        self.values = [p.values() for p in self.coo]
        self.free()
        self.sp_info_ptr = pim_ops.spmv_coo_to_device_group(
            self.row_indices, self.col_indices, self.values, [p.size(0) for p in self.coo],
            [p.size(1) for p in self.coo], split_widths(hidden_size, hidden_size), hidden_size, rank_pre_spmv)
        self._plan_created()

    def mul_single(self, B: torch.Tensor):
        assert self.hidden_size == B.size(1)
        pad = self.coo[0].size(1) - B.size(0)
        if pad > 0:                                   # the padded columns of A multiply nothing
            B = torch.nn.functional.pad(B, (0, 0, 0, pad))
        res = pim_ops.spmm_run_dense(self.sp_info_ptr, B)
        return res[:self.raw.size(0), ...]

    def mul(self, B: torch.Tensor):
        batches = dense_split(B, B.size(1) // self.groups)
        return torch.cat([self.mul_single(b) for b in batches], dim=1)

    def col_split(self, nparts=4):
        assert False   # spmv.py:107-108


def prepare_pim_spmv(adj_t, args):
    assert args.sp_format == "COO"
    A = SparseTensorCOO(adj_t, dtype=args.data_type, groups=args.ds_parts)
    A.to_pim_group_coo(args.ds_parts, args.sp_parts)
    return A


def pim_spmv(x, adj_t: SparseTensorCOO):
    return adj_t.mul(x)
