"""`--version=grande`: every unit holds the whole sparse tile and a slice of the feature columns.

Follows backend_pim/grande.py of the reference (SparseTensorCOO :25-121, dense_split :12-23,
TYPES_MUL :11, prepare_pim_spmm_grande :124-128, pim_spmm_grande :130-131).  CSR only.  On the GPU a
"unit's column slice" is just a column tile of one launch, so `mul` hands B over whole; the per-unit
width list is still computed and kept (`dense_ncols`) because it is part of the public surface, and
`dense_split` still produces the reference's padded slices for callers that drive
`pim_ops.spmm_csr_run_group` themselves.

Intended semantics C = sum_i A[:, cols_i] * B[cols_i, :] are implemented; the reference's merge bug
for sp_parts > 1 (memadd / memadd_2D, spmm_grande/spmm.h:102-106, spmm_mul_csr.c:65-74) is not.
"""
from __future__ import annotations

import torch

from . import pim_ops
from ._common import TORCH_TYPES, SparseTensorBase  # noqa: F401

# elements per 8 bytes: the DPU DMA granularity the slices are padded to
TYPES_MUL = {torch.int64: 1, torch.int32: 2, torch.int16: 4, torch.int8: 8, torch.float32: 2, torch.float64: 1}


def dense_split(B, ncols, dim=1):
    """Cut B into len(ncols) column slices, each `pad` columns wide where pad = ncols[0] rounded up
    to 8 bytes; slice k starts at sum(ncols[:k]) so a padded slice overlaps its right neighbour and
    only the last one needs real zero padding."""
    unit = TYPES_MUL[B.dtype]
    pad = (ncols[0] + unit - 1) // unit * unit
    if ncols[-1] % pad != 0:
        B = torch.nn.functional.pad(B, (0, pad - ncols[-1] % pad))
    if len(ncols) == 1:
        return [B.contiguous()]
    out, start = [], 0
    for w in ncols:
        out.append(B[:, start:start + pad].contiguous())
        start += w
    return out


class SparseTensorCOO(SparseTensorBase):
    def __init__(self, coo, dtype=torch.int32, dpus_per_rank=[], format=""):
        super().__init__(coo.int(), dtype=dtype, format="")
        self.dpus_per_rank = dpus_per_rank

    def to_pim_group_csr(self, hidden_size, B_parts=4):
        self.format = "CSR"
        self.hidden_size = hidden_size
        if len(self.csr) != len(self.parts):
            self.build_csr()
        widths = []
        for i in range(len(self.csr)):
            units = self.dpus_per_rank[i]
            per = [hidden_size // units] * units
            for k in range(hidden_size - per[0] * units):
                per[k] += 1
            widths.append(torch.tensor(per, dtype=torch.int32))
        self.dense_ncols = widths
        self.free()
        self.sp_info_ptr = pim_ops.spmm_csr_to_device_group(
            [p.crow_indices() for p in self.csr], [p.col_indices() for p in self.csr],
            [p.values() for p in self.csr], [p.size(0) for p in self.csr], [p.size(1) for p in self.csr],
            self.dense_ncols, hidden_size)
        self._plan_created()

    def mul(self, B: torch.Tensor, out=None):
        assert self.hidden_size == B.size(1)
        assert len(self.dpus_per_rank) == len(self.csr)
        if self.format != "CSR":
            return None
        return pim_ops.spmm_run_dense(self.sp_info_ptr, B, out=out)


def prepare_pim_spmm_grande(adj_t, args, dpus_per_rank):
    A = SparseTensorCOO(adj_t, dtype=args.data_type, dpus_per_rank=dpus_per_rank, format=args.sp_format)
    A.col_split(args.sp_parts)
    A.to_pim_group_csr(args.hidden_size)
    return A


def pim_spmm_grande(x, adj_t: SparseTensorCOO):
    return adj_t.mul(x)
