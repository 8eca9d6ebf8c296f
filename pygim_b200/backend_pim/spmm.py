"""`--version=spmm`: the default 2-D partitioned SpMM front-end.

Public names and behaviour follow backend_pim/spmm.py of the reference (SparseTensorCOO :15-136,
dense_split :9-13, TORCH_TYPES :141, prepare_pim_spmm :143-147, pim_spmm :149-150); the work is
done by libbackend_pim.so (sm_100a) through backend_pim.pim_ops.
"""
from __future__ import annotations

import torch

from . import pim_ops
from ._common import TORCH_TYPES, SparseTensorBase, split_widths  # noqa: F401  (TORCH_TYPES re-exported)


def dense_split(B, nparts, dim=1):
    """torch.chunk into contiguous column groups (the reference's per-rank B slices)."""
    if nparts == 1:
        return [B.contiguous()]
    return [piece.contiguous() for piece in torch.chunk(B, nparts, dim)]


class SparseTensorCOO(SparseTensorBase):
    def _plan(self, fmt, hidden_size, B_parts):
        self.format = fmt
        self.hidden_size = hidden_size
        self.dense_parts = B_parts
        self.max_B_parts_ncols = (hidden_size + B_parts - 1) / B_parts
        self.free()
        return split_widths(hidden_size, B_parts)

    def to_pim_group_csr(self, hidden_size, B_parts=4):
        h_size = self._plan("CSR", hidden_size, B_parts)
        if len(self.csr) != len(self.parts):
            self.build_csr()
        self.sp_info_ptr = pim_ops.spmm_csr_to_device_group(
            [p.crow_indices() for p in self.csr], [p.col_indices() for p in self.csr],
            [p.values() for p in self.csr], [p.size(0) for p in self.csr], [p.size(1) for p in self.csr],
            h_size, hidden_size)
        self._plan_created()

    def to_pim_group_coo(self, hidden_size, B_parts=4):
        h_size = self._plan("COO", hidden_size, B_parts)
        if len(self.coo) != len(self.parts):
            self.build_coo()
        self.row_indices = [p.row_indices() for p in self.coo]
        self.col_indices = [p.col_indices() for p in self.coo]
        self.values = [p.values() for p in self.coo]
        self.sp_info_ptr = pim_ops.spmm_coo_to_device_group(
            self.row_indices, self.col_indices, self.values, [p.size(0) for p in self.coo],
            [p.size(1) for p in self.coo], h_size, hidden_size)
        self._plan_created()

    def to_pim_group(self, hidden_size, B_parts=4):
        if self.format == "COO":
            self.to_pim_group_coo(hidden_size, B_parts)
        elif self.format == "CSR":
            self.to_pim_group_csr(hidden_size, B_parts)
        else:
            assert False

    def mul(self, B: torch.Tensor, out=None):
        assert self.hidden_size == B.size(1)
        if self.format not in ("CSR", "COO"):
            return None
        # torch.chunk and split_widths disagree when ds_parts does not divide well (e.g. 10 columns in
        # 4 parts: chunk gives 3,3,3,1 and 3 parts would be missing for 9 in 4); the reference then
        # trips its asserts.  The column tiles of the plan are authoritative here, so B is passed whole.
        return pim_ops.spmm_run_dense(self.sp_info_ptr, B, out=out)


def prepare_pim_spmm(adj_t, args):
    """backend_pim/spmm.py:143-147.  Two optional attributes of `args` (absent in the reference's drivers, so they
    run unchanged) switch on the prepare-time work SURVEY.md 8f lists: `reorder` ("cluster" | "tiles" | "degree"; also
    the PYGIM_REORDER environment variable) permutes A's rows so that rows sharing neighbours are adjacent ("tiles"
    additionally builds the hot/cold plan whose hot feature rows are gathered from shared memory), and `tune` (True, or ds_parts == 0) lets
    utils.autotuner pick the column tiling and the kernel options from the graph statistics."""
    import os
    method = getattr(args, "reorder", None) or os.environ.get("PYGIM_REORDER") or None
    perm, hot = None, None
    if method and method != "none":
        from ..reorder import hot_cold_plan, reorder_rows
        adj_t, perm, stats = reorder_rows(adj_t, method)
        row_bytes = args.hidden_size * torch.empty((), dtype=args.data_type).element_size()
        if method == "tiles" and args.sp_format == "CSR" and args.sp_parts == 1 and row_bytes >= 64 and row_bytes % 16 == 0:
            adj_t, hot = hot_cold_plan(adj_t, stats.get("group_of_row"))
    ds_parts, options = args.ds_parts, {}
    if getattr(args, "tune", False) or not ds_parts:
        from ..utils import autotuner
        choice = autotuner.tune_plan(adj_t, args.hidden_size, args.data_type, args.sp_format,
                                     cache_dir=getattr(args, "datadir", None), dataset=getattr(args, "dataset", None),
                                     reordered=perm is not None)
        ds_parts, options = choice["ds_parts"], choice["options"]
    A = SparseTensorCOO(adj_t, dtype=args.data_type, format=args.sp_format)
    A.row_perm = perm
    A.hot_plan = hot
    A.plan_options = options
    A.col_split(args.sp_parts)
    A.to_pim_group(args.hidden_size, ds_parts)
    return A


def pim_spmm(x, adj_t: SparseTensorCOO):
    return adj_t.mul(x)
