"""Row-sharded aggregation across the GPUs of one box (one process per GPU, torch.distributed).

PyGim spreads one SpMM over UPMEM ranks/DPUs with a 2-D partitioning orchestrated by one host thread
(backend_pim/spmm.py:105-136; support/partition.c); the DPU row blocks are gathered back by
dpu_push_xfer(FROM_DPU) and merged on the host (spmm_default/spmm_mul_csr.c:385-410,479-554).  The B200
mapping:

* the adjacency is cut into `world` contiguous row ranges of near-equal nnz
  (pygim_partition_rows_by_nnz, the GPU-level partition_by_nnz_csr, support/partition.c:51-99);
* the dense operand is replicated (the reference broadcasts the B slice to every DPU of a rank,
  spmm_mul_csr.c:352-367);
* every rank runs the single-GPU plan on its row range, writing straight into its slice of the full
  [N x H] result;
* ONE all-gather of the (unequal) row blocks over NCCL/NVLink leaves the full result on every rank,
  ready to be the next layer's dense operand.  There is no host merge and no CPU fallback.
"""
from __future__ import annotations

import types
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .sparse_tensor import SparseTensor


def row_splits_by_nnz(rowptr: torch.Tensor, world: int) -> List[int]:
    from .backend_pim import pim_ops
    return pim_ops.partition_rows_by_nnz(rowptr, world)


def shard_rows(adj: SparseTensor, r0: int, r1: int) -> SparseTensor:
    """Rows [r0, r1) of `adj` as a SparseTensor of shape (r1 - r0, ncols)."""
    rowptr, col, value = adj.csr()
    e0, e1 = int(rowptr[r0]), int(rowptr[r1])
    return SparseTensor(rowptr=(rowptr[r0:r1 + 1] - rowptr[r0]).clone(), col=col[e0:e1],
                        value=None if value is None else value[e0:e1], sparse_sizes=(r1 - r0, adj.size(1)),
                        is_sorted=True)


def all_gather_rows(out: torch.Tensor, row_counts: Sequence[int], mine: torch.Tensor, group=None) -> None:
    """All-gather unequal row blocks IN PLACE: block i of `out` (row_counts[i] rows) is rank i's `mine`.
    NCCL takes the uneven list directly (one grouped collective); backends that insist on equal sizes
    (gloo, used by the CPU tests) get one broadcast per block."""
    chunks = list(torch.split(out, list(row_counts), dim=0))
    if len(set(row_counts)) == 1 or dist.get_backend(group) == "nccl":
        dist.all_gather(chunks, mine, group=group)
        return
    ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
    for i, chunk in enumerate(chunks):
        if chunk.numel():
            dist.broadcast(chunk, src=ranks[i], group=group)


class ShardedSpMM:
    """C = A @ B with A row-sharded over the ranks of `group`.

    `adj` is the FULL adjacency (every rank passes the same one; only the local row range is planned)
    unless `splits` and `local_adj` are given.  `make_local` builds the per-rank operator from the local
    shard - by default the CUDA plan of backend_pim.spmm; tests on CPU inject a checker there.
    """

    def __init__(self, adj: Optional[SparseTensor], args, group=None, splits: Optional[Sequence[int]] = None,
                 local_adj: Optional[SparseTensor] = None,
                 make_local: Optional[Callable[[SparseTensor, object], object]] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if splits is None:
            rowptr = adj.csr()[0]
            splits = row_splits_by_nnz(rowptr, self.world) if self.world > 1 else [0, adj.size(0)]
        self.splits = [int(s) for s in splits]
        assert len(self.splits) == self.world + 1 and self.splits[0] == 0
        self.r0, self.r1 = self.splits[self.rank], self.splits[self.rank + 1]
        self.nrows = self.splits[-1]
        self.row_counts = [self.splits[i + 1] - self.splits[i] for i in range(self.world)]
        if local_adj is None:
            local_adj = shard_rows(adj, self.r0, self.r1)
        self.local_adj = local_adj
        self.hidden_size = args.hidden_size
        if make_local is None:
            from .backend_pim.spmm import prepare_pim_spmm
            make_local = prepare_pim_spmm
        local_args = types.SimpleNamespace(**vars(args))
        self.local = make_local(local_adj, local_args)

    def mul(self, B: torch.Tensor, out: Optional[torch.Tensor] = None, gather: bool = True) -> torch.Tensor:
        """Returns the full [N x H] result (on B's device) when `gather`, else a view of the local block."""
        assert B.size(1) == self.hidden_size
        if out is None:
            out = torch.empty((self.nrows, self.hidden_size), dtype=B.dtype, device=B.device)
        mine = out[self.r0:self.r1]
        self.local.mul(B, out=mine)
        if gather and self.world > 1:
            all_gather_rows(out, self.row_counts, mine, self.group)
        return out if gather else mine

    def mul_gather_only(self, out: torch.Tensor) -> torch.Tensor:
        """The collective half of `mul` on its own (lets a caller time compute and all-gather apart)."""
        if self.world > 1:
            all_gather_rows(out, self.row_counts, out[self.r0:self.r1], self.group)
        return out

    def free(self):
        if hasattr(self.local, "free"):
            self.local.free()
