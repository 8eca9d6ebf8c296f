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


class ColumnShardedSpMM:
    """C = A @ B with the FEATURE COLUMNS dealt over the ranks (the cross-GPU form of `dense_split`, spmm.py:9-13):
    every rank holds all of A and computes all rows of its column block C[:, c0:c1] from B[:, c0:c1].  A
    stand-alone SpMM needs no exchange at all in this form (outputs are concatenated, never summed); `gather=True`
    all-gathers the column blocks for callers that need the full C everywhere.  Row sharding (ShardedSpMM) is the
    default because it also divides A's stream; this form is for wide features on graphs whose A is small."""

    def __init__(self, adj: SparseTensor, args, group=None, make_local=None, align: int = 4):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.hidden_size = args.hidden_size
        self.nrows = adj.size(0)
        units = -(-self.hidden_size // align)                   # columns are dealt in blocks of `align`
        per, rest = divmod(units, self.world)
        bounds = [0]
        for r in range(self.world):
            bounds.append(min(self.hidden_size, bounds[-1] + (per + (1 if r < rest else 0)) * align))
        self.col_splits = bounds
        self.c0, self.c1 = bounds[self.rank], bounds[self.rank + 1]
        if make_local is None:
            from .backend_pim.spmm import prepare_pim_spmm
            make_local = prepare_pim_spmm
        local_args = types.SimpleNamespace(**vars(args))
        local_args.hidden_size = self.c1 - self.c0
        self.local = make_local(adj, local_args) if self.c1 > self.c0 else None

    def mul(self, B: torch.Tensor, gather: bool = True) -> torch.Tensor:
        assert B.size(1) == self.hidden_size
        w = self.c1 - self.c0
        mine = self.local.mul(B[:, self.c0:self.c1]) if w else B.new_empty((self.nrows, 0))
        if not gather or self.world == 1:
            return mine
        blocks = [B.new_empty((self.nrows, self.col_splits[r + 1] - self.col_splits[r])) for r in range(self.world)]
        if len({b.size(1) for b in blocks}) == 1 or dist.get_backend(self.group) == "nccl":
            dist.all_gather(blocks, mine.contiguous(), group=self.group)
        else:
            blocks[self.rank].copy_(mine)
            _gather_views(blocks, self.group)
        return torch.cat(blocks, dim=1)

    def free(self):
        if self.local is not None and hasattr(self.local, "free"):
            self.local.free()


def _gather_views(views: Sequence[torch.Tensor], group=None) -> None:
    ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
    for i, v in enumerate(views):
        if v.numel():
            dist.broadcast(v, src=ranks[i], group=group)


class ShardedSpMM:
    """C = A @ B with A row-sharded over the ranks of `group`.

    `adj` is the FULL adjacency (every rank passes the same one; only the local row range is planned)
    unless `splits` and `local_adj` are given.  `make_local` builds the per-rank operator from the local
    shard - by default the CUDA plan of backend_pim.spmm; tests on CPU inject a checker there.

    `chunks` > 1 cuts the local row range into that many nnz-balanced sub-blocks, each with its own plan:
    the all-gather of sub-block k (NCCL's stream) then overlaps the SpMM of sub-block k+1 (compute stream).

    `fused=True` (CUDA, CSR or sorted COO, sp_parts == 1) removes the collective altogether: the result lives in a
    symmetric-memory buffer and the SpMM kernel's epilogue stores every output row straight into every
    peer's copy over NVLink (one multimem.st through the NVSwitch when multicast is available, else one store
    per peer); the kernel itself signals completion to the peers (`sync="flags"`), so there is no barrier and no
    extra launch between two layers beyond a one-warp wait.  The transfer overlaps the gathers row by row.
    """

    def __init__(self, adj: Optional[SparseTensor], args, group=None, splits: Optional[Sequence[int]] = None,
                 local_adj: Optional[SparseTensor] = None,
                 make_local: Optional[Callable[[SparseTensor, object], object]] = None, chunks: int = 1,
                 fused: bool = False, use_multicast: bool = True, sync: str = "flags",
                 world: Optional[int] = None, rank: Optional[int] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is not None:        # e.g. one rank running the whole graph alone inside a multi-rank job
            self.world, self.rank = int(world), int(rank or 0)
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
        self.chunks = max(1, int(chunks)) if self.world > 1 else 1
        if self.chunks == 1:
            self.sub = [0, self.r1 - self.r0]
            self.locals = [make_local(local_adj, types.SimpleNamespace(**vars(args)))]
            self.all_sub = None
        else:
            self.sub = row_splits_by_nnz(local_adj.csr()[0], self.chunks)      # local row offsets of the sub-blocks
            self.locals = [make_local(shard_rows(local_adj, self.sub[k], self.sub[k + 1]),
                                      types.SimpleNamespace(**vars(args))) for k in range(self.chunks)]
            gathered = [None] * self.world
            dist.all_gather_object(gathered, self.sub, group=group)            # every rank's sub-block offsets
            self.all_sub = gathered
        self.local = self.locals[0]
        self.fused = bool(fused) and self.world > 1
        self.use_multicast = use_multicast
        assert sync in ("flags", "barrier")
        self.sync = sync         # how a fused call learns that every peer's rows have landed
        self.peer_mask = None    # optional uint8 [local rows]: bit p = peer p needs the row (halo exchange)
        self._symm = {}          # dtype -> [slots, next slot, flags, flag ptrs, epochs]
        if self.fused and self.chunks != 1:
            raise ValueError("fused=True replaces the chunked NCCL schedule; use chunks=1")

    # -- fused all-gather: symmetric result buffer + peer stores from the kernel epilogue
    def _symmetric_out(self, dtype: torch.dtype, device: torch.device):
        """Symmetric result buffers per dtype, used in rotation (two with a barrier per call, three with arrival
        flags), plus - for sync="flags" - one symmetric int32 flag vector [slots x world]."""
        if dtype not in self._symm:
            import torch.distributed._symmetric_memory as symm_mem
            group = self.group if self.group is not None else dist.group.WORLD
            n_slots = 3 if self.sync == "flags" else 2
            slots = []
            for _ in range(n_slots):
                buf = symm_mem.empty((self.nrows, self.hidden_size), dtype=dtype, device=device)
                hdl = symm_mem.rendezvous(buf, group.group_name)
                mc = int(hdl.multicast_ptr) if (self.use_multicast and hdl.has_multicast_support) else 0
                slots.append((buf, hdl, [int(p) for p in hdl.buffer_ptrs], mc))
            flags = symm_mem.empty((n_slots * self.world,), dtype=torch.int32, device=device)
            flags.zero_()
            fh = symm_mem.rendezvous(flags, group.group_name)
            torch.cuda.synchronize(device)
            self._symm[dtype] = [slots, 0, flags, [int(p) for p in fh.buffer_ptrs], [0] * n_slots]
            slots[0][1].barrier(channel=0)    # buffers and zeroed flags exist everywhere before the first remote store
        return self._symm[dtype]

    def _mul_fused(self, B: torch.Tensor) -> torch.Tensor:
        """SpMM whose epilogue stores this rank's rows into every peer's buffer.

        sync="flags" (default): NO collective and no barrier.  The kernel's last warp writes this call's epoch
        into every peer's flag slot once all of its rows are stored (release, system scope); each rank then
        enqueues a one-warp wait on its own flag vector.  Three result buffers rotate: call k writes buffer k%3,
        which peers can only start to overwrite (call k+3) after this rank's kernel k+2 has finished - so a result
        stays valid until the call after next has been ENQUEUED here.
        sync="barrier": two buffers and one symmetric-memory barrier per call (the round-1 scheme)."""
        from .backend_pim import pim_ops
        state = self._symmetric_out(B.dtype, B.device)
        slots, k = state[0], state[1]
        buf, hdl, ptrs, mc = slots[k]
        state[1] = (k + 1) % len(slots)
        if self.sync == "flags":
            state[4][k] += 1
            epoch = state[4][k]
            off = k * self.world * 4
            pim_ops.spmm_run_dense_peers(self.locals[0].sp_info_ptr, B, ptrs, mc, self.hidden_size, self.r0,
                                         peer_mask=self.peer_mask, flag_ptrs=[p + off for p in state[3]],
                                         my_rank=self.rank, epoch=epoch)
            pim_ops.wait_flags(state[2][k * self.world:(k + 1) * self.world], epoch)
        else:
            pim_ops.spmm_run_dense_peers(self.locals[0].sp_info_ptr, B, ptrs, mc, self.hidden_size, self.r0,
                                         peer_mask=self.peer_mask)
            hdl.barrier(channel=0)          # every rank's rows have landed everywhere
        return buf

    def _sub_block_views(self, out: torch.Tensor, k: int):
        """Rows of sub-block k of every rank, as views of the full result."""
        return [out[self.splits[i] + self.all_sub[i][k]: self.splits[i] + self.all_sub[i][k + 1]]
                for i in range(self.world)]

    def mul(self, B: torch.Tensor, out: Optional[torch.Tensor] = None, gather: bool = True) -> torch.Tensor:
        """Returns the full [N x H] result (on B's device) when `gather`, else a view of the local block."""
        assert B.size(1) == self.hidden_size
        if self.fused and gather:
            res = self._mul_fused(B)      # the symmetric buffer IS the result (valid until the next mul)
            if out is not None:
                out.copy_(res)
                return out
            return res
        if out is None:
            out = torch.empty((self.nrows, self.hidden_size), dtype=B.dtype, device=B.device)
        mine = out[self.r0:self.r1]
        if self.chunks == 1:
            self.locals[0].mul(B, out=mine)
            if gather and self.world > 1:
                all_gather_rows(out, self.row_counts, mine, self.group)
            return out if gather else mine
        pending = []
        for k in range(self.chunks):
            block = mine[self.sub[k]:self.sub[k + 1]]
            self.locals[k].mul(B, out=block)
            if gather:
                # async: NCCL waits for the kernel just enqueued, then runs beside the next sub-block's SpMM
                views = self._sub_block_views(out, k)
                if dist.get_backend(self.group) == "nccl":
                    pending.append(dist.all_gather(views, block, group=self.group, async_op=True))
                else:   # equal-size-only backends (gloo in the CPU tests): one broadcast per block
                    _gather_views(views, self.group)
        for w in pending:
            w.wait()
        return out if gather else mine

    def mul_gather_only(self, out: torch.Tensor) -> torch.Tensor:
        """The collective half of `mul` on its own (lets a caller time compute and all-gather apart)."""
        if self.world > 1:
            all_gather_rows(out, self.row_counts, out[self.r0:self.r1], self.group)
        return out

    def free(self):
        for op in self.locals:
            if hasattr(op, "free"):
                op.free()
