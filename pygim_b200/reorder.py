"""Prepare-time row reordering for L1 locality.

The reference clusters large graphs before handing them to the backend (`ClusterData`, i.e. METIS, at
spmm_test.py:57-65) because a DPU can only hold a 64 MiB slice; on B200 the reason to cluster is the memory
hierarchy: the SpMM gathers `s * nnz * H` bytes of feature rows, on a graph in arbitrary node order every one of
them crosses L2 -> SM, and the only way below that is to let the warps that share an SM's L1 work
on rows that share neighbours.  The CSR kernel schedules supertickets of consecutive rows SM-affinely
(csrc/spmm_csr.cuh); this module supplies the row order.

Method - shared-neighbour clustering with sampled pivots, computed BY the aggregation kernel itself:

1. sample P pivot rows; D[c, p] = 1 if column c is a neighbour of pivot p (a dense [ncols x P] indicator);
2. O = A @ D through the backend: O[r, p] = |N(r) & N(p)|, the number of neighbours row r shares with pivot p;
3. every row joins the pivot it shares most neighbours with (rows that share none form a last group);
4. rows are ordered by (group, original index); A's rows are permuted accordingly.

Only ROWS are permuted - columns, and therefore the dense operand, are untouched.  The plan is told the
permutation (`pygim_plan_set_row_map`) and the kernels scatter on store, so callers see results in the original
row order and every row's sum is formed in the same nonzero order as before: results are bit-identical.
"""
from __future__ import annotations

import types
from typing import Optional, Tuple

import torch

from .sparse_tensor import SparseTensor


def permute_rows(adj: SparseTensor, perm: torch.Tensor) -> SparseTensor:
    """Row r of the result is row perm[r] of `adj` (columns and values untouched)."""
    rowptr, col, value = adj.csr()
    perm = perm.to(rowptr.device, torch.int64)
    deg = (rowptr[1:] - rowptr[:-1])[perm]
    new_ptr = torch.zeros_like(rowptr)
    torch.cumsum(deg, 0, out=new_ptr[1:])
    nnz = int(col.numel())
    new_col = torch.empty_like(col)
    new_val = None if value is None else torch.empty_like(value)
    # chunked gather: source index of new nonzero e = rowptr[perm[row(e)]] + (e - new_ptr[row(e)])
    n = perm.numel()
    rows_per_chunk = max(1, int(n * (32_000_000 / max(nnz, 1))))
    for r0 in range(0, n, rows_per_chunk):
        r1 = min(n, r0 + rows_per_chunk)
        e0, e1 = int(new_ptr[r0]), int(new_ptr[r1])
        if e1 == e0:
            continue
        shift = rowptr[:-1][perm[r0:r1]] - new_ptr[r0:r1]
        src = torch.arange(e0, e1, device=col.device) + torch.repeat_interleave(shift, deg[r0:r1])
        new_col[e0:e1] = col[src]
        if new_val is not None:
            new_val[e0:e1] = value[src]
    return SparseTensor(rowptr=new_ptr, col=new_col, value=new_val, sparse_sizes=adj.sparse_sizes(), is_sorted=True)


def _default_pivots(n_rows: int) -> int:
    return int(min(4096, max(32, (n_rows // 256 + 31) // 32 * 32)))


def cluster_rows(adj: SparseTensor, n_pivots: Optional[int] = None, seed: int = 0, batch: int = 256,
                 min_shared: int = 2) -> Tuple[torch.Tensor, dict]:
    """Row order (int64 permutation: new position -> original row) that puts rows sharing neighbours next to each
    other, plus statistics.  Needs the backend initialised and `adj` on the GPU (the overlap counts are ONE SpMM per
    `batch` pivots through libbackend_pim.so)."""
    from .backend_pim.spmm import SparseTensorCOO
    rowptr, col, _ = adj.csr()
    dev = col.device
    if dev.type != "cuda":
        raise ValueError("cluster_rows needs the adjacency on the GPU")
    n, m = adj.size(0), adj.size(1)
    P = int(n_pivots or _default_pivots(n))
    P = max(batch, (P + batch - 1) // batch * batch) if P > batch else P
    batch = min(batch, P)
    g = torch.Generator(device="cpu").manual_seed(seed)
    deg = rowptr[1:] - rowptr[:-1]
    # pivots: random rows that have neighbours
    cand = torch.nonzero(deg > 0).flatten()
    if cand.numel() == 0:
        return torch.arange(n, device=dev), {"pivots": 0, "groups": 1, "assigned": 0.0}
    pick = cand[torch.randperm(cand.numel(), generator=g)[:P].to(dev)]
    P = int(pick.numel())
    # value-less plan over the same index arrays: counts are exact in float32 (< 2^24 neighbours per row)
    pattern = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(n, m), is_sorted=True)
    best = torch.full((n,), -1.0, device=dev)
    arg = torch.full((n,), P, dtype=torch.int64, device=dev)
    for b0 in range(0, P, batch):
        b1 = min(P, b0 + batch)
        width = batch if (b1 - b0) == batch else ((b1 - b0 + 3) // 4 * 4)
        D = torch.zeros((m, width), dtype=torch.float32, device=dev)
        piv = pick[b0:b1]
        cnt = deg[piv]
        src = torch.repeat_interleave(rowptr[:-1][piv], cnt) + \
            (torch.arange(int(cnt.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
        D[col[src], torch.repeat_interleave(torch.arange(b1 - b0, device=dev), cnt)] = 1.0
        A = SparseTensorCOO(pattern, dtype=torch.float32, format="CSR")
        A.to_pim_group(width, 1)
        O = A.mul(D)
        A.free()
        val, idx = O[:, : b1 - b0].max(dim=1)
        better = val > best
        best = torch.where(better, val, best)
        arg = torch.where(better, idx + b0, arg)
        del D, O
    # a pivot row shares all of its own neighbours with itself; rows sharing fewer than min_shared form the tail
    arg = torch.where(best >= float(min_shared), arg, torch.full_like(arg, P))
    perm = torch.argsort(arg, stable=True)
    groups = int(torch.unique(arg).numel())
    stats = {"pivots": P, "groups": groups, "assigned": float((arg < P).float().mean()),
             "mean_shared": float(best.clamp(min=0).mean()), "mean_degree": float(deg.float().mean())}
    return perm, stats


def reorder_rows(adj: SparseTensor, method: str = "cluster", **kw) -> Tuple[SparseTensor, torch.Tensor, dict]:
    """(row-permuted adjacency, perm, stats).  perm[r] = original row of new row r - exactly the row map the plan
    needs.  Methods: "cluster" (shared-neighbour pivots), "degree" (rows by descending degree - keeps rows of
    similar length together, which evens out the work items; no locality claim)."""
    if method == "cluster":
        perm, stats = cluster_rows(adj, **kw)
    elif method == "degree":
        rowptr = adj.csr()[0]
        perm = torch.argsort(rowptr[1:] - rowptr[:-1], descending=True, stable=True)
        stats = {"groups": 1}
    else:
        raise ValueError("unknown reorder method %r" % (method,))
    stats["method"] = method
    return permute_rows(adj, perm), perm, stats
