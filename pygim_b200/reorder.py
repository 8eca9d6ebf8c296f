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
             "mean_shared": float(best.clamp(min=0).mean()), "mean_degree": float(deg.float().mean()),
             "group_of_row": arg[perm]}           # group of every row of the NEW order
    return perm, stats


def auto_seg_len(nnz: int, row_bytes: int = 0, sm_count: int = 148, nrows: int = 0) -> int:
    """The plan's automatic segment length (csrc/backend_pim.cu::auto_seg_len), needed here because rows that will be
    cut into segments must stay all-cold."""
    short = nrows > 0 and nnz < 96 * nrows
    s = nnz // max(1, sm_count * (64 * 8 if short else 16 * 6))
    if not short and 0 < row_bytes <= 128:
        s *= 2
    p = 512
    while p < s and p < 4096:
        p *= 2
    return p


def hot_cold_plan(adj: SparseTensor, group_of_row: Optional[torch.Tensor] = None, hot_k: int = 1280,
                  seg_len: Optional[int] = None, super_nnz: int = 65536) -> Tuple[SparseTensor, dict]:
    """Hot/cold form of a (row-clustered) adjacency for csrc/spmm_csr_hc.cuh.

    The rows are cut into supertickets of about `super_nnz` nonzeros (never across the groups in `group_of_row`,
    small neighbouring groups are merged); per superticket the `hot_k` most referenced columns become its shared-memory
    tile.  Returns the re-encoded adjacency - every row's hot nonzeros first, holding TILE SLOTS instead of column
    ids, values moved along - and {"super_rows", "hot_cols", "hot_cnt", "seg_len", "coverage"} for
    pygim_plan_set_hot_tiles.  Rows longer than `seg_len` (cut into segments by the plan) stay all-cold."""
    rowptr, col, value = adj.csr()
    dev = col.device
    n, m = adj.size(0), adj.size(1)
    nnz = int(col.numel())
    seg_len = int(seg_len or auto_seg_len(nnz, nrows=n))
    deg = rowptr[1:] - rowptr[:-1]
    short = deg <= seg_len
    w = torch.where(short, deg, torch.zeros_like(deg))
    cum = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(w, 0, out=cum[1:])
    cum_h = cum.cpu()
    # ---- superticket boundaries (host loop over the groups: a few thousand iterations)
    if group_of_row is not None and n > 0:
        g = group_of_row.to(dev)
        change = torch.nonzero(g[1:] != g[:-1]).flatten() + 1
        bounds = [0] + change.cpu().tolist() + [n]
    else:
        bounds = [0, n]
    cuts = [0]
    pending = 0                      # start row of groups merged so far
    for gi in range(len(bounds) - 1):
        ge = bounds[gi + 1]
        load = int(cum_h[ge] - cum_h[pending])
        if load < super_nnz // 2 and ge < n:
            continue                  # too little work for a tile of its own: merge with the next group
        pieces = max(1, int(round(load / super_nnz)))
        base = int(cum_h[pending])
        targets = torch.tensor([base + (load * j) // pieces for j in range(1, pieces)], dtype=torch.int64)
        inner = torch.searchsorted(cum_h, targets).tolist() if pieces > 1 else []
        for r in inner:
            r = min(max(int(r), cuts[-1] + 1), ge - 1)
            if r > cuts[-1]:
                cuts.append(r)
        if ge > cuts[-1]:
            cuts.append(ge)
        pending = ge
    if cuts[-1] != n:
        cuts.append(n)
    super_rows = torch.tensor(cuts, dtype=torch.int64, device=dev)
    S = super_rows.numel() - 1
    st_of_row = torch.repeat_interleave(torch.arange(S, device=dev), super_rows[1:] - super_rows[:-1])
    # ---- hot columns: the hot_k most referenced columns of each superticket (short rows only)
    e_row = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    e_short = short[e_row]
    key = st_of_row[e_row] * m + col
    uniq, counts = torch.unique(key[e_short], return_counts=True)       # sorted
    ust = uniq // m
    o1 = torch.argsort(counts, descending=True, stable=True)
    order = o1[torch.argsort(ust[o1], stable=True)]                     # by superticket, most referenced first
    ust_o = ust[order]
    seg_start = torch.searchsorted(ust_o, torch.arange(S, device=dev))
    rank = torch.arange(order.numel(), device=dev) - seg_start[ust_o]
    hot = (rank < hot_k) & (counts[order] >= 2)                          # a column used once gains nothing
    hot_cols = torch.full((S, hot_k), -1, dtype=torch.int32, device=dev)
    hot_cols[ust_o[hot], rank[hot]] = (uniq[order][hot] % m).to(torch.int32)
    hk, hk_perm = torch.sort(uniq[order][hot])
    hslot = rank[hot][hk_perm]
    del uniq, counts, ust, o1, order, ust_o, rank
    if hk.numel():
        pos = torch.searchsorted(hk, key).clamp_(max=hk.numel() - 1)
        is_hot = e_short & (hk[pos] == key)
        slot = hslot[pos]
    else:
        is_hot = torch.zeros_like(e_short)
        slot = torch.zeros_like(col)
    del key, e_short
    # ---- every row hot-first
    order2 = torch.argsort(e_row * 2 + (~is_hot).to(torch.int64), stable=True)
    new_col = torch.where(is_hot, slot, col)[order2]
    new_val = None if value is None else value[order2]
    hot_cnt = torch.bincount(e_row[is_hot], minlength=n).to(torch.int32)
    coverage = float(is_hot.float().mean()) if nnz else 0.0
    out = SparseTensor(rowptr=rowptr, col=new_col, value=new_val, sparse_sizes=(n, m), is_sorted=True)
    return out, {"super_rows": super_rows.to(torch.int32), "hot_cols": hot_cols, "hot_cnt": hot_cnt, "seg_len": seg_len,
                 "coverage": coverage, "supertickets": S, "hot_k": hot_k}


def reorder_rows(adj: SparseTensor, method: str = "cluster", **kw) -> Tuple[SparseTensor, torch.Tensor, dict]:
    """(row-permuted adjacency, perm, stats).  perm[r] = original row of new row r - exactly the row map the plan
    needs.  Methods: "cluster" (shared-neighbour pivots), "tiles" (the same order; the caller then builds the
    hot/cold plan with hot_cold_plan(adj, stats["group_of_row"])), "degree" (rows by descending degree - keeps rows of
    similar length together, which evens out the work items; no locality claim)."""
    if method in ("cluster", "tiles"):
        perm, stats = cluster_rows(adj, **kw)
    elif method == "degree":
        rowptr = adj.csr()[0]
        perm = torch.argsort(rowptr[1:] - rowptr[:-1], descending=True, stable=True)
        stats = {"groups": 1}
    else:
        raise ValueError("unknown reorder method %r" % (method,))
    stats["method"] = method
    return permute_rows(adj, perm), perm, stats
