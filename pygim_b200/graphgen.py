"""Synthetic graphs of the dataset SHAPES PyGim's drivers load (no dataset download is possible).

spmm_test.py:40-71 loads Reddit / ogbn-arxiv / AmazonProducts / ogbn-products through PyG and
ToSparseTensor, which yields a value-less adjacency (=> ones) with row-major sorted, duplicate-free
int indices, and draws the features as `torch.randint(-2^6, 2^6, ...)` - in Python `^` is XOR, so
that is randint(-8, 4) (spmm_test.py:70).

The generator reproduces N, nnz (exactly) and the degree skew of each shape:
* out-degrees: log-normal, rescaled so that sum(deg) == nnz, capped at the dataset's max degree,
  rows in random order (no degree sorting);
* columns of a row of degree d: one uniformly random column inside each of d equal strata of
  [0, ncols) - strictly increasing, unique, and as scattered as uniform sampling without
  replacement (no artificial locality), but O(nnz) with no sort.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from .sparse_tensor import SparseTensor

# name -> (nodes, edges, max degree)
SHAPES = {
    "arxiv": (169_343, 1_166_243, 13_161),        # ogbn-arxiv
    "reddit": (232_965, 114_615_892, 21_657),     # Reddit
    "products": (2_449_029, 61_859_140, 17_481),  # ogbn-products
    "pubmed": (19_717, 88_648, 171),              # PubMed (spmm_test.py default dataset)
}


def _sigma_for(z: torch.Tensor, ratio: float) -> float:
    """log-normal sigma for which max(w) / mean(w) of THIS sample w = exp(sigma*z) equals `ratio`."""
    target = math.log(max(ratio, 1.0))
    zmax = float(z.max())

    def log_ratio(s: float) -> float:
        return s * zmax - float(torch.logsumexp(s * z, 0)) + math.log(z.numel())

    lo, hi = 0.0, 8.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if log_ratio(mid) < target:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def degree_sequence(n: int, nnz: int, max_deg: int, ncols: Optional[int] = None, seed: int = 0) -> torch.Tensor:
    """int64[n] degrees with sum == nnz, every row >= 1 when nnz >= n (the datasets have no isolated
    nodes to speak of) and max == min(max_deg, ncols) up to rounding; generated on the CPU so the same
    seed gives the same graph structure on every device."""
    cap = min(max_deg, ncols if ncols is not None else n)
    assert nnz <= n * cap, "shape is infeasible"
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, generator=g, dtype=torch.float64)
    base = 1 if nnz >= n else 0
    spare = nnz - base * n
    w = torch.exp(_sigma_for(z, (cap - base) / max(spare / n, 1e-9)) * z)
    deg = torch.full((n,), base, dtype=torch.int64)
    remaining, weights = spare, w.clone()
    for _ in range(8):   # rescale the not-yet-capped rows until the cap no longer binds
        free = deg < cap
        if remaining <= 0 or not bool(free.any()):
            break
        add = torch.floor(weights * (remaining / weights[free].sum())).to(torch.int64)
        add = torch.where(free, torch.minimum(add, cap - deg), torch.zeros_like(add))
        deg += add
        remaining = nnz - int(deg.sum())
        weights = torch.where(deg < cap, w, torch.zeros_like(w))
        if int(add.sum()) == 0:
            break
    if remaining > 0:    # hand the rounding remainder out one by one over a random order of open rows
        order = torch.randperm(n, generator=g)
        while remaining > 0:
            open_rows = order[deg[order] < cap]
            take = open_rows[:remaining]
            deg[take] += 1
            remaining -= int(take.numel())
    return deg


def synthetic_csr(n: int, nnz: int, max_deg: int, ncols: Optional[int] = None, seed: int = 0,
                  device: str = "cpu", rows: Optional[Tuple[int, int]] = None,
                  deg: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(rowptr int64[n+1], col int64[nnz]) of a random graph with the given shape.  `rows=(r0, r1)`
    generates only that row block (a shard: rowptr has r1-r0+1 entries starting at 0) of the SAME degree
    sequence; the column jitter of a shard is seeded by (seed, r0)."""
    m = n if ncols is None else ncols
    if deg is None:
        deg = degree_sequence(n, nnz, max_deg, m, seed)
    jitter_seed = seed + 1
    if rows is not None:
        deg = deg[rows[0]:rows[1]]
        n, nnz = int(deg.numel()), int(deg.sum())
        jitter_seed = seed + 1 + 7919 * rows[0]
    deg = deg.to(device)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=rowptr[1:])
    gen = torch.Generator(device=device).manual_seed(jitter_seed)
    col = torch.empty(nnz, dtype=torch.int64, device=device)
    # chunked over nnz so the temporaries stay small next to a 114.6 M-edge graph
    rows_per_chunk = max(1, int(n * (32_000_000 / max(nnz, 1))))
    for r0 in range(0, n, rows_per_chunk):
        r1 = min(n, r0 + rows_per_chunk)
        e0, e1 = int(rowptr[r0]), int(rowptr[r1])
        if e1 == e0:
            continue
        d = deg[r0:r1]
        row = torch.repeat_interleave(torch.arange(r1 - r0, device=device), d)
        j = torch.arange(e0, e1, device=device) - rowptr[r0:r1][row]
        dd = d[row]
        lo = (j * m) // dd
        hi = ((j + 1) * m) // dd
        u = torch.rand(e1 - e0, generator=gen, device=device, dtype=torch.float64)
        col[e0:e1] = lo + torch.clamp((u * (hi - lo).to(torch.float64)).to(torch.int64), max=(hi - lo - 1))
    return rowptr, col


def clustered_csr(n: int, nnz: int, max_deg: int, seed: int = 0, device: str = "cpu", community: int = 1024,
                  p_in: float = 0.7, hidden_order: bool = True, deg: Optional[torch.Tensor] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same N, nnz and degree sequence as synthetic_csr, but with COMMUNITY STRUCTURE (a stochastic block model):
    nodes form communities of `community` members and a row draws round(p_in * degree) of its neighbours inside
    its own community (as many as fit), the rest uniformly from all other nodes.  Real graphs of this kind (Reddit:
    posts of one subreddit; products: one category) look like this; synthetic_csr deliberately does not.

    With `hidden_order` the node ids are a random permutation of the community layout - like a real dataset the
    structure is there but not visible in the numbering, so a locality-aware pipeline has to FIND it
    (pygim_b200/reorder.py).  hidden_order=False numbers the nodes community by community (the best case a
    reordering can reach)."""
    if deg is None:
        deg = degree_sequence(n, nnz, max_deg, n, seed)
    g = torch.Generator().manual_seed(seed + 17)
    pi = torch.randperm(n, generator=g) if hidden_order else torch.arange(n)      # layout position -> node id
    pos_of = torch.empty(n, dtype=torch.int64)
    pos_of[pi] = torch.arange(n)                                                   # node id -> layout position
    deg = deg.to(device)
    pi_d, pos_d = pi.to(device), pos_of.to(device)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=rowptr[1:])
    gen = torch.Generator(device=device).manual_seed(seed + 1)
    col = torch.empty(int(rowptr[-1]), dtype=torch.int64, device=device)
    n_comm = max(1, (n + community - 1) // community)
    rows_per_chunk = max(1, int(n * (16_000_000 / max(nnz, 1))))
    for r0 in range(0, n, rows_per_chunk):
        r1 = min(n, r0 + rows_per_chunk)
        e0, e1 = int(rowptr[r0]), int(rowptr[r1])
        if e1 == e0:
            continue
        d = deg[r0:r1]
        k = pos_d[r0:r1] // community                                    # community of each row (layout space)
        c_lo = k * community
        c_sz = torch.clamp(c_lo + community, max=n) - c_lo
        d_in = torch.minimum(torch.round(d.to(torch.float64) * p_in).to(torch.int64), c_sz)
        d_in = torch.maximum(d_in, d - (n - c_sz))                       # the rest must fit outside
        row = torch.repeat_interleave(torch.arange(r1 - r0, device=device), d)
        j = torch.arange(e0, e1, device=device) - rowptr[r0:r1][row]
        inside = j < d_in[row]
        # stratified sample without replacement: d_in strata of the community, d - d_in strata of the complement
        cnt = torch.where(inside, d_in[row], (d - d_in)[row])
        span = torch.where(inside, c_sz[row], n - c_sz[row])
        jj = torch.where(inside, j, j - d_in[row])
        lo = (jj * span) // cnt
        hi = ((jj + 1) * span) // cnt
        u = torch.rand(e1 - e0, generator=gen, device=device, dtype=torch.float64)
        t = lo + torch.clamp((u * (hi - lo).to(torch.float64)).to(torch.int64), max=(hi - lo - 1))
        lay = torch.where(inside, c_lo[row] + t, torch.where(t < c_lo[row], t, t + c_sz[row]))   # layout position
        ids = pi_d[lay]
        # columns ascending inside a row
        key = row * n + ids
        col[e0:e1] = torch.sort(key).values - row * n
    del n_comm
    return rowptr, col


def synthetic_adj(shape: str = "arxiv", seed: int = 0, device: str = "cpu", scale: float = 1.0) -> SparseTensor:
    """A value-less SparseTensor (what ToSparseTensor produces) of a named dataset shape.  `scale` < 1
    shrinks nodes and edges proportionally (tests)."""
    n, nnz, max_deg = SHAPES[shape]
    if scale != 1.0:
        n = max(8, int(n * scale))
        nnz = max(n, int(nnz * scale))
        max_deg = max(4, min(int(max_deg * math.sqrt(scale)) + 1, n))
        nnz = min(nnz, n * max_deg)
    rowptr, col = synthetic_csr(n, nnz, max_deg, seed=seed, device=device)
    return SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(n, n), is_sorted=True)


def reference_features(n: int, hidden: int, dtype: torch.dtype, seed: int = 0, device: str = "cpu") -> torch.Tensor:
    """data.x of spmm_test.py:70: integers in [-8, 3] in the requested dtype."""
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randint(-8, 4, (n, hidden), generator=g, device=device, dtype=torch.int32).to(dtype)
