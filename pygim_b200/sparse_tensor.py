"""Minimal stand-in for ``torch_sparse.SparseTensor`` - the container PyGim's Python backend
receives as ``adj_t`` (backend_pim/spmm.py:15-55,127-136; spmm_test.py:54-68).

``torch_sparse`` is a third-party dependency of the reference (rusty1s/pytorch_sparse, unpinned in
Libs/install_libs.sh:13) that is not installable here.  Only the data-structure surface the
backend touches is provided: construction from COO or CSR, ``coo() / csr() / nnz() / sizes() /
size(d) / sparse_sizes() / device() / int() / to() / t()`` and column slicing ``adj[:, a:b]``.
There is deliberately NO ``matmul`` here: the CPU SpMM of ``--version=cpu`` is the baseline, it
lives with the oracle (oracle/spmm_oracle.c::oracle_spmm_csr_rowpar_*), not in the product.

A real ``torch_sparse.SparseTensor`` works with pygim_b200.backend_pim too (duck typing).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class SparseTensor:
    def __init__(self, row: Optional[torch.Tensor] = None, rowptr: Optional[torch.Tensor] = None,
                 col: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None,
                 sparse_sizes: Optional[Tuple[int, int]] = None, is_sorted: bool = False):
        assert col is not None and (row is not None or rowptr is not None)
        col = col.to(torch.int64)
        if row is None:
            rowptr = rowptr.to(torch.int64)
            counts = rowptr[1:] - rowptr[:-1]
            row = torch.repeat_interleave(torch.arange(rowptr.numel() - 1, device=col.device), counts)
            is_sorted_rows = True
        else:
            row = row.to(torch.int64)
            is_sorted_rows = is_sorted
        if sparse_sizes is None:
            n = int(row.max()) + 1 if row.numel() else 0
            m = int(col.max()) + 1 if col.numel() else 0
            sparse_sizes = (n, m)
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        if not (is_sorted and is_sorted_rows):
            # row-major, columns ascending inside a row (what torch_sparse's storage guarantees)
            key = row * max(self._sizes[1], 1) + col
            perm = torch.argsort(key, stable=True)
            if not torch.equal(perm, torch.arange(perm.numel(), device=perm.device)):
                row, col = row[perm], col[perm]
                if value is not None:
                    value = value[perm]
        self._row, self._col, self._value = row, col, value
        self._rowptr = rowptr if (rowptr is not None and is_sorted) else None

    # -- construction helpers
    @classmethod
    def from_edge_index(cls, edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor] = None,
                        sparse_sizes: Optional[Tuple[int, int]] = None, is_sorted: bool = False):
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr, sparse_sizes=sparse_sizes,
                   is_sorted=is_sorted)

    @classmethod
    def from_scipy(cls, mat, has_value: bool = True):
        mat = mat.tocsr()
        mat.sort_indices()
        value = torch.from_numpy(mat.data) if has_value else None
        return cls(rowptr=torch.from_numpy(mat.indptr.astype("int64")), col=torch.from_numpy(mat.indices.astype("int64")),
                   value=value, sparse_sizes=mat.shape, is_sorted=True)

    # -- accessors (torch_sparse API)
    def coo(self):
        return self._row, self._col, self._value

    def csr(self):
        if self._rowptr is None:
            counts = torch.bincount(self._row, minlength=self._sizes[0]) if self._row.numel() else \
                torch.zeros(self._sizes[0], dtype=torch.int64, device=self._col.device)
            rowptr = torch.zeros(self._sizes[0] + 1, dtype=torch.int64, device=self._col.device)
            torch.cumsum(counts, 0, out=rowptr[1:])
            self._rowptr = rowptr
        return self._rowptr, self._col, self._value

    def nnz(self) -> int:
        return int(self._col.numel())

    def sparse_sizes(self):
        return self._sizes

    def sizes(self):
        extra = tuple(self._value.shape[1:]) if self._value is not None else ()
        return list(self._sizes + extra)

    def size(self, dim: int) -> int:
        return self.sizes()[dim]

    def device(self):
        return self._col.device

    def dtype(self):
        return self._value.dtype if self._value is not None else torch.float

    def has_value(self) -> bool:
        return self._value is not None

    def _like(self, row, col, value, sizes, rowptr=None):
        out = SparseTensor.__new__(SparseTensor)
        out._row, out._col, out._value, out._sizes, out._rowptr = row, col, value, sizes, rowptr
        return out

    def int(self):
        v = None if self._value is None else self._value.to(torch.int)
        return self._like(self._row, self._col, v, self._sizes, self._rowptr)

    def to(self, device=None, dtype=None):
        mv = lambda t: None if t is None else t.to(device) if device is not None else t
        v = mv(self._value)
        if dtype is not None and v is not None:
            v = v.to(dtype)
        return self._like(mv(self._row), mv(self._col), v, self._sizes, mv(self._rowptr))

    def set_value(self, value: Optional[torch.Tensor], layout: Optional[str] = None):
        return self._like(self._row, self._col, value, self._sizes, self._rowptr)

    def t(self):
        return SparseTensor(row=self._col, col=self._row, value=self._value,
                            sparse_sizes=(self._sizes[1], self._sizes[0]))

    def __getitem__(self, index):
        # the backend only slices columns: self.raw[:, a:b] (spmm.py:131-133)
        if not (isinstance(index, tuple) and len(index) == 2):
            raise NotImplementedError("only adj[rows, cols] slicing is supported")
        rs, cs = index
        n, m = self._sizes
        r0, r1, rstep = rs.indices(n) if isinstance(rs, slice) else (None, None, None)
        c0, c1, cstep = cs.indices(m) if isinstance(cs, slice) else (None, None, None)
        if rstep != 1 or cstep != 1:
            raise NotImplementedError("only contiguous slices are supported")
        r1, c1 = max(r1, r0), max(c1, c0)
        keep = (self._col >= c0) & (self._col < c1) & (self._row >= r0) & (self._row < r1)
        value = None if self._value is None else self._value[keep]
        return self._like(self._row[keep] - r0, self._col[keep] - c0, value, (r1 - r0, c1 - c0))

    def __repr__(self):
        return "SparseTensor(sizes=%s, nnz=%d, has_value=%s)" % (self._sizes, self.nnz(), self.has_value())
