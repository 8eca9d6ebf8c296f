"""Shared helpers of the test-suite: small random sparse matrices in the layouts PyGim builds."""
from __future__ import annotations

import types

import numpy as np
import torch

from pygim_b200.sparse_tensor import SparseTensor

NP_DTYPES = {torch.int8: np.int8, torch.int16: np.int16, torch.int32: np.int32, torch.int64: np.int64,
             torch.float32: np.float32, torch.float64: np.float64}
ALL_DTYPES = list(NP_DTYPES)


def random_adj(n, m, density, seed=0, value_dtype=None, empty_rows=(), long_row=None, value_range=(-5, 6),
               real_valued=False):
    """SparseTensor with sorted unique indices.  value_dtype None => value-less (implicit ones).  Float
    values are integer-valued unless real_valued (integer-valued floats make f32/f64 sums exact, so the
    comparison with the oracle can be bit-exact regardless of summation order)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((n, m)) < density
    for r in empty_rows:
        mask[r, :] = False
    if long_row is not None:
        mask[long_row, :] = rng.random(m) < 0.9
    row, col = np.nonzero(mask)
    value = None
    if value_dtype is not None:
        if value_dtype.is_floating_point and real_valued:
            value = torch.from_numpy(rng.standard_normal(row.shape[0])).to(value_dtype)
        else:
            value = torch.from_numpy(rng.integers(value_range[0], value_range[1], row.shape[0])).to(value_dtype)
    return SparseTensor(row=torch.from_numpy(row), col=torch.from_numpy(col), value=value, sparse_sizes=(n, m),
                        is_sorted=True)


def make_args(dtype, fmt="CSR", hidden=32, sp_parts=1, ds_parts=1):
    return types.SimpleNamespace(data_type=dtype, sp_format=fmt, hidden_size=hidden, sp_parts=sp_parts,
                                 ds_parts=ds_parts)


def features(n, hidden, dtype, seed=0, integer_valued=True):
    g = torch.Generator().manual_seed(seed)
    if dtype.is_floating_point and not integer_valued:
        return torch.randn((n, hidden), generator=g, dtype=torch.float64).to(dtype)
    return torch.randint(-8, 4, (n, hidden), generator=g, dtype=torch.int32).to(dtype)


def oracle_spmm(O, adj, x, dtype):
    """C = A x through the oracle's COO definition (spmm_default/spmm_mul_coo.c:40-51); value-less A => ones."""
    row, col, value = adj.coo()
    nd = NP_DTYPES[dtype]
    val = np.ones(row.numel(), dtype=nd) if value is None else value.type(dtype).numpy()
    return torch.from_numpy(O.spmm_coo(row.numpy(), col.numpy(), val, x.numpy().astype(nd, copy=False), adj.size(0)))
