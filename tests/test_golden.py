"""The oracle restatement (and, with -m gpu, the CUDA path) against golden vectors produced by the
REFERENCE's own host oracles (tests/golden/make_golden.py).  Bit-exact for every dtype."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_host_oracles.npz"))
DTYPES = ["int8", "int16", "int32", "int64", "float32", "float64"]
CASES = ["small", "wide", "wrap"]


def _case(dt, case):
    k = "%s_%s_" % (dt, case)
    return {name[len(k):]: GOLD[name] for name in GOLD.files if name.startswith(k)}


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(oracle, dt, case):
    c = _case(dt, case)
    n = int(c["n"][0])
    assert np.array_equal(oracle.spmm_coo(c["row"], c["col"], c["val"], c["x"], n), c["y_coo"])
    assert np.array_equal(oracle.spmm_coo(c["row"], c["col"], c["val"], c["x"], n), c["y_spmv"])
    assert np.array_equal(oracle.spmm_csr(c["rowptr"], c["col"], c["val"], c["x"]), c["y_csr"])
    assert np.array_equal(oracle.spmm_csr_ones(c["rowptr"], c["col"], c["x"]), c["y_csr_ones"])
    assert np.array_equal(oracle.spmm_csr(c["rowptr"], c["col"], c["val"], c["xpad"], ncols=c["x"].shape[1]),
                          c["y_csr_pad"])
    # the row-parallel form (the CPU baseline / large-case oracle) is the same function of the inputs
    assert np.array_equal(oracle.spmm_csr_rowpar(c["rowptr"], c["col"], c["val"], c["x"]), c["y_csr"])
    # COO and CSR definitions agree (both wrap / both exact on these integer-valued inputs)
    assert np.array_equal(c["y_coo"], c["y_csr"])


@pytest.mark.parametrize("dt", DTYPES)
def test_oracle_group_matches_reference_golden(oracle, dt):
    k = "%s_group_" % dt
    c = {name[len(k):]: GOLD[name] for name in GOLD.files if name.startswith(k)}
    n, m, h = [int(v) for v in c["n"]]
    parts, cur = [], 0
    rowptr_parts = []
    for w in c["widths"]:
        sel = (c["col"] >= cur) & (c["col"] < cur + w)
        parts.append({"nrows": n, "ncols": int(w), "rowind": c["row"][sel], "colind": c["col"][sel] - cur,
                      "values": c["val"][sel]})
        rp = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(np.bincount(c["row"][sel], minlength=n), out=rp[1:])
        rowptr_parts.append(rp)
        cur += int(w)
    B_parts, col = [], 0
    for hj in c["hs"]:
        B_parts.append(np.ascontiguousarray(c["x"][:, col:col + hj]))
        col += int(hj)
    assert np.array_equal(oracle.spmm_group("COO", parts, B_parts), c["y"])
    csr_parts = [dict(p, rowptr=rp) for p, rp in zip(parts, rowptr_parts)]
    assert np.array_equal(oracle.spmm_group("CSR", csr_parts, B_parts), c["y"])


# ------------------------------------------------------------------ the CUDA path against the same fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_cuda_matches_reference_golden(gpu_backend, dt, case, fmt):
    from helpers import make_args
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    from pygim_b200.sparse_tensor import SparseTensor
    c = _case(dt, case)
    n, m, h = [int(v) for v in c["n"]]
    tdt = getattr(torch, dt)
    adj = SparseTensor(row=torch.from_numpy(c["row"].astype(np.int64)), col=torch.from_numpy(c["col"].astype(np.int64)),
                       value=torch.from_numpy(c["val"]), sparse_sizes=(n, m), is_sorted=True)
    A = prepare_pim_spmm(adj, make_args(tdt, fmt, h))
    out = A.mul(torch.from_numpy(c["x"]))
    assert np.array_equal(out.numpy(), c["y_coo"])
    A.free()


@pytest.mark.gpu
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_cuda_group_matches_reference_golden(gpu_backend, dt, fmt):
    """sp_parts=3 x ds_parts=2 through the public API reproduces the reference's group composition."""
    from helpers import make_args
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    from pygim_b200.sparse_tensor import SparseTensor
    k = "%s_group_" % dt
    c = {name[len(k):]: GOLD[name] for name in GOLD.files if name.startswith(k)}
    n, m, h = [int(v) for v in c["n"]]
    adj = SparseTensor(row=torch.from_numpy(c["row"].astype(np.int64)), col=torch.from_numpy(c["col"].astype(np.int64)),
                       value=torch.from_numpy(c["val"]), sparse_sizes=(n, m), is_sorted=True)
    A = prepare_pim_spmm(adj, make_args(getattr(torch, dt), fmt, h, sp_parts=3, ds_parts=2))
    assert [p.size(1) for p in A.parts] == list(c["widths"])       # col_split widths (spmm.py:129-133)
    out = A.mul(torch.from_numpy(c["x"]))
    assert np.array_equal(out.numpy(), c["y"])
    A.free()
