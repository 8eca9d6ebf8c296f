"""Pins the oracle restatement to the reference's own host oracles compiled in place from
/root/reference (oracle/_ref).  Skipped where the compiled reference objects are absent."""
import numpy as np
import pytest

DTYPES = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle


def _rand(seed, n, m, h, dens, dt, full=False):
    rng = np.random.default_rng(seed)
    mask = rng.random((n, m)) < dens
    row, col = np.nonzero(mask)
    if np.issubdtype(dt, np.integer) and full:
        info = np.iinfo(dt)
        lo, hi = max(info.min, -2 ** 31), min(info.max, 2 ** 31 - 1)
        val = rng.integers(lo, hi, row.shape[0]).astype(dt)
        x = rng.integers(lo, hi, (m, h)).astype(dt)
    elif np.issubdtype(dt, np.integer):
        val = rng.integers(-9, 9, row.shape[0]).astype(dt)
        x = rng.integers(-8, 4, (m, h)).astype(dt)
    else:
        val = rng.standard_normal(row.shape[0]).astype(dt)     # real-valued: same order => same rounding
        x = rng.standard_normal((m, h)).astype(dt)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(row, minlength=n), out=rowptr[1:])
    return row.astype(np.int32), col.astype(np.int32), val, rowptr, x


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("seed", range(4))
def test_coo_csr_spmv_definitions(ref, dt, seed):
    n, m, h = 30 + 11 * seed, 25 + 7 * seed, [1, 4, 17, 32][seed]
    row, col, val, rowptr, x = _rand(seed, n, m, h, 0.15, dt, full=(seed == 3))
    # identical loop order and no FP contraction: bit-exact even for real-valued floats
    assert np.array_equal(ref.spmm_coo(row, col, val, x, n), ref.ref_spmm_host_coo(row, col, val, x, n, "default"))
    assert np.array_equal(ref.spmm_coo(row, col, val, x, n), ref.ref_spmm_host_coo(row, col, val, x, n, "spmv"))
    assert np.array_equal(ref.spmm_csr(rowptr, col, val, x), ref.ref_spmm_host_csr(rowptr, col, val, x, "grande"))
    assert np.array_equal(ref.spmm_csr_ones(rowptr, col, x), ref.ref_spmm_host_csr(rowptr, col, val, x, "default"))


@pytest.mark.parametrize("dt", DTYPES)
def test_add_2d(ref, dt):
    rng = np.random.default_rng(5)
    A = rng.integers(-50, 50, (9, 13)).astype(dt)
    B = rng.integers(-50, 50, (4, 6)).astype(dt)
    want = A.copy()
    ref.ref_add_2d(want, B, 3, 5)
    got = A.copy()
    import ctypes as C
    f = getattr(ref.lib(), "oracle_add_2d_" + ref.SUFFIX[np.dtype(dt)])
    f(got.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), C.c_int64(13), C.c_int64(6), C.c_int64(3),
      C.c_int64(5), C.c_int64(4), C.c_int64(6))
    assert np.array_equal(got, want)


def test_rowpar_matches_scipy_and_torch(oracle):
    """Independent cross-check of the restatement: scipy CSR @ dense and torch.sparse.mm."""
    import scipy.sparse as sp
    import torch
    for dt in (np.float64, np.int64, np.int32):
        row, col, val, rowptr, x = _rand(11, 120, 90, 24, 0.1, dt)
        want = sp.csr_matrix((val, col, rowptr), shape=(120, 90)) @ x
        assert np.array_equal(oracle.spmm_csr_rowpar(rowptr, col, val, x), want)
    for dt in (np.int8, np.int16, np.float32):
        row, col, val, rowptr, x = _rand(12, 80, 70, 16, 0.2, dt, full=np.issubdtype(dt, np.integer))
        if dt == np.float32:
            val, x = np.round(val * 4), np.round(x * 4)
        coo = torch.sparse_coo_tensor(torch.from_numpy(np.stack([row, col]).astype(np.int64)), torch.from_numpy(val),
                                      (80, 70))
        want = torch.sparse.mm(coo, torch.from_numpy(x)).numpy()      # wraps modulo 2^n for ints
        assert np.array_equal(oracle.spmm_coo(row, col, val, x, 80), want)


def test_row_partitioners_against_the_reference(ref):
    """pygim_partition_rows_even == the reference's partition_by_row_csr; pygim_partition_rows_by_nnz returns the
    bottleneck-optimal contiguous partition, so its heaviest part is never heavier than that of the reference's
    greedy partition_by_nnz_csr (support/partition.c:14-44, 51-99)."""
    import torch
    from pygim_b200.backend_pim import pim_ops
    rng = np.random.default_rng(3)
    total_ours = total_theirs = 0
    for trial in range(30):
        nrows = int(rng.integers(1, 400))
        nparts = int(rng.integers(1, 12))
        deg = (rng.pareto(1.2, nrows) * 5).astype(np.int64)
        rowptr = np.zeros(nrows + 1, dtype=np.int32)
        np.cumsum(deg, out=rowptr[1:])
        assert list(ref.ref_partition_rows(rowptr, nparts, "row")) == pim_ops.partition_rows_even(nrows, nparts)
        ours = pim_ops.partition_rows_by_nnz(torch.from_numpy(rowptr), nparts)
        theirs = ref.ref_partition_rows(rowptr, nparts, "nnz")
        assert ours[0] == 0 and ours[-1] == nrows and ours == sorted(ours)
        load = lambda sp: max(int(rowptr[sp[i + 1]] - rowptr[sp[i]]) for i in range(nparts))
        assert load(ours) <= load(theirs), (trial, ours, list(theirs))
        assert load(ours) >= max(int(deg.max()), -(-int(rowptr[-1]) // nparts))          # the lower bound
        total_ours += load(ours)
        total_theirs += load(theirs)
    assert total_ours <= total_theirs
