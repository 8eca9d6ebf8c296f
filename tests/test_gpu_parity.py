"""GPU parity: the sm_100a kernels behind the backend_pim API vs the CPU oracle, bit-exact for integers
and for integer-valued floats (the reference's own inputs), within the stated tolerance otherwise."""
import numpy as np
import pytest
import torch

from helpers import ALL_DTYPES, features, make_args, oracle_spmm, random_adj

pytestmark = pytest.mark.gpu


def _run(front, adj, args, x, **kw):
    if front == "spmm":
        from pygim_b200.backend_pim.spmm import prepare_pim_spmm
        A = prepare_pim_spmm(adj, args)
    elif front == "grande":
        from pygim_b200.backend_pim.grande import prepare_pim_spmm_grande
        A = prepare_pim_spmm_grande(adj, args, kw["dpus_per_rank"])
    else:
        from pygim_b200.backend_pim.spmv import prepare_pim_spmv
        A = prepare_pim_spmv(adj, args)
    if kw.get("seg_len"):
        from pygim_b200.backend_pim import pim_ops
        pim_ops.plan_set_option(A.sp_info_ptr, "seg_len", kw["seg_len"])
    out = A.mul(x)
    if out.is_cuda:
        torch.cuda.synchronize()
    A.free()
    return out


@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("fmt", ["CSR", "COO"])
@pytest.mark.parametrize("hidden", [1, 3, 16, 32, 33, 64, 128, 256])
def test_spmm_all_dtypes_host_operand(gpu_backend, oracle, dtype, fmt, hidden):
    adj = random_adj(301, 301, 0.05, seed=hidden, empty_rows=(0, 7, 300), long_row=5)
    x = features(301, hidden, dtype, seed=1)
    out = _run("spmm", adj, make_args(dtype, fmt, hidden), x, seg_len=64)     # the 270-nnz row is cut into segments
    assert out.dtype == dtype and out.device.type == "cpu"
    assert torch.equal(out, oracle_spmm(oracle, adj, x, dtype))


@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_spmm_device_operand_with_values(gpu_backend, oracle, dtype, fmt):
    adj = random_adj(500, 500, 0.04, seed=3, value_dtype=dtype, long_row=11)
    x = features(500, 64, dtype, seed=2)
    out = _run("spmm", adj.to("cuda"), make_args(dtype, fmt, 64), x.cuda())
    assert out.is_cuda
    assert torch.equal(out.cpu(), oracle_spmm(oracle, adj, x, dtype))


@pytest.mark.parametrize("fmt", ["CSR", "COO"])
@pytest.mark.parametrize("sp_parts,ds_parts", [(1, 1), (2, 1), (1, 2), (3, 4), (4, 3), (7, 5)])
def test_sp_ds_partitioning(gpu_backend, oracle, fmt, sp_parts, ds_parts):
    for dtype in (torch.int32, torch.float32, torch.int8):
        adj = random_adj(257, 257, 0.06, seed=5, value_dtype=dtype)
        x = features(257, 40, dtype, seed=4)
        out = _run("spmm", adj, make_args(dtype, fmt, 40, sp_parts, ds_parts), x)
        assert torch.equal(out, oracle_spmm(oracle, adj, x, dtype)), (dtype, sp_parts, ds_parts)


def test_int8_overflow_wraps(gpu_backend, oracle):
    adj = random_adj(64, 64, 0.9, seed=9, value_dtype=torch.int8, value_range=(-128, 128))
    x = torch.randint(-128, 128, (64, 32), dtype=torch.int32).to(torch.int8)
    for fmt in ("CSR", "COO"):
        out = _run("spmm", adj, make_args(torch.int8, fmt, 32), x)
        assert torch.equal(out, oracle_spmm(oracle, adj, x, torch.int8))


@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_long_rows_are_segmented(gpu_backend, oracle, fmt):
    # one row far above seg_len (512 minimum) so the segment + last-arriver merge path runs
    n = 3000
    adj = random_adj(40, n, 0.01, seed=1, long_row=3)
    for dtype in (torch.float32, torch.int16, torch.float64, torch.int64):
        x = features(n, 32, dtype, seed=6)
        out = _run("spmm", adj, make_args(dtype, fmt, 32), x)
        assert torch.equal(out, oracle_spmm(oracle, adj, x, dtype)), dtype


def test_float_tolerance_real_valued(gpu_backend, oracle):
    """FLT32 with real-valued inputs: |gpu - exact| <= 1e-5 * sum|a x| + 1e-6 per element."""
    adj = random_adj(400, 400, 0.1, seed=2, value_dtype=torch.float32, long_row=8, real_valued=True)
    x = features(400, 64, torch.float32, seed=3, integer_valued=False)
    rowptr, col, val = adj.csr()
    exact, mag = oracle.spmm_csr_f32_exact(rowptr.numpy(), col.numpy(), val.numpy(), x.numpy())
    for fmt in ("CSR", "COO"):
        out = _run("spmm", adj, make_args(torch.float32, fmt, 64), x).double().numpy()
        assert np.all(np.abs(out - exact) <= 1e-5 * mag + 1e-6), fmt


def test_grande_and_spmv_frontends(gpu_backend, oracle):
    adj = random_adj(203, 203, 0.05, seed=12)
    for dtype in (torch.int32, torch.float32, torch.int8):
        x = features(203, 32, dtype, seed=7)
        want = oracle_spmm(oracle, adj, x, dtype)
        for sp in (1, 2):
            out = _run("grande", adj, make_args(dtype, "CSR", 32, sp, 1), x, dpus_per_rank=[5] * sp)
            assert torch.equal(out, want), ("grande", dtype, sp)
        out = _run("spmv", adj, make_args(dtype, "COO", 32, 1, 8), x)
        assert out.shape == want.shape and torch.equal(out, want), ("spmv", dtype)


def test_empty_matrix_and_empty_rows(gpu_backend, oracle):
    adj = random_adj(50, 50, 0.0, seed=0)
    x = features(50, 16, torch.float32)
    for fmt in ("CSR", "COO"):
        out = _run("spmm", adj, make_args(torch.float32, fmt, 16), x)
        assert torch.equal(out, torch.zeros(50, 16))


def test_sharded_single_rank_and_reddit_like(gpu_backend, oracle):
    """ShardedSpMM with world == 1 is the plain plan; checked on a skewed Reddit-like graph (long rows are
    segmented, the persistent kernel's ticket counter is reused across calls)."""
    import types
    from pygim_b200 import graphgen
    from pygim_b200.sharded import ShardedSpMM
    adj = graphgen.synthetic_adj("reddit", scale=0.01, seed=2)
    rowptr, col, _ = adj.csr()
    n = adj.size(0)
    for dtype, hidden in ((torch.float32, 32), (torch.float32, 128), (torch.int8, 64), (torch.float64, 16)):
        args = types.SimpleNamespace(data_type=dtype, sp_format="CSR", hidden_size=hidden, sp_parts=1, ds_parts=1)
        op = ShardedSpMM(adj.to("cuda"), args)
        x = graphgen.reference_features(n, hidden, dtype, seed=3)
        want = oracle.spmm_csr_rowpar(rowptr.numpy(), col.numpy(), None, x.numpy())
        for _ in range(3):                       # repeated launches on one plan
            got = op.mul(x.cuda())
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), want), (dtype, hidden)
        op.free()


@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_unit_value_fast_path_equals_general_kernel(gpu_backend, oracle, fmt):
    """A value-less adjacency (=> ones) takes the unit-value kernels; forcing the general kernels on the same
    plan gives the same bits, and explicit non-unit values never take the fast path."""
    from pygim_b200.backend_pim import pim_ops
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    for dtype in (torch.float32, torch.int8, torch.int64):
        adj = random_adj(300, 300, 0.08, seed=21, long_row=4)
        x = features(300, 48, dtype, seed=5)
        want = oracle_spmm(oracle, adj, x, dtype)
        A = prepare_pim_spmm(adj, make_args(dtype, fmt, 48))
        fast = A.mul(x)
        pim_ops.plan_set_option(A.sp_info_ptr, "unit_values", 0)
        general = A.mul(x)
        assert torch.equal(fast, want) and torch.equal(general, want), dtype
        A.free()
        adj2 = random_adj(300, 300, 0.08, seed=21, value_dtype=dtype, long_row=4)     # mixed values incl. ones
        A2 = prepare_pim_spmm(adj2, make_args(dtype, fmt, 48))
        assert torch.equal(A2.mul(x), oracle_spmm(oracle, adj2, x, dtype)), dtype
        A2.free()


@pytest.mark.parametrize("rows_per_ticket", [1, 5, 31])
def test_streamed_row_tickets(gpu_backend, oracle, rows_per_ticket):
    """short_rows = 2: the rows of a ticket are read as one contiguous nonzero stream (products-/citation-like
    graphs).  Covers empty rows at every position, rows longer than seg_len inside a ticket (handled by segments),
    accumulation over sparse parts, column tiles and the unit-value path."""
    from pygim_b200 import graphgen
    from pygim_b200.backend_pim import pim_ops
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = graphgen.synthetic_adj("products", scale=0.002, seed=4)           # ~4.9 K rows, mean degree ~25, long rows
    n = adj.size(0)
    rowptr, col, _ = adj.csr()
    for dtype, hidden, sp, ds in ((torch.float32, 32, 1, 1), (torch.float32, 64, 2, 2), (torch.int32, 16, 1, 1),
                                  (torch.float64, 32, 1, 1), (torch.float32, 128, 1, 1)):
        x = features(n, hidden, dtype, seed=3)
        want = torch.from_numpy(oracle.spmm_csr_rowpar(rowptr.numpy(), col.numpy(), None, x.numpy()))
        A = prepare_pim_spmm(adj.to("cuda"), make_args(dtype, "CSR", hidden, sp_parts=sp, ds_parts=ds))
        pim_ops.plan_set_option(A.sp_info_ptr, "short_rows", 2)
        pim_ops.plan_set_option(A.sp_info_ptr, "rows_per_ticket", rows_per_ticket)
        for seg_len in (-1, 64):
            pim_ops.plan_set_option(A.sp_info_ptr, "seg_len", seg_len)
            for unit in (-1, 0):
                pim_ops.plan_set_option(A.sp_info_ptr, "unit_values", unit)
                got = A.mul(x.cuda())
                torch.cuda.synchronize()
                assert torch.equal(got.cpu(), want), (dtype, hidden, sp, ds, seg_len, unit)
        A.free()
    # empty rows everywhere + explicit values
    adj2 = random_adj(500, 400, 0.01, seed=7, value_dtype=torch.float32, empty_rows=tuple(range(0, 500, 3)))
    x2 = features(400, 32, torch.float32, seed=1)
    A2 = prepare_pim_spmm(adj2, make_args(torch.float32, "CSR", 32))
    pim_ops.plan_set_option(A2.sp_info_ptr, "short_rows", 2)
    pim_ops.plan_set_option(A2.sp_info_ptr, "rows_per_ticket", rows_per_ticket)
    assert torch.equal(A2.mul(x2), oracle_spmm(oracle, adj2, x2, torch.float32))
    A2.free()


@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_duplicate_edges_are_summed(gpu_backend, oracle, fmt):
    """Duplicate (row, col) entries: the COO front-end coalesces them (spmm.py:40-42 `.coalesce()`), the CSR
    front-end keeps them as separate nonzeros - either way they add up, in the value dtype (int8 wraps)."""
    from pygim_b200.sparse_tensor import SparseTensor
    rng = np.random.default_rng(5)
    n, m, nnz = 150, 120, 4000
    row = torch.from_numpy(rng.integers(0, n, nnz))
    col = torch.from_numpy(rng.integers(0, m, nnz))          # ~20 % of the pairs repeat
    for dtype in (torch.int8, torch.float32, torch.int64):
        val = torch.from_numpy(rng.integers(-100, 100, nnz)).to(dtype)
        adj = SparseTensor(row=row, col=col, value=val, sparse_sizes=(n, m))
        x = features(m, 24, dtype, seed=2)
        out = _run("spmm", adj, make_args(dtype, fmt, 24), x)
        r, c, v = adj.coo()
        want = torch.from_numpy(oracle.spmm_coo(r.numpy(), c.numpy(), v.numpy(), x.numpy(), n))
        assert torch.equal(out, want), dtype
