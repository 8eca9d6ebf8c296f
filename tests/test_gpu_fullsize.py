"""Whole-matrix parity at the FULL sizes BASELINE.json names (not the scaled-down graphs of the other GPU tests):
every element of C against the CPU oracle's row-parallel CSR SpMM - the reference's exact-equality criterion
(spmm_multigroup/mul_csr_multigroup.c:550-620) on its own input distribution (A = ones, X in {-8..3}: every partial
sum is an integer below 2^24, so FLT32 is exact in any summation order).

configs[0] arxiv-shape FLT32 CSR H=32 | configs[1] Reddit-shape FLT32 CSR H in {16, 128} |
configs[2] Reddit-shape INT8 / INT32 COO H=32 | configs[4] products-shape FLT32 CSR H=32."""
import numpy as np
import pytest
import torch

from helpers import make_args

pytestmark = pytest.mark.gpu


def _graph(shape, clustered=False):
    from pygim_b200 import graphgen
    from pygim_b200.sparse_tensor import SparseTensor
    n, nnz, max_deg = graphgen.SHAPES[shape]
    gen = graphgen.clustered_csr if clustered else graphgen.synthetic_csr
    rowptr, col = gen(n, nnz, max_deg, seed=0, device="cuda")
    adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(n, n), is_sorted=True)
    return adj, rowptr.cpu().numpy().astype(np.int32), col.cpu().numpy().astype(np.int32)


def _check(oracle, adj, rp, cl, dtype, fmt, hidden, reorder=None, ds_parts=1):
    from pygim_b200 import graphgen
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    n = adj.size(0)
    x = graphgen.reference_features(n, hidden, dtype, seed=hidden)
    args = make_args(dtype, fmt, hidden, ds_parts=ds_parts)
    if reorder:
        args.reorder = reorder
    A = prepare_pim_spmm(adj, args)
    got = A.mul(x.cuda())
    torch.cuda.synchronize()
    want = oracle.spmm_csr_rowpar(rp, cl, None, x.numpy(), nthreads=oracle.max_threads())
    bad = int((got.cpu().numpy() != want).sum())
    A.free()
    assert bad == 0, "%d of %d elements differ (%s %s H=%d reorder=%s)" % (bad, want.size, dtype, fmt, hidden, reorder)


@pytest.fixture(scope="module")
def reddit(gpu_backend):
    g = _graph("reddit")
    yield g
    del g
    torch.cuda.empty_cache()


@pytest.mark.parametrize("hidden,ds", [(16, 1), (128, 2), (128, 1)])
def test_reddit_shape_flt32_csr(gpu_backend, oracle, reddit, hidden, ds):
    _check(oracle, *reddit, torch.float32, "CSR", hidden, ds_parts=ds)


@pytest.mark.parametrize("dtype", [torch.int8, torch.int32])
def test_reddit_shape_quantised_coo(gpu_backend, oracle, reddit, dtype):
    _check(oracle, *reddit, dtype, "COO", 32)


def test_arxiv_shape_flt32_csr(gpu_backend, oracle):
    _check(oracle, *_graph("arxiv"), torch.float32, "CSR", 32)


def test_products_shape_flt32_csr(gpu_backend, oracle):
    _check(oracle, *_graph("products"), torch.float32, "CSR", 32)
    torch.cuda.empty_cache()


def test_clustered_reddit_shape_with_prepare_time_reordering(gpu_backend, oracle):
    """The block-model Reddit-shape graph (same N, nnz, degrees; hidden communities) through reorder="cluster":
    results in the original row order, every element equal."""
    g = _graph("reddit", clustered=True)
    _check(oracle, *g, torch.float32, "CSR", 32, reorder="cluster")
    _check(oracle, *g, torch.float32, "CSR", 32, reorder="tiles")
    _check(oracle, *g, torch.float32, "CSR", 128, reorder="tiles")
    del g
    torch.cuda.empty_cache()
