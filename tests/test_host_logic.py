"""Host-side logic that needs no GPU: the SparseTensor stand-in, the front-ends' preprocessing
(col_split, CSR/COO building, coalescing, spmv padding, grande column dealing), the graph generator,
the autotuner's column tiling."""
import types

import pytest
import torch

from helpers import random_adj
from pygim_b200 import graphgen
from pygim_b200.backend_pim import _common, grande, spmm, spmv
from pygim_b200.sparse_tensor import SparseTensor
from pygim_b200.utils import autotuner


def test_sparse_tensor_sorts_and_slices():
    row = torch.tensor([2, 0, 1, 0, 2])
    col = torch.tensor([1, 3, 0, 1, 0])
    val = torch.tensor([5., 1., 3., 2., 4.])
    a = SparseTensor(row=row, col=col, value=val, sparse_sizes=(3, 4))
    r, c, v = a.coo()
    assert r.tolist() == [0, 0, 1, 2, 2] and c.tolist() == [1, 3, 0, 0, 1] and v.tolist() == [2., 1., 3., 4., 5.]
    assert a.csr()[0].tolist() == [0, 2, 3, 5]
    assert a.nnz() == 5 and a.sizes() == [3, 4] and a.size(1) == 4
    left, right = a[:, 0:2], a[:, 2:]
    assert left.sizes() == [3, 2] and right.sizes() == [3, 2]
    assert left.coo()[1].tolist() == [1, 0, 0, 1] and right.coo()[1].tolist() == [1]
    assert a.int().coo()[2].dtype == torch.int32
    assert a.t().sizes() == [4, 3]


def test_col_split_widths_follow_reference():
    adj = random_adj(10, 23, 0.3, seed=1)
    A = spmm.SparseTensorCOO(adj, dtype=torch.int32, format="CSR")
    parts = A.col_split(4)                         # ceil(23/4) = 6, 6, 6, 5 (spmm.py:129-133)
    assert [p.size(1) for p in parts] == [6, 6, 6, 5]
    assert sum(p.nnz() for p in parts) == adj.nnz()
    assert _common.split_widths(10, 4) == [3, 3, 3, 1]      # h_size list (spmm.py:60-72)
    assert _common.split_widths(32, 1) == [32]
    with pytest.raises(AssertionError):
        A.row_split(2)


def test_build_csr_and_coo_values_and_dtypes():
    adj = random_adj(12, 9, 0.4, seed=2, value_dtype=torch.float32, real_valued=True)
    A = spmm.SparseTensorCOO(adj, dtype=torch.int8, format="CSR")
    A.build_csr()
    csr = A.csr[0]
    assert csr.crow_indices().dtype == torch.int32 and csr.col_indices().dtype == torch.int32
    assert csr.values().dtype == torch.int8
    # present values are cast with .type(dtype): truncation toward zero (spmm.py:38-39)
    assert torch.equal(csr.values(), adj.coo()[2].type(torch.int8))
    B = spmm.SparseTensorCOO(random_adj(12, 9, 0.4, seed=2), dtype=torch.float64, format="COO")
    B.build_coo()
    assert torch.equal(B.coo[0].values(), torch.ones(B.raw.nnz(), dtype=torch.float64))   # None => ones


def test_coalesce_sums_duplicates_with_wraparound():
    row = torch.tensor([1, 0, 1, 1, 0])
    col = torch.tensor([2, 1, 2, 0, 1])
    val = torch.tensor([100, 7, 100, 3, -7], dtype=torch.int8)
    r, c, v = _common.coalesce(row, col, val, 3)
    assert r.tolist() == [0, 1, 1] and c.tolist() == [1, 0, 2]
    assert v.tolist() == [0, 3, -56]               # 100 + 100 wraps to -56 in int8, like torch's coalesce
    want = torch.sparse_coo_tensor(torch.stack([row, col]), val, (2, 3)).coalesce()
    assert torch.equal(v, want.values()) and torch.equal(torch.stack([r, c]), want.indices())


def test_spmv_pads_both_dims_and_grande_deals_columns():
    adj = random_adj(13, 13, 0.3, seed=3)
    A = spmv.SparseTensorCOO(adj, dtype=torch.int16, groups=4)
    A.build_coo()
    assert A.coo[0].size() == (16, 16)             # multiple of 64/16 = 4 (spmv.py:45-51)
    with pytest.raises(AssertionError):
        A.col_split(2)
    # grande.dense_split: pad = ncols[0] rounded to 8 bytes, slices overlap, last one zero padded
    B = torch.arange(3 * 10, dtype=torch.float32).reshape(3, 10)
    pieces = grande.dense_split(B, [4, 3, 3])
    assert [tuple(p.shape) for p in pieces] == [(3, 4), (3, 4), (3, 4)]
    assert torch.equal(pieces[1][:, :3], B[:, 4:7]) and torch.equal(pieces[2][:, :3], B[:, 7:10])
    assert torch.all(pieces[2][:, 3] == 0)


def test_prepare_spmv_requires_coo():
    args = types.SimpleNamespace(data_type=torch.int32, sp_format="CSR", hidden_size=8, sp_parts=1, ds_parts=4)
    with pytest.raises(AssertionError):
        spmv.prepare_pim_spmv(random_adj(8, 8, 0.5), args)


@pytest.mark.parametrize("shape", ["pubmed", "arxiv"])
def test_graph_generator_shape(shape):
    n, nnz, max_deg = graphgen.SHAPES[shape]
    adj = graphgen.synthetic_adj(shape, seed=0)
    rowptr, col, value = adj.csr()
    assert adj.sizes() == [n, n] and adj.nnz() == nnz and value is None
    deg = rowptr[1:] - rowptr[:-1]
    assert int(deg.max()) <= max_deg and int(deg.max()) >= 0.9 * max_deg and int(deg.min()) >= 1
    key = adj.coo()[0] * n + col
    assert bool((key[1:] > key[:-1]).all())        # row-major sorted, no duplicates
    assert 0 <= int(col.min()) and int(col.max()) < n


def test_graph_generator_shard_matches_degree_sequence():
    n, nnz, md = 5000, 200_000, 900
    deg = graphgen.degree_sequence(n, nnz, md, seed=3)
    rp, col = graphgen.synthetic_csr(n, nnz, md, seed=3, rows=(1000, 2500), deg=deg)
    assert torch.equal(rp[1:] - rp[:-1], deg[1000:2500]) and col.numel() == int(deg[1000:2500].sum())
    x = graphgen.reference_features(50, 8, torch.int8, seed=1)
    assert x.dtype == torch.int8 and int(x.min()) >= -8 and int(x.max()) <= 3    # randint(-8, 4), spmm_test.py:70


def test_choose_ds_parts():
    l2 = 126 * 2 ** 20
    assert autotuner.choose_ds_parts(232_965, 64, 4, l2) == 1        # 59.6 MB tile fits the budget
    assert autotuner.choose_ds_parts(232_965, 128, 4, l2) == 2       # 119 MB does not: two 64-column tiles
    assert autotuner.choose_ds_parts(2_449_029, 128, 4, l2) == 1     # products-shape: B >> L2, tiling cannot help
    assert autotuner.choose_ds_parts(1000, 32, 4, l2) == 1
    # a feature row gathered fewer than 32 times per launch has nothing to keep resident: no tiling (arxiv-shape
    # H = 128 is 86 MB, above the budget, but runs as one tile); with enough reuse the size rule decides
    assert autotuner.choose_ds_parts(169_343, 128, 4, l2) == 2
    assert autotuner.choose_ds_parts(169_343, 128, 4, l2, nnz=1_166_243) == 1
    assert autotuner.choose_ds_parts(232_965, 128, 4, l2, nnz=114_615_892) == 2


def test_space_algebra():
    from pygim_b200.utils.space import For, Table, Unit
    s = For("a", [1, 2]) * For("b", "xy")
    assert len(s) == 4 and s.fields() == ("a", "b")
    assert list(s.iter_dict())[1] == {"a": 1, "b": "y"}
    t = Table(["a", "b"], [(7, "z")]) + s
    assert len(t) == 5 and list(t.iter_dict())[0] == {"a": 7, "b": "z"}
    assert list(Unit() * For("c", [3])) == [(("c", 3),)]
    with pytest.raises(RuntimeError):
        For("a", [1]) * For("a", [2])
    with pytest.raises(RuntimeError):
        For("a", [1]) + For("b", [2])
    assert len(Table.from_dicts([{"p": 1, "q": 2}, {"p": 3, "q": 4}])) == 2


def test_autotune_picks_gpu_friendly_splits():
    reddit = autotuner.GraphStats(232_965, 232_965, 114_615_892, 492.0, 21_657, 1.2, 0)
    products = autotuner.GraphStats(2_449_029, 2_449_029, 61_859_140, 25.3, 17_481, 2.0, 0)
    cfg = autotuner.autotune(reddit, 128)
    assert cfg[0] == 1 and cfg[1] == 2 and cfg[2] == "nnz" and cfg[3] == "nnz" and cfg[4]["seg_len"] == 2048
    assert autotuner.autotune(reddit, 32)[:2] == [1, 1]
    assert autotuner.autotune(products, 128)[:2] == [1, 1]           # B >> L2: no tiling
    # the reference's own candidate pairs are accepted and ranked: (1,32) loses to (2,16)?  no - both are poor,
    # but the model must still return one of them
    assert autotuner.autotune(reddit, 256, split_set=[(1, 32), (2, 16)])[:2] in ([1, 32], [2, 16])
    # model sanity: Reddit H=32 within 2x of the measured 0.91 ms, products H=128 within 2x of 5.2 ms
    assert 0.45 < autotuner.predict_ms(reddit, 32, 4, 1, 1) < 1.8
    assert 2.6 < autotuner.predict_ms(products, 128, 4, 1, 1) < 10.4
    st = autotuner.GraphStats.from_rowptr(torch.tensor([0, 2, 2, 5]), 4)
    assert (st.nrows, st.nnz, st.max_degree, st.empty_rows) == (3, 5, 3, 1)


def test_bench_reference_arm_and_byte_model(tmp_path):
    """bench.py --impl reference (the CPU arm the driver runs) on a tiny sample prints one well-formed JSON line;
    the algorithmic-byte model matches SURVEY.md 8d."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    n, nnz = 232_965, 114_615_892
    assert bench.alg_bytes_csr(n, n, nnz, 32) == 4 * (n + 1) + 8 * nnz + 2 * 4 * n * 32          # 977.5 MB
    assert round(bench.alg_bytes_csr(n, n, nnz, 128) / 1e6, 1) == 1156.4
    assert bench.alg_bytes_csr(n, n, nnz, 32, 1, "COO") == 9 * nnz + 2 * n * 32                  # INT8 COO: 1046 MB
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-rows", "512"], stdout=subprocess.PIPE, text=True, check=True)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GFLOP/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    # both arms describe the workload with the same `config` block (the driver compares them)
    assert line["config"] == bench.workload_config("reddit", False, "FLT32", "CSR", [16, 32, 64, 128], n, nnz)
