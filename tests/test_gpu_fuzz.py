"""Seeded random sweep over shapes, dtypes, formats, partitionings and plan options: the CUDA path must equal the
oracle bit for bit on every case (integer-valued inputs, so floats are exact too)."""
import numpy as np
import pytest
import torch

from helpers import ALL_DTYPES, features, make_args, oracle_spmm, random_adj

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(40))
def test_random_configuration(gpu_backend, oracle, seed):
    from pygim_b200.backend_pim import pim_ops
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    rng = np.random.default_rng(1000 + seed)
    dtype = ALL_DTYPES[int(rng.integers(len(ALL_DTYPES)))]
    fmt = "CSR" if rng.random() < 0.5 else "COO"
    n = int(rng.integers(1, 900))
    m = int(rng.integers(1, 900))
    hidden = int(rng.choice([1, 2, 3, 4, 7, 8, 16, 24, 32, 33, 48, 64, 100, 128, 160, 256]))
    density = float(rng.choice([0.0, 0.002, 0.02, 0.1, 0.5]))
    sp = int(rng.integers(1, min(m, 4) + 1))
    ds = int(rng.integers(1, min(hidden, 4) + 1))
    if -(-hidden // ds) * (ds - 1) >= hidden:        # split_widths would produce an empty / negative last part
        ds = 1
    with_values = rng.random() < 0.6
    long_row = int(rng.integers(n)) if (rng.random() < 0.5 and m > 8) else None
    empty = tuple(int(r) for r in rng.integers(0, n, size=int(rng.integers(0, 4))))
    adj = random_adj(n, m, density, seed=seed, value_dtype=dtype if with_values else None, empty_rows=empty,
                     long_row=long_row if long_row not in empty else None)
    x = features(m, hidden, dtype, seed=seed)
    want = oracle_spmm(oracle, adj, x, dtype)
    on_gpu = rng.random() < 0.5
    A = prepare_pim_spmm(adj.to("cuda") if on_gpu else adj, make_args(dtype, fmt, hidden, sp_parts=sp, ds_parts=ds))
    if rng.random() < 0.5:
        pim_ops.plan_set_option(A.sp_info_ptr, "seg_len", int(rng.choice([32, 64, 128])))
    if rng.random() < 0.3:
        pim_ops.plan_set_option(A.sp_info_ptr, "rows_per_ticket", int(rng.integers(1, 32)))
    if rng.random() < 0.3:
        pim_ops.plan_set_option(A.sp_info_ptr, "chunk_nnz", int(rng.choice([32, 64, 256])))
    if rng.random() < 0.6:
        pim_ops.plan_set_option(A.sp_info_ptr, "short_rows", int(rng.integers(0, 3)))   # 2 = streamed row tickets
    if rng.random() < 0.3:
        pim_ops.plan_set_option(A.sp_info_ptr, "host_chunks", int(rng.integers(1, 4)))
    info = (seed, dtype, fmt, n, m, hidden, density, sp, ds, with_values, on_gpu)
    got_dev = A.mul(x.cuda())
    got_host = A.mul(x)
    torch.cuda.synchronize()
    assert torch.equal(got_dev.cpu(), want), info
    assert torch.equal(got_host, want), info
    A.free()
