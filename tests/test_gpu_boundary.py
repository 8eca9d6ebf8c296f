"""The drop-in boundary on the GPU: the registered torch.ops.pim_ops operators with the reference's call shapes,
the raw C ABI with strided buffers, option handling and error reporting."""
import ctypes as C

import pytest
import torch

from helpers import features, make_args, oracle_spmm, random_adj

pytestmark = pytest.mark.gpu


def test_torch_ops_namespace_like_the_reference_drivers(gpu_backend, oracle):
    """spmm.py:57-122 call sequence: to_device_group on int32 tensors, run_group on the dense_split list."""
    import pygim_b200.backend_pim  # noqa: F401  (registers torch.ops.pim_ops)
    ops = torch.ops.pim_ops
    adj = random_adj(120, 120, 0.1, seed=3, value_dtype=torch.int32)
    rowptr, col, val = adj.csr()
    x = features(120, 24, torch.int32, seed=1)
    handle = ops.spmm_csr_to_device_group([rowptr.int()], [col.int()], [val], [120], [120], [12, 12], 24)
    parts = [p.contiguous() for p in torch.chunk(x, 2, 1)]
    out = ops.spmm_csr_run_group(handle, parts)
    assert torch.equal(out, oracle_spmm(oracle, adj, x, torch.int32))
    row, col2, _ = adj.coo()
    h2 = ops.spmm_coo_to_device_group([row.int()], [col2.int()], [val], [120], [120], [24], 24)
    assert torch.equal(ops.spmm_coo_run_group(h2, [x]), out)
    ops.spmm_free_group(handle)
    ops.spmm_free_group(h2)
    assert isinstance(ops.dpu_init_ranks(2), list)          # grande consumes the returned list


def test_grande_run_group_with_padded_slices(gpu_backend, oracle):
    """grande.py:83-107 drives spmm_csr_run_group with padded per-unit column slices; the op reassembles them."""
    from pygim_b200.backend_pim import grande, pim_ops
    adj = random_adj(90, 90, 0.1, seed=4)
    x = features(90, 10, torch.float32, seed=2)
    args = make_args(torch.float32, "CSR", 10, sp_parts=2)
    A = grande.prepare_pim_spmm_grande(adj, args, [3, 3])
    row_blocks = torch.split(x, [p.size(1) for p in A.csr], dim=0)
    B_parts = []
    for i, blk in enumerate(row_blocks):
        B_parts += grande.dense_split(blk, A.dense_ncols[i].tolist())
    out = pim_ops.spmm_csr_run_group(A.sp_info_ptr, B_parts)
    assert torch.equal(out, oracle_spmm(oracle, adj, x, torch.float32))
    assert torch.equal(A.mul(x), out)
    A.free()


def test_c_abi_with_strided_device_buffers(gpu_backend, oracle):
    """pygim_spmm_device on a column window of wider B / C matrices (ldb, ldc > width; 16-byte aligned or not)."""
    from pygim_b200 import _lib
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    lib = _lib.lib()
    adj = random_adj(200, 150, 0.08, seed=6, value_dtype=torch.float32)
    for width, off in ((32, 8), (20, 3)):
        xw = features(150, 64, torch.float32, seed=3).cuda()
        cw = torch.full((200, 80), -1.0, device="cuda")
        A = prepare_pim_spmm(adj.to("cuda"), make_args(torch.float32, "CSR", width))
        b_view, c_view = xw[:, off:off + width], cw[:, off:off + width]
        _lib.check(lib.pygim_spmm_device(A.sp_info_ptr, b_view.data_ptr(), xw.stride(0), c_view.data_ptr(),
                                         cw.stride(0), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        want = oracle_spmm(oracle, adj, b_view.cpu().contiguous(), torch.float32)
        assert torch.equal(c_view.cpu(), want)
        untouched = torch.ones(80, dtype=torch.bool)
        untouched[off:off + width] = False
        assert bool((cw[:, untouched.cuda()] == -1.0).all())          # nothing outside the window was written
        A.free()


def test_options_and_errors(gpu_backend, oracle):
    from pygim_b200 import _lib
    from pygim_b200.backend_pim import pim_ops
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = random_adj(300, 300, 0.2, seed=8, long_row=1)
    x = features(300, 32, torch.float32, seed=4)
    want = oracle_spmm(oracle, adj, x, torch.float32)
    A = prepare_pim_spmm(adj, make_args(torch.float32, "CSR", 32))
    stats0 = pim_ops.plan_stats(A.sp_info_ptr)
    assert stats0["nnz"] == adj.nnz() and stats0["nrows"] == 300
    for key, value in (("seg_len", 32), ("rows_per_ticket", 7), ("short_rows", 1), ("short_rows", 2), ("short_rows", 0),
                       ("unit_values", 0), ("host_chunks", 3), ("l2_persist", 1), ("l2_persist", 0), ("seg_len", -1)):
        pim_ops.plan_set_option(A.sp_info_ptr, key, value)
        assert torch.equal(A.mul(x), want), (key, value)
        assert torch.equal(A.mul(x.cuda()).cpu(), want), (key, value)
    pim_ops.plan_set_option(A.sp_info_ptr, "seg_len", 32)
    assert pim_ops.plan_stats(A.sp_info_ptr)["segments"] > stats0["segments"]
    with pytest.raises(_lib.PygimError, match="unknown option"):
        pim_ops.plan_set_option(A.sp_info_ptr, "no_such_knob", 1)
    with pytest.raises(_lib.PygimError, match="dtype"):
        A.mul(x.double())
    with pytest.raises(AssertionError):
        A.mul(x[:, :16])                                              # hidden_size mismatch (spmm.py:114)
    timers = pim_ops.last_timers(A.sp_info_ptr)
    assert set(timers) == {"load_sparse_time", "load_dense_time", "kernel_time", "retrieve_result_time",
                           "alignment_time"} and timers["alignment_time"] == 0.0
    A.free()
    with pytest.raises(_lib.PygimError, match="freed"):
        pim_ops.spmm_run_dense(12345, x)


def test_example_drivers_run(gpu_backend):
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, argv in (("spmm_test", ["--dataset", "PubMed", "--version", "spmm", "--hidden_size", "32",
                                      "--data_type", "INT8", "--sp_format", "COO", "--sp_parts", "2", "--repeat", "1"]),
                       ("spmm_test", ["--dataset", "PubMed", "--version", "spmv", "--hidden_size", "32",
                                      "--data_type", "INT32", "--sp_format", "COO", "--ds_parts", "8", "--repeat", "1"]),
                       ("inference", ["--dataset", "PubMed", "--model", "sage", "--hidden_size", "32", "--repeat", "1",
                                      "--data_type", "INT16", "--sp_format", "CSR"])):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "examples", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out = mod.main(mod.get_args(argv))
        assert out is not None and torch.isfinite(out.float()).all()
    from pygim_b200.backend_pim import pim_ops
    pim_ops.dpu_init_ranks(1)     # the drivers release the backend; restore it for the session fixture


@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_cuda_graph_capture_and_replay(gpu_backend, oracle, fmt):
    """The device entry point only enqueues stream work and keeps no host-side launch state (self-resetting ticket
    counters), so a `mul` can be captured in a CUDA graph and replayed on new operand contents."""
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    from pygim_b200 import graphgen
    adj = graphgen.synthetic_adj("reddit", scale=0.004, seed=9)
    n = adj.size(0)
    A = prepare_pim_spmm(adj.to("cuda"), make_args(torch.float32, fmt, 64))
    x = torch.zeros((n, 64), device="cuda")
    out = torch.empty((n, 64), device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):                       # warm-up: scratch buffers are allocated outside the capture
            A.mul(x, out=out)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        A.mul(x, out=out)
    for seed in (1, 2, 3):
        xs = features(n, 64, torch.float32, seed=seed)
        x.copy_(xs)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), oracle_spmm(oracle, adj, xs, torch.float32)), seed
    A.free()


@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_pipelined_host_entry_point(gpu_backend, oracle, fmt):
    """Host operands with the upload/compute/download pipeline forced on (column tiles of 256 bytes by default, 128 /
    512 by option, a 128-byte remainder tile + row chunks of the last tile), with several sparse and dense parts,
    16-byte aligned and unaligned widths."""
    from pygim_b200.backend_pim import pim_ops
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    for dtype, hidden, sp, ds in ((torch.float32, 128, 1, 1), (torch.float32, 96, 2, 2), (torch.int8, 256, 1, 1),
                                  (torch.int32, 72, 3, 1), (torch.float64, 33, 1, 2)):
        adj = random_adj(700, 700, 0.05, seed=hidden, long_row=9, value_dtype=dtype)
        x = features(700, hidden, dtype, seed=2)
        want = oracle_spmm(oracle, adj, x, dtype)
        A = prepare_pim_spmm(adj, make_args(dtype, fmt, hidden, sp_parts=sp, ds_parts=ds))
        for chunks, tile in ((3, 128), (3, -1), (1, 512), (0, -1)):
            pim_ops.plan_set_option(A.sp_info_ptr, "host_chunks", chunks)
            pim_ops.plan_set_option(A.sp_info_ptr, "host_tile_bytes", tile)
            out = torch.empty((700, hidden), dtype=dtype).pin_memory()
            got = A.mul(x.pin_memory(), out=out)
            assert torch.equal(got, want), (dtype, hidden, sp, ds, chunks, tile)
            assert torch.equal(A.mul(x), want), (dtype, hidden, sp, ds, chunks, tile, "pageable")
        A.free()
