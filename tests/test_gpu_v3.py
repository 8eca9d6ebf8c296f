"""GPU parity of the round-2 machinery: SM-affine supertickets + stealing under every scheduling option, vector
index loads on misaligned views, row maps / prepare-time clustering, sorted-COO-as-CSR and the unsorted COO
fallback, the fused quantise / de-quantise / residual epilogue, in-kernel arrival flags, per-stream plan state and
never-reused handles.  Everything is compared with the CPU oracle (integers and integer-valued floats bit-exact)."""
import numpy as np
import pytest
import torch

from helpers import ALL_DTYPES, NP_DTYPES, features, make_args, oracle_spmm, random_adj

pytestmark = pytest.mark.gpu


def _reddit_like(scale=0.01, seed=2):
    from pygim_b200 import graphgen
    return graphgen.synthetic_adj("reddit", scale=scale, seed=seed)


def _want(oracle, adj, x):
    rowptr, col, val = adj.csr()
    v = None if val is None else val.numpy()
    return torch.from_numpy(oracle.spmm_csr_rowpar(rowptr.numpy(), col.numpy(), v, x.numpy()))


@pytest.mark.parametrize("opts", [
    dict(item_nnz=32), dict(item_nnz=2048), dict(super_nnz=64), dict(super_nnz=1 << 20), dict(cta_threads=1024),
    dict(cta_threads=512, max_g=8), dict(max_g=4), dict(max_g=1), dict(rows_per_ticket=1), dict(seg_len=64, item_nnz=64),
    dict(short_rows=2, item_nnz=96), dict(short_rows=1), dict(short_rows=2, max_g=8, cta_threads=1024),
    dict(short_rows=3), dict(short_rows=4), dict(short_rows=4, seg_len=64, max_g=2),
])
def test_scheduling_options_do_not_change_results(gpu_backend, oracle, opts):
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = _reddit_like()
    n = adj.size(0)
    for dtype, hidden in ((torch.float32, 128), (torch.float32, 48), (torch.int8, 64), (torch.int64, 24), (torch.int16, 7)):
        x = features(n, hidden, dtype, seed=3)
        want = _want(oracle, adj, x)
        A = prepare_pim_spmm(adj.to("cuda"), make_args(dtype, "CSR", hidden))
        for k, v in opts.items():
            gpu_backend.plan_set_option(A.sp_info_ptr, k, v)
        for _ in range(2):                       # counters return to zero between launches
            got = A.mul(x.cuda())
        torch.cuda.synchronize()
        assert torch.equal(got.cpu(), want), (dtype, hidden, opts)
        A.free()


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_weighted_values_on_misaligned_views(gpu_backend, oracle, dtype):
    """Index/value arrays that start 1, 2, 3 elements past a 16-byte boundary (row shards are such views), with
    explicit values: the vector index loads use the arrays' own alignment; when colind and val are NOT co-aligned the
    kernel falls back to element loads."""
    from pygim_b200.sharded import shard_rows
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = random_adj(400, 380, 0.08, seed=11, value_dtype=dtype, long_row=17)
    x = features(380, 32, dtype, seed=4)
    full = oracle_spmm(oracle, adj, x, dtype)
    rowptr = adj.csr()[0]
    for r0 in (1, 2, 3, 7, 50):
        if int(rowptr[r0]) % 4 == 0:
            continue
        sh = shard_rows(adj.to("cuda"), r0, 400)
        A = prepare_pim_spmm(sh, make_args(dtype, "CSR", 32))
        got = A.mul(x.cuda())
        torch.cuda.synchronize()
        assert torch.equal(got.cpu(), full[r0:]), (dtype, r0)
        A.free()
    # raw op with hand-made views: co-aligned at every offset (vector loads), and misaligned differently (element loads)
    rp, col, val = adj.to("cuda").csr()
    for oc, ov in ((0, 0), (1, 1), (2, 2), (3, 3), (1, 2), (0, 3)):
        pad_c = torch.zeros(col.numel() + 8, dtype=torch.int32, device="cuda")
        pad_v = torch.zeros(val.numel() + 8, dtype=dtype, device="cuda")
        c_view, v_view = pad_c[oc:oc + col.numel()], pad_v[ov:ov + val.numel()]
        c_view.copy_(col.int())
        v_view.copy_(val.type(dtype))
        h = gpu_backend.spmm_csr_to_device_group([rp.int()], [c_view], [v_view], [400], [380], [32], 32)
        got = gpu_backend.spmm_csr_run_group(h, [x.cuda()])
        torch.cuda.synchronize()
        assert torch.equal(got.cpu(), full), (dtype, oc, ov)
        gpu_backend.spmm_free_group(h)


@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_row_map_and_prepare_time_clustering(gpu_backend, oracle, fmt):
    """A row-permuted plan with a row map returns rows in the ORIGINAL order, bit for bit; the shared-neighbour
    clustering finds the communities of a block-model graph whose node numbering hides them."""
    from pygim_b200 import graphgen, reorder
    from pygim_b200.backend_pim.spmm import SparseTensorCOO, prepare_pim_spmm
    from pygim_b200.sparse_tensor import SparseTensor
    n, nnz = 6000, 600_000
    rowptr, col = graphgen.clustered_csr(n, nnz, 900, seed=3, community=300, p_in=0.7)
    adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(n, n), is_sorted=True)
    for dtype, hidden in ((torch.float32, 32), (torch.int32, 64), (torch.int8, 16)):
        x = features(n, hidden, dtype, seed=9)
        want = _want(oracle, adj, x)
        # (a) arbitrary permutation
        g = torch.Generator().manual_seed(1)
        perm = torch.randperm(n, generator=g)
        A = SparseTensorCOO(reorder.permute_rows(adj, perm).to("cuda"), dtype=dtype, format=fmt)
        A.row_perm = perm.cuda()
        A.to_pim_group(hidden, 1)
        assert gpu_backend.plan_layout(A.sp_info_ptr)["row_map"] == 1
        assert torch.equal(A.mul(x.cuda()).cpu(), want), (dtype, "randperm")
        assert torch.equal(A.mul(x), want), (dtype, "randperm, host operand")
        A.free()
        # (b) through the public switch
        args = make_args(dtype, fmt, hidden)
        args.reorder = "cluster"
        A = prepare_pim_spmm(adj.to("cuda"), args)
        assert torch.equal(A.mul(x.cuda()).cpu(), want), (dtype, "cluster")
        A.free()
    perm, stats = reorder.cluster_rows(adj.to("cuda"), n_pivots=128, seed=0)
    assert sorted(perm.cpu().tolist()) == list(range(n))
    # hidden community of every node (the generator's layout): consecutive rows of the new order mostly share one
    pi = torch.randperm(n, generator=torch.Generator().manual_seed(3 + 17))
    pos = torch.empty(n, dtype=torch.int64)
    pos[pi] = torch.arange(n)
    comm = (pos // 300)[perm.cpu()]
    same = float((comm[1:] == comm[:-1]).float().mean())
    assert same > 0.85 and stats["assigned"] > 0.9, (same, stats)


@pytest.mark.parametrize("weighted", [False, True])
def test_hot_cold_tiles(gpu_backend, oracle, weighted):
    """reorder="tiles": rows clustered, hot feature rows staged in shared memory (csrc/spmm_csr_hc.cuh).  Same results
    as the oracle in the ORIGINAL row order - across column chunks (H x s > 128 bytes), segments of long rows (all
    cold), weighted values moved along with their nonzeros, host and device operands."""
    from pygim_b200 import graphgen, reorder
    from pygim_b200.backend_pim.spmm import SparseTensorCOO, prepare_pim_spmm
    from pygim_b200.sparse_tensor import SparseTensor
    n, nnz = 6000, 600_000
    rowptr, col = graphgen.clustered_csr(n, nnz, 2500, seed=4, community=300, p_in=0.7)
    for dtype, hidden in ((torch.float32, 32), (torch.float32, 128), (torch.float32, 16), (torch.int32, 48), (torch.int8, 64),
                          (torch.float64, 8), (torch.int16, 40)):
        val = None
        if weighted:
            val = torch.from_numpy(np.random.default_rng(1).integers(-3, 4, nnz)).to(dtype)
        adj = SparseTensor(rowptr=rowptr, col=col, value=val, sparse_sizes=(n, n), is_sorted=True)
        x = features(n, hidden, dtype, seed=5)
        want = _want(oracle, adj, x)
        args = make_args(dtype, "CSR", hidden)
        args.reorder = "tiles"
        A = prepare_pim_spmm(adj.to("cuda"), args)
        assert A.hot_plan is not None and A.hot_plan["coverage"] > 0.3, (dtype, hidden)
        for _ in range(2):
            got = A.mul(x.cuda())
        torch.cuda.synchronize()
        assert torch.equal(got.cpu(), want), (dtype, hidden, weighted)
        assert torch.equal(A.mul(x), want), (dtype, hidden, weighted, "host operand")
        A.free()
    # explicit small tile, tiny supertickets, long rows cut into segments
    adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(n, n), is_sorted=True).to("cuda")
    radj, perm, stats = reorder.reorder_rows(adj, "tiles", n_pivots=64)
    hc, hot = reorder.hot_cold_plan(radj, stats["group_of_row"], hot_k=96, seg_len=128, super_nnz=4096)
    x = features(n, 64, torch.float32, seed=6)
    A = SparseTensorCOO(hc, dtype=torch.float32, format="CSR")
    A.row_perm, A.hot_plan = perm, hot
    A.to_pim_group(64, 1)
    assert gpu_backend.plan_stats(A.sp_info_ptr)["segments"] > 0
    assert torch.equal(A.mul(x.cuda()).cpu(), _want(oracle, adj.to("cpu"), x))
    A.free()


def test_sorted_coo_runs_through_the_csr_kernels_and_unsorted_coo_is_still_right(gpu_backend, oracle):
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = _reddit_like(0.006, seed=7)
    n = adj.size(0)
    for dtype in (torch.float32, torch.int8, torch.int32):
        x = features(n, 32, dtype, seed=1)
        want = _want(oracle, adj, x)
        A = prepare_pim_spmm(adj.to("cuda"), make_args(dtype, "COO", 32))
        lay = gpu_backend.plan_layout(A.sp_info_ptr)
        assert lay["coo_sorted"] == 1 and lay["coo_runs_as_csr"] == 1
        assert torch.equal(A.mul(x.cuda()).cpu(), want)
        assert gpu_backend.last_launches(A.sp_info_ptr) == 1          # no zero-fill launch, one kernel
        gpu_backend.plan_set_option(A.sp_info_ptr, "coo_native", 1)  # the segmented-reduction COO kernel
        assert gpu_backend.plan_layout(A.sp_info_ptr)["coo_runs_as_csr"] == 0
        assert torch.equal(A.mul(x.cuda()).cpu(), want)
        A.free()
        # the raw op with an UNSORTED stream (the C ABI does not require what .coalesce() guarantees)
        row, col, _ = adj.coo()
        g = torch.Generator().manual_seed(5)
        shuffle = torch.randperm(row.numel(), generator=g)
        val = torch.ones(row.numel(), dtype=dtype)
        h = gpu_backend.spmm_coo_to_device_group([row[shuffle].int().cuda()], [col[shuffle].int().cuda()], [val.cuda()],
                                                 [n], [n], [32], 32)
        assert gpu_backend.plan_layout(h)["coo_sorted"] == 0
        got = gpu_backend.spmm_coo_run_group(h, [x.cuda()])
        torch.cuda.synchronize()
        assert torch.equal(got.cpu(), want), dtype            # integer-valued inputs: exact in any order
        gpu_backend.spmm_free_group(h)


@pytest.mark.parametrize("dtype", [torch.int8, torch.int16, torch.int32, torch.float32])
@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_fused_quantise_dequantise_residual_epilogue(gpu_backend, oracle, dtype, fmt):
    """quantize kernel == symmetric_quantize; SpMM with scale (+ residual) == the torch expression, bit for bit."""
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    from pygim_b200.models.quantize import symmetric_dequantize, symmetric_quantize
    adj = _reddit_like(0.006, seed=4)
    n = adj.size(0)
    torch.manual_seed(3)
    for hidden in (16, 64, 40):
        x = torch.randn(n, hidden) * 3.0
        scale_ref, xq_ref = symmetric_quantize(x, dtype)
        scale, xq = gpu_backend.quantize(x.cuda(), dtype)
        assert float(scale) == float(scale_ref) and torch.equal(xq.cpu(), xq_ref), (dtype, hidden)
        out_q = _want(oracle, adj, xq_ref)
        want = symmetric_dequantize(out_q, 1.0, scale_ref)
        A = prepare_pim_spmm(adj.to("cuda"), make_args(dtype, fmt, hidden))
        got = gpu_backend.spmm_run_dense_ex(A.sp_info_ptr, xq, scale=scale)
        assert got.dtype == torch.float32 and torch.equal(got.cpu(), want), (dtype, hidden, "dequantise")
        eps = torch.tensor([0.25])
        coeff = float((1 + eps).item())
        want_r = want + (1 + eps) * x
        got_r = A.mul_fused(x.cuda(), residual=x.cuda(), residual_coeff=coeff)
        assert torch.equal(got_r.cpu(), want_r), (dtype, hidden, "residual")
        A.free()


def test_single_gpu_peer_stores_arrival_flags_and_halo_mask(gpu_backend, oracle):
    """The fused all-gather epilogue with this GPU as its own (only) peer: rows land at row_offset of the peer
    buffer, the last warp raises the arrival flag, wait_flags returns, and a peer mask suppresses rows."""
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = _reddit_like(0.006, seed=8)
    n = adj.size(0)
    for dtype, hidden in ((torch.float32, 32), (torch.int8, 16), (torch.float64, 8)):
        x = features(n, hidden, dtype, seed=2)
        want = _want(oracle, adj, x)
        A = prepare_pim_spmm(adj.to("cuda"), make_args(dtype, "CSR", hidden))
        big = torch.full((n + 10, hidden), 7, dtype=dtype, device="cuda")
        flags = torch.zeros(4, dtype=torch.int32, device="cuda")
        mask = torch.ones(n, dtype=torch.uint8, device="cuda")
        mask[::3] = 0
        gpu_backend.spmm_run_dense_ex(A.sp_info_ptr, x.cuda(), peer_ptrs=[big.data_ptr()], ldc=hidden, row_offset=5,
                                      peer_mask=mask, flag_ptrs=[flags.data_ptr()], my_rank=2, epoch=41)
        gpu_backend.wait_flags(flags[2:3], 41)
        torch.cuda.synchronize()
        assert flags.cpu().tolist() == [0, 0, 41, 0]
        got = big.cpu()
        keep = mask.cpu().bool()
        assert torch.equal(got[5:5 + n][keep], want[keep]) and bool((got[5:5 + n][~keep] == 7).all())
        assert bool((got[:5] == 7).all()) and bool((got[5 + n:] == 7).all())
        A.free()


def test_one_handle_on_two_streams(gpu_backend, oracle):
    """Launch state (draw counters, partial sums) is per (plan, stream): interleaved launches of one handle on two
    streams do not disturb each other."""
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = _reddit_like(0.01, seed=6)
    n = adj.size(0)
    A = prepare_pim_spmm(adj.to("cuda"), make_args(torch.float32, "CSR", 64))
    gpu_backend.plan_set_option(A.sp_info_ptr, "seg_len", 128)       # many segmented rows => partial sums in use
    xs = [features(n, 64, torch.float32, seed=s).cuda() for s in (1, 2)]
    wants = [_want(oracle, adj, x.cpu()) for x in xs]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    torch.cuda.synchronize()
    for _ in range(6):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                outs[k].append(A.mul(xs[k]))
    torch.cuda.synchronize()
    for k in (0, 1):
        for o in outs[k]:
            assert torch.equal(o.cpu(), wants[k]), k
    A.free()


def test_handles_are_never_reused(gpu_backend, oracle):
    """prepare; release; init; prepare: the old object's late free() must not hit the new plan (round-1 advice)."""
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = random_adj(100, 100, 0.1, seed=1)
    x = features(100, 16, torch.float32)
    old = prepare_pim_spmm(adj, make_args(torch.float32, "CSR", 16))
    h_old = old.sp_info_ptr
    gpu_backend.dpu_release()
    gpu_backend.dpu_init_ranks(1)
    new = prepare_pim_spmm(adj, make_args(torch.float32, "CSR", 16))
    assert new.sp_info_ptr != h_old
    del old                                    # __del__ -> free() of a stale handle: a no-op
    assert torch.equal(new.mul(x), oracle_spmm(oracle, adj, x, torch.float32))
    with pytest.raises(Exception):
        gpu_backend.plan_stats(h_old)
    new.free()


def test_spmv_batch_is_one_launch(gpu_backend, oracle):
    from pygim_b200.backend_pim.spmv import prepare_pim_spmv
    adj = random_adj(203, 203, 0.05, seed=12)
    x = features(203, 64, torch.int32, seed=7)
    A = prepare_pim_spmv(adj, make_args(torch.int32, "COO", 64, 1, 32))
    out = A.mul(x)
    assert torch.equal(out, oracle_spmm(oracle, adj, x, torch.int32))
    # one 32-column batch per call, not one launch per vector (this degree-10 graph takes the two-launch family:
    # tiny rows + the rest)
    assert gpu_backend.last_launches(A.sp_info_ptr) <= 2
    # the op-level surface: `groups` single-column vectors
    pad = A.coo[0].size(1) - 203
    xb = torch.nn.functional.pad(x[:, :32], (0, 0, 0, pad))
    res = gpu_backend.spmv_coo_run_group(A.sp_info_ptr, [xb[:, k:k + 1].contiguous() for k in range(32)])
    assert torch.equal(res[:203], oracle_spmm(oracle, adj, x[:, :32].contiguous(), torch.int32))
    A.free()


def test_autotuned_pick_is_close_to_the_best_candidate(gpu_backend, oracle):
    """prepare_pim_spmm(..., tune=True) consults utils.autotuner; its analytic pick must be within 10 % of the best
    option set of the candidate space when they are all timed on the device."""
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    from pygim_b200.utils import autotuner
    from pygim_b200 import graphgen
    for shape, scale, hidden in (("reddit", 0.05, 64), ("products", 0.05, 32)):
        adj = graphgen.synthetic_adj(shape, scale=scale, seed=1).to("cuda")
        n = adj.size(0)
        args = make_args(torch.float32, "CSR", hidden)
        args.tune = True
        A = prepare_pim_spmm(adj, args)
        x = features(n, hidden, torch.float32, seed=1).cuda()
        stats = autotuner.GraphStats.from_rowptr(adj.csr()[0], n)
        cands = autotuner.candidate_options(stats, hidden, 4)
        assert cands[0] == autotuner.kernel_options(stats, hidden, 4) == A.plan_options
        times = autotuner.measure_options(A, x, cands, repeats=7)
        assert times[0] <= 1.10 * min(times), (shape, list(zip(cands, times)))
        A.free()


def test_batched_host_pipeline_matches_the_oracle(gpu_backend, oracle):
    """pygim_spmm_run_many_host: the hidden-size sweep as ONE upload / compute / download pipeline over pinned host
    operands.  Big enough that the 128-byte column tiles and the row chunks of the last tile are exercised
    (>= 8 MB results), CSR and sorted COO, float and integer; results equal the oracle element for element; a
    second pass re-uses the staging buffers and the row-chunk plans."""
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = _reddit_like(scale=0.12, seed=5)
    n = adj.size(0)
    adj_d = adj.to("cuda")
    cases = [(torch.float32, "CSR", 128), (torch.float32, "CSR", 16), (torch.int32, "COO", 96), (torch.float32, "CSR", 64),
             (torch.int8, "CSR", 32), (torch.float32, "COO", 7)]
    plans, xs, outs, wants = [], [], [], []
    for dtype, fmt, hidden in cases:
        x = features(n, hidden, dtype, seed=hidden)
        wants.append(_want(oracle, adj, x))
        plans.append(prepare_pim_spmm(adj_d, make_args(dtype, fmt, hidden)))
        xs.append(x.pin_memory())
        outs.append(torch.empty((n, hidden), dtype=dtype).pin_memory())
    for _ in range(2):
        for o in outs:
            o.fill_(99)
        gpu_backend.spmm_run_dense_many([p.sp_info_ptr for p in plans], xs, outs)
        for (dtype, fmt, hidden), got, want in zip(cases, outs, wants):
            assert torch.equal(got, want), (dtype, fmt, hidden)
    # a strided (column-sliced) result and operand: the batch honours the row strides
    wide_x = torch.zeros((n, 160), dtype=torch.float32).pin_memory()
    wide_c = torch.full((n, 200), -1.0).pin_memory()
    wide_x[:, 16:144] = xs[0]
    gpu_backend.spmm_run_dense_many([plans[0].sp_info_ptr], [wide_x[:, 16:144]], [wide_c[:, 8:136]])
    assert torch.equal(wide_c[:, 8:136], wants[0]) and bool((wide_c[:, :8] == -1).all()) and bool((wide_c[:, 136:] == -1).all())
    # one plan twice in a batch is refused (its staging buffers are per plan), nothing is left in flight
    with pytest.raises(Exception):
        gpu_backend.spmm_run_dense_many([plans[1].sp_info_ptr, plans[1].sp_info_ptr], [xs[1], xs[1]], [outs[1], outs[1].clone()])
    for p in plans:
        p.free()


@pytest.mark.parametrize("fmt", ["CSR", "COO"])
def test_tiny_row_launch_on_a_citation_shaped_graph(gpu_backend, oracle, fmt):
    """Graphs of mean degree < 12 take the two-launch family by default (short_rows = 4): rows of at most 8 nonzeros
    by a matrix-wide grid of lane groups (csr_tiny_rows_kernel), the others - including segmented hubs - by the
    persistent kernel.  Empty rows, rows of exactly 8 and 9 nonzeros, every dtype, explicit values, ragged widths,
    accumulating sparse parts; bit-exact against the oracle."""
    from pygim_b200 import graphgen
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = graphgen.synthetic_adj("arxiv", scale=0.2, seed=4)
    n = adj.size(0)
    rowptr, col, _ = adj.csr()
    deg = rowptr[1:] - rowptr[:-1]
    assert int((deg <= 8).sum()) > n // 2 and int(deg.max()) > 512 and bool((deg == 8).any()) and bool((deg == 9).any())
    for dtype, hidden in ((torch.float32, 32), (torch.float32, 128), (torch.float32, 5), (torch.int8, 48), (torch.int16, 64),
                          (torch.int32, 16), (torch.int64, 20), (torch.float64, 256)):
        g = torch.Generator().manual_seed(hidden)
        val = torch.randint(-3, 4, (col.numel(),), generator=g, dtype=torch.int32).to(dtype)
        for value in (None, val):
            a = type(adj)(rowptr=rowptr, col=col, value=value, sparse_sizes=(n, n), is_sorted=True)
            x = features(n, hidden, dtype, seed=hidden + 1)
            want = torch.from_numpy(oracle.spmm_csr_rowpar(rowptr.numpy(), col.numpy(), None if value is None else value.numpy(),
                                                           x.numpy()))
            for sp in (1, 2):
                A = prepare_pim_spmm(a.to("cuda"), make_args(dtype, fmt, hidden, sp_parts=sp))
                got = A.mul(x.cuda())
                got = A.mul(x.cuda())
                torch.cuda.synchronize()
                assert torch.equal(got.cpu(), want), (dtype, hidden, fmt, value is not None, sp)
                if hidden >= 128:      # host operands: column tiles + row-chunk plans of the two-launch family
                    gpu_backend.plan_set_option(A.sp_info_ptr, "host_chunks", 3)
                    assert torch.equal(A.mul(x.pin_memory()), want), (dtype, hidden, fmt, value is not None, sp, "host")
                A.free()


def test_tiny_row_family_with_a_row_map_and_options(gpu_backend, oracle):
    """The two-launch family under a row permutation (SM-affine piece supertickets with stealing, both launches
    scatter through the row map), under explicit seg_len / super_nnz / cta_threads options, forced onto a graph that
    would not choose it, and switched off again on one plan; results in the original row order, bit-exact."""
    from pygim_b200 import graphgen
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    for shape, scale in (("arxiv", 0.2), ("products", 0.01)):
        adj = graphgen.synthetic_adj(shape, scale=scale, seed=9)
        n = adj.size(0)
        for dtype, hidden in ((torch.float32, 64), (torch.int16, 24)):
            x = features(n, hidden, dtype, seed=5)
            want = _want(oracle, adj, x)
            for reorder in (None, "degree", "cluster"):
                args = make_args(dtype, "CSR", hidden)
                if reorder:
                    args.reorder = reorder
                A = prepare_pim_spmm(adj.to("cuda"), args)
                for opts in (dict(short_rows=4), dict(short_rows=4, seg_len=32, super_nnz=4096), dict(short_rows=4, seg_len=4), dict(short_rows=4, cta_threads=512),
                             dict(short_rows=3), dict(short_rows=4)):
                    for k, v in opts.items():
                        gpu_backend.plan_set_option(A.sp_info_ptr, k, v)
                    got = A.mul(x.cuda())
                    got = A.mul(x.cuda())
                    torch.cuda.synchronize()
                    assert torch.equal(got.cpu(), want), (shape, dtype, hidden, reorder, opts)
                A.free()
