#!/usr/bin/env python
"""How much of a SMALL SpMM call is host launch overhead?  Times the same plan (a) call by call through the Python
API with CUDA events around each call and (b) as a CUDA-graph replay of the same calls.  Shapes: a 1/8 Reddit-shape
row shard (what each GPU runs at N = 8) and arxiv-shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygim_b200 import graphgen  # noqa: E402
from pygim_b200.backend_pim import pim_ops  # noqa: E402
from pygim_b200.backend_pim.spmm import SparseTensorCOO  # noqa: E402
from pygim_b200.sparse_tensor import SparseTensor  # noqa: E402

pim_ops.dpu_init_ranks(1)
dev = "cuda"


def plans_for(shape, frac):
    n, nnz, max_deg = graphgen.SHAPES[shape]
    deg = graphgen.degree_sequence(n, nnz, max_deg, n, seed=0)
    rp = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(deg, 0, out=rp[1:])
    splits = pim_ops.partition_rows_by_nnz(rp, frac) if frac > 1 else [0, n]
    rowptr, col = graphgen.synthetic_csr(n, nnz, max_deg, seed=0, device=dev, rows=(splits[0], splits[1]), deg=deg)
    adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(splits[1], n), is_sorted=True)
    out = {}
    for h in (16, 32, 64, 128):
        A = SparseTensorCOO(adj, dtype=torch.float32, format="CSR")
        A.to_pim_group(h, 1)
        for kv in filter(None, os.environ.get("PROBE_OPTS", "").split(",")):
            k, v = kv.split("=")
            pim_ops.plan_set_option(A.sp_info_ptr, k, int(v))
        out[h] = (A, graphgen.reference_features(n, h, torch.float32, seed=h, device=dev),
                  torch.empty((splits[1], h), device=dev))
    return out


for shape, frac in (("reddit", 8), ("arxiv", 1)):
    P = plans_for(shape, frac)
    for h, (A, x, c) in P.items():
        for _ in range(5):
            A.mul(x, out=c)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for e0, e1 in ev:
            e0.record()
            A.mul(x, out=c)
            e1.record()
        torch.cuda.synchronize()
        per_call = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[25] * 1e3
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(50):
            A.mul(x, out=c)
        t1.record()
        torch.cuda.synchronize()
        back_to_back = t0.elapsed_time(t1) / 50 * 1e3
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            A.mul(x, out=c)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                A.mul(x, out=c)
        g.replay()
        torch.cuda.synchronize()
        t0.record()
        for _ in range(5):
            g.replay()
        t1.record()
        torch.cuda.synchronize()
        graphed = t0.elapsed_time(t1) / 100 * 1e3
        print("%-8s 1/%d  H=%3d   per-call events %7.1f us   back-to-back %7.1f us   CUDA graph %7.1f us" %
              (shape, frac, h, per_call, back_to_back, graphed), flush=True)
    # the sweep as it runs in the benchmark: the four operands one after the other, so every launch finds ITS dense
    # operand evicted from L2 by the previous ones (the per-H loops above re-run one operand, which stays resident)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for A, x, c in P.values():
            A.mul(x, out=c)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5):
            for A, x, c in P.values():
                A.mul(x, out=c)
    g.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        g.replay()
    t1.record()
    torch.cuda.synchronize()
    print("%-8s 1/%d  sweep 16+32+64+128 as one CUDA graph: %7.1f us per sweep" % (shape, frac, t0.elapsed_time(t1) / 25 * 1e3),
          flush=True)
    for A, _, _ in P.values():
        A.free()
