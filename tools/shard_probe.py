#!/usr/bin/env python
"""Tuning helper (1 GPU): time the CSR kernel on ONE row shard of the Reddit-shaped graph (what a rank of an
N-GPU run owns) for several seg_len / rows_per_ticket settings."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygim_b200 import graphgen  # noqa: E402
from pygim_b200.backend_pim import pim_ops  # noqa: E402
from pygim_b200.backend_pim.spmm import prepare_pim_spmm  # noqa: E402
from pygim_b200.sparse_tensor import SparseTensor  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, nnz, md = graphgen.SHAPES["reddit"]
pim_ops.dpu_init_ranks(1)
deg = graphgen.degree_sequence(n, nnz, md, n, seed=0)
rp = torch.zeros(n + 1, dtype=torch.int64)
torch.cumsum(deg, 0, out=rp[1:])
splits = pim_ops.partition_rows_by_nnz(rp, world)
r0, r1 = splits[0], splits[1]
rowptr, col = graphgen.synthetic_csr(n, nnz, md, seed=0, device="cuda", rows=(r0, r1), deg=deg)
adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(r1 - r0, n), is_sorted=True)
print("shard rows", r1 - r0, "nnz", col.numel())
for hidden in (32, 64):
    x = graphgen.reference_features(n, hidden, torch.float32, device="cuda")
    args = types.SimpleNamespace(data_type=torch.float32, sp_format="CSR", hidden_size=hidden, sp_parts=1, ds_parts=1)
    A = prepare_pim_spmm(adj, args)
    out = torch.empty((r1 - r0, hidden), device="cuda")
    for seg in (-1, 256, 512, 1024, 2048, 4096, 1 << 30):
        pim_ops.plan_set_option(A.sp_info_ptr, "seg_len", seg)
        st = pim_ops.plan_stats(A.sp_info_ptr)
        for _ in range(5):
            A.mul(x, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            A.mul(x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("H=%d seg_len=%s (eff %d, %d segs): %.3f ms  gather %.1f TB/s" % (
            hidden, seg, st["seg_len"], st["segments"], ms, 4.0 * col.numel() * hidden / ms / 1e9))
    A.free()
