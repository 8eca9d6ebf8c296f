#!/usr/bin/env python
"""Host cost of ONE device-operand SpMM call (what bounds a sweep of small launches, e.g. the 1/8 shards at N = 8):
wall-clock per call over many back-to-back calls on a graph so small that the GPU is never the bottleneck."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygim_b200 import graphgen  # noqa: E402
from pygim_b200.backend_pim import pim_ops  # noqa: E402
from pygim_b200.backend_pim.spmm import SparseTensorCOO  # noqa: E402

pim_ops.dpu_init_ranks(1)
adj = graphgen.synthetic_adj("reddit", scale=0.002, seed=1).to("cuda")
n = adj.size(0)
A = SparseTensorCOO(adj, dtype=torch.float32, format="CSR")
A.to_pim_group(32, 1)
x = torch.ones((n, 32), device="cuda")
c = torch.empty((n, 32), device="cuda")


def per_call(fn, reps=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / reps * 1e6


print("A.mul(x, out=c)                  %6.1f us per call" % per_call(lambda: A.mul(x, out=c)))
print("pim_ops.spmm_run_dense           %6.1f us per call" % per_call(lambda: pim_ops.spmm_run_dense(A.sp_info_ptr, x, out=c)))
pim_ops.plan_set_option(A.sp_info_ptr, "l2_persist", 0)
print("  ... with l2_persist = 0        %6.1f us per call" % per_call(lambda: pim_ops.spmm_run_dense(A.sp_info_ptr, x, out=c)))
pim_ops.plan_set_option(A.sp_info_ptr, "l2_persist", 1)
print("  ... with l2_persist = 1        %6.1f us per call" % per_call(lambda: pim_ops.spmm_run_dense(A.sp_info_ptr, x, out=c)))
import ctypes as C  # noqa: E402
from pygim_b200 import _lib  # noqa: E402
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
h, xp, cp = int(A.sp_info_ptr), x.data_ptr(), c.data_ptr()
print("raw C ABI pygim_spmm_device      %6.1f us per call" % per_call(lambda: lib.pygim_spmm_device(h, xp, 32, cp, 32, C.c_void_p(st))))
pim_ops.plan_set_option(A.sp_info_ptr, "l2_persist", 0)
print("  ... with l2_persist = 0        %6.1f us per call" % per_call(lambda: lib.pygim_spmm_device(h, xp, 32, cp, 32, C.c_void_p(st))))
