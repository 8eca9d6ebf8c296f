#!/usr/bin/env python
"""CPU-baseline protocol of BASELINE.md section 3 / SURVEY.md 8(d), run on the GPU box's HOST cores.

For each BASELINE.json config: the reference's `--version=cpu` path (`torch_sparse.matmul`, restated as the oracle's
row-parallel CSR SpMM - torch_sparse is not installable) timed the way spmm_test.py:24-27,130-132 does - one call,
datetime, NO warm-up (first-call) - plus best-of-N, on all host threads; and `torch.sparse.mm` on the same CSR as a
second CPU reference.  Prints one JSON object (gpurun_out/cpu_protocol.json is what BASELINE.md's table is filled from).
The graphs are generated on the GPU when one is present (seconds instead of minutes) and moved to the host."""
import datetime
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from pygim_b200 import graphgen  # noqa: E402


def timed_calls(fn, n=3):
    out = []
    for _ in range(n):
        t0 = datetime.datetime.now()
        fn()
        out.append((datetime.datetime.now() - t0).total_seconds() * 1e3)
    return out


def main():
    O.build()
    native = O.build_native()
    clib = O.lib(native) if native else O.lib()
    threads = O.max_threads()
    torch.set_num_threads(threads)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    rec = {"cores": threads, "torch_threads": torch.get_num_threads(), "native_build": bool(native), "rows": []}
    for shape, dtype, sweep in (("arxiv", torch.float32, [32]), ("reddit", torch.float32, [16, 32, 64, 128]),
                                ("reddit", torch.int8, [32]), ("reddit", torch.int32, [32]),
                                ("products", torch.float32, [16, 32, 64, 128])):
        n, nnz, max_deg = graphgen.SHAPES[shape]
        rowptr, col = graphgen.synthetic_csr(n, nnz, max_deg, seed=0, device=dev)
        rp, cl = rowptr.cpu().numpy().astype(np.int32), col.cpu().numpy().astype(np.int32)
        del rowptr, col
        csr = None
        if dtype == torch.float32:
            csr = torch.sparse_csr_tensor(torch.from_numpy(rp), torch.from_numpy(cl), torch.ones(nnz), size=(n, n))
        for h in sweep:
            x = graphgen.reference_features(n, h, dtype, seed=h)
            xn = x.numpy()
            out = np.empty((n, h), dtype=xn.dtype)
            calls = timed_calls(lambda: O.spmm_csr_rowpar(rp, cl, None, xn, nthreads=threads, out=out, clib=clib), 4)
            row = {"shape": shape, "dtype": str(dtype).replace("torch.", ""), "hidden": h, "nnz": nnz,
                   "oracle_first_call_ms": calls[0], "oracle_best_ms": min(calls),
                   "oracle_best_gflops": 2.0 * nnz * h / min(calls) / 1e6}
            if csr is not None:
                t = timed_calls(lambda: torch.sparse.mm(csr, x), 3)
                row.update(torch_sparse_mm_first_ms=t[0], torch_sparse_mm_best_ms=min(t),
                           torch_sparse_mm_best_gflops=2.0 * nnz * h / min(t) / 1e6)
            rec["rows"].append(row)
            print(json.dumps(row), file=sys.stderr, flush=True)
        del csr
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
