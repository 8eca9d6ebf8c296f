"""Generates tests/golden/reference_models.npz: the reference's OWN GCN / GIN / SAGE classes
(/root/reference/models/models.py with pyg_*_conv.py and quantize.py, imported unchanged through tests/shims) run
on a small Reddit-like graph with the aggregation operator `adj_t.mul` served by the CPU oracle, for the quantised
INT32 / INT8 paths and the FLT32 path.  Inputs, weights and outputs travel to the GPU box, where
tests/test_reference_models.py runs pygim_b200.models on the B200 backend against them.

    python tests/golden/make_model_golden.py        (needs /root/reference; run in the build container)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
import run_reference  # noqa: E402

run_reference.prepare()

from models.models import GCN, GIN, SAGE  # noqa: E402  (the reference's classes)
from oracle import oracle as O  # noqa: E402
from pygim_b200 import graphgen  # noqa: E402

SCALE, SEED, IN, HID, OUT = 0.004, 2, 32, 64, 7


class OracleAdj:
    """What prepare_pim_spmm returns, as far as the conv layers can tell: `.dtype` and `.mul`."""

    def __init__(self, adj, dtype):
        self.rowptr, self.col, _ = adj.csr()
        self.dtype = dtype

    def mul(self, x):
        y = O.spmm_csr_rowpar(self.rowptr.numpy(), self.col.numpy(), None, x.detach().numpy(), nthreads=1)
        return torch.from_numpy(y)


def main():
    O.build()
    adj = graphgen.synthetic_adj("reddit", scale=SCALE, seed=SEED)
    n = adj.size(0)
    out = {"meta": np.array([SCALE, SEED, IN, HID, OUT, n], dtype=np.float64)}
    torch.manual_seed(1)
    x = torch.randn(n, IN)
    out["x"] = x.numpy()
    for name, net in (("gcn", GCN), ("gin", GIN), ("sage", SAGE)):
        torch.manual_seed(10)
        model = net(IN, HID, OUT, num_layers=2).eval()
        # non-trivial BatchNorm statistics and GIN eps, so nothing is hidden by defaults
        g = torch.Generator().manual_seed(11)
        for key, val in model.state_dict().items():
            if key.endswith("running_mean"):
                val.copy_(torch.randn(val.shape, generator=g) * 0.1)
            elif key.endswith("running_var"):
                val.copy_(torch.rand(val.shape, generator=g) + 0.5)
            elif key.endswith(".eps"):
                val.fill_(0.25)
        for key, val in model.state_dict().items():
            out["%s/w/%s" % (name, key)] = val.numpy()
        for tag, dtype in (("i32", torch.int32), ("i8", torch.int8), ("f32", torch.float32)):
            with torch.no_grad():
                y = model(x, OracleAdj(adj, dtype), None)
            out["%s/y/%s" % (name, tag)] = y.numpy()
    path = os.path.join(HERE, "reference_models.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
