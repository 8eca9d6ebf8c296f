"""The reference's callers around the path.

* In the build container (where /root/reference exists) the reference's OWN classes - models/models.py with
  pyg_gcn_conv.py / pyg_gin_conv.py / pyg_sage_conv.py / quantize.py, imported UNCHANGED through tests/shims - are run
  next to pygim_b200.models with the same weights and the same aggregation operator: outputs must be identical.
* Everywhere (the GPU box has no /root/reference) the golden outputs those classes produced
  (tests/golden/reference_models.npz, generator: tests/golden/make_model_golden.py) pin pygim_b200.models on the CPU
  and - marked gpu - on the B200 backend, where the quantised aggregation itself is bit-exact and only the dense
  torch layers differ between CPU and GPU arithmetic."""
import os
import sys

import numpy as np
import pytest
import torch

from pygim_b200 import graphgen
from pygim_b200.models import GCN, GIN, SAGE

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_models.npz")
REF = os.environ.get("PYGIM_REFERENCE_ROOT", "/root/reference")
OURS = {"gcn": GCN, "gin": GIN, "sage": SAGE}
TAGS = {"i32": torch.int32, "i8": torch.int8, "f32": torch.float32}


def _our_key(name, key):
    """State-dict key of pygim_b200.models for a key of the reference's model (only GIN's MLP is laid out
    differently: torch_geometric's MLP(lins, norms) vs nn.Sequential(Linear, BatchNorm, ReLU, Linear))."""
    if name == "gin":
        key = key.replace(".nn.lins.0.", ".nn.0.").replace(".nn.norms.0.", ".nn.1.").replace(".nn.lins.1.", ".nn.3.")
    return key


def _load_into(model, name, weights):
    sd = model.state_dict()
    used = set()
    for key, val in weights.items():
        k = _our_key(name, key)
        assert k in sd, (name, key, k)
        assert tuple(sd[k].shape) == tuple(val.shape), (k, sd[k].shape, val.shape)
        sd[k].copy_(torch.as_tensor(val))
        used.add(k)
    assert used == set(sd), set(sd) - used
    return model


def _golden():
    blob = np.load(GOLDEN)
    scale, seed, n_in, hid, n_out, n = blob["meta"]
    adj = graphgen.synthetic_adj("reddit", scale=float(scale), seed=int(seed))
    assert adj.size(0) == int(n)
    weights = {name: {k.split("/w/")[1]: blob[k] for k in blob.files if k.startswith(name + "/w/")} for name in OURS}
    return blob, adj, (int(n_in), int(hid), int(n_out)), weights


class OracleAdj:
    def __init__(self, adj, dtype, O):
        self.rowptr, self.col, _ = adj.csr()
        self.dtype, self.O = dtype, O

    def mul(self, x):
        y = self.O.spmm_csr_rowpar(self.rowptr.numpy(), self.col.numpy(), None, x.detach().cpu().numpy(), nthreads=1)
        return torch.from_numpy(y).to(x.device)


@pytest.mark.parametrize("name", ["gcn", "gin", "sage"])
def test_our_models_reproduce_the_reference_classes_golden_outputs(oracle, name):
    blob, adj, (n_in, hid, n_out), weights = _golden()
    model = _load_into(OURS[name](n_in, hid, n_out, num_layers=2).eval(), name, weights[name])
    x = torch.from_numpy(blob["x"])
    for tag, dtype in TAGS.items():
        with torch.no_grad():
            y = model(x, OracleAdj(adj, dtype, oracle))
        assert torch.equal(y, torch.from_numpy(blob["%s/y/%s" % (name, tag)])), (name, tag)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="needs the reference tree")
@pytest.mark.parametrize("name", ["gcn", "gin", "sage"])
def test_reference_classes_run_unchanged_and_agree_with_ours(oracle, name):
    sys.path.insert(0, os.path.join(HERE, "shims"))
    import run_reference
    run_reference.prepare()
    from models.models import GCN as RGCN, GIN as RGIN, SAGE as RSAGE          # the reference's own classes
    ref_cls = {"gcn": RGCN, "gin": RGIN, "sage": RSAGE}[name]
    assert ref_cls.__module__ == "models.models" and REF in sys.modules["models.models"].__file__
    adj = graphgen.synthetic_adj("pubmed", scale=0.05, seed=3)
    n = adj.size(0)
    torch.manual_seed(5)
    ref = ref_cls(20, 32, 6, num_layers=3).eval()
    ours = _load_into(OURS[name](20, 32, 6, num_layers=3).eval(), name,
                      {k: v.numpy() for k, v in ref.state_dict().items()})
    x = torch.randn(n, 20)
    for dtype in (torch.int32, torch.int16, torch.int8, torch.float32):
        op = OracleAdj(adj, dtype, oracle)
        with torch.no_grad():
            assert torch.equal(ref(x, op, None), ours(x, op)), (name, dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gcn", "gin", "sage"])
@pytest.mark.parametrize("tag,fmt", [("i32", "COO"), ("i8", "COO"), ("i32", "CSR"), ("f32", "CSR")])
def test_gpu_backend_reproduces_the_reference_classes_golden_outputs(gpu_backend, name, tag, fmt):
    """pygim_b200.models on the GPU, aggregation through libbackend_pim.so (fused epilogue on and off), against what
    the reference's classes produced.  Integer aggregation is exact; the dense layers run in GPU float arithmetic."""
    from helpers import make_args
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    from pygim_b200.models import layers
    blob, adj, (n_in, hid, n_out), weights = _golden()
    model = _load_into(OURS[name](n_in, hid, n_out, num_layers=2).eval(), name, weights[name]).cuda()
    x = torch.from_numpy(blob["x"]).cuda()
    want = torch.from_numpy(blob["%s/y/%s" % (name, tag)])
    A = prepare_pim_spmm(adj.to("cuda"), make_args(TAGS[tag], fmt, hid))
    outs = []
    for fused in (True, False):
        layers.FUSED_EPILOGUE = fused
        try:
            with torch.no_grad():
                outs.append(model(x, A).cpu())
        finally:
            layers.FUSED_EPILOGUE = True
    A.free()
    tol = dict(rtol=2e-3, atol=2e-3) if tag != "i8" else dict(rtol=5e-2, atol=5e-2)   # int8: 5-bit grid flips on GPU/CPU matmul noise
    for y in outs:
        assert torch.isfinite(y).all()
        assert torch.allclose(y, want, **tol), (name, tag, float((y - want).abs().max()))
    assert torch.equal(outs[0], outs[1]), "fused and unfused epilogues must give the same bits"
