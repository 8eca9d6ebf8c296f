"""The callers around the path: quantise -> aggregate -> dequantise conv layers and the 2-layer stacks."""
import pytest
import torch

from pygim_b200 import graphgen
from pygim_b200.models import GCN, GIN, SAGE, symmetric_dequantize, symmetric_quantize
from pygim_b200.models.layers import aggregate


class OracleAdj:
    """Checker-side aggregation operator with the surface the conv layers use (`.dtype`, `.mul`)."""

    def __init__(self, adj, dtype, O):
        self.rowptr, self.col, _ = adj.csr()
        self.dtype, self.O = dtype, O

    def mul(self, x):
        y = self.O.spmm_csr_rowpar(self.rowptr.numpy(), self.col.numpy(), None, x.detach().cpu().numpy(), nthreads=1)
        return torch.from_numpy(y).to(x.device)


def test_quantisation_grid_matches_reference_definition():
    v = torch.tensor([[-3.0, 0.5], [1.5, 3.0]])
    for dtype, bits in ((torch.int8, 5), (torch.int16, 10), (torch.int32, 20)):
        scale, q = symmetric_quantize(v, dtype)
        assert q.dtype == dtype and float(scale) == pytest.approx(6.0 / 2 ** bits)
        assert int(q.abs().max()) == 2 ** (bits - 1)                 # |x_q| <= 16 / 512 / 2^19
    scale, q = symmetric_quantize(v, torch.float64)                  # "anything else": float on the 2^19 grid
    assert q.dtype == torch.float and float(q.abs().max()) == 2 ** 19
    assert torch.allclose(symmetric_dequantize(q, 1.0, scale), v, atol=1e-5)


@pytest.mark.parametrize("net", [GCN, GIN, SAGE])
def test_stacks_run_with_a_generic_aggregation_operator(oracle, net):
    adj = graphgen.synthetic_adj("pubmed", scale=0.05, seed=1)
    n = adj.size(0)
    torch.manual_seed(0)
    model = net(24, 16, 5, num_layers=2).eval()
    x = torch.randn(n, 24)
    with torch.no_grad():
        y = model(x, OracleAdj(adj, torch.int32, oracle))
    assert y.shape == (n, 5) and torch.isfinite(y).all()
    # int32 quantisation is fine-grained: the result tracks the float aggregation closely
    with torch.no_grad():
        y_f = model(x, OracleAdj(adj, torch.float32, oracle))
    assert torch.allclose(y, y_f, rtol=1e-2, atol=1e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("net", [GCN, GIN, SAGE])
@pytest.mark.parametrize("dtype,fmt", [(torch.int32, "COO"), (torch.int8, "COO"), (torch.float32, "CSR")])
def test_gpu_inference_matches_oracle_aggregation(gpu_backend, oracle, net, dtype, fmt):
    """Same model, same device for the dense layers; only the aggregation operator differs (CUDA plan vs
    oracle).  Quantised integers are identical on both sides, so integer paths agree bit for bit."""
    from helpers import make_args
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = graphgen.synthetic_adj("reddit", scale=0.004, seed=2)
    n = adj.size(0)
    torch.manual_seed(1)
    model = net(32, 64, 7, num_layers=2).cuda().eval()
    x = torch.randn(n, 32, device="cuda")
    A = prepare_pim_spmm(adj.to("cuda"), make_args(dtype, fmt, 64))
    with torch.no_grad():
        got = model(x, A)
        want = model(x, OracleAdj(adj, dtype, oracle))
    torch.cuda.synchronize()
    if dtype.is_floating_point:
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-4)
    else:
        assert torch.equal(got, want)
    A.free()


@pytest.mark.gpu
def test_aggregate_helper_on_gpu(gpu_backend, oracle):
    from helpers import make_args
    from pygim_b200.backend_pim.spmm import prepare_pim_spmm
    adj = graphgen.synthetic_adj("arxiv", scale=0.02, seed=3)
    x = torch.randn(adj.size(0), 32, device="cuda")
    for dtype in (torch.int8, torch.int16, torch.int32):
        A = prepare_pim_spmm(adj.to("cuda"), make_args(dtype, "COO", 32))
        got = aggregate(A, x)
        want = aggregate(OracleAdj(adj, dtype, oracle), x)
        assert torch.equal(got, want)
        A.free()
