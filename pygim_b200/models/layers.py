"""GCN / GIN / SAGE convolution layers whose aggregation is `adj_t.mul(x_q)` on a backend_pim
SparseTensorCOO (or any object with `.dtype` and `.mul`), wrapped in quantise / dequantise exactly like the
reference's overridden `message_and_aggregate` (models/pyg_gcn_conv.py:130-137, pyg_gin_conv.py:93-101,
pyg_sage_conv.py:147-155).  Like the reference's GCNConv.forward (:116-125) there is no normalisation and no
self-loop insertion: forward = lin -> aggregate -> + bias."""
from __future__ import annotations

import math

import torch
from torch import nn

from .quantize import symmetric_dequantize, symmetric_quantize


FUSED_EPILOGUE = True      # module switch (bench.py times both settings)


def aggregate(adj_t, x: torch.Tensor, residual: torch.Tensor = None, coeff: float = 1.0) -> torch.Tensor:
    """quantise -> sparse x dense -> dequantise (+ coeff * residual).  `adj_t.dtype` selects the quantisation grid.
    On the GPU the elementwise passes are fused into the aggregation kernels (SparseTensorCOO.mul_fused: same bits,
    ~6 fewer N x H passes per layer); otherwise they are the reference's torch expressions."""
    fused = getattr(adj_t, "mul_fused", None) if FUSED_EPILOGUE else None
    if fused is not None and x.is_cuda:
        out = fused(x, residual, coeff)
        if out is not None:
            return out
    scale, x_q = symmetric_quantize(x, dtype=adj_t.dtype)
    out_q = adj_t.mul(x_q)
    out = symmetric_dequantize(out_q, 1.0, scale)
    return out if residual is None else out + coeff * residual


class GCNConv(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, **_ignored):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        bound = math.sqrt(6.0 / (self.in_channels + self.out_channels))      # glorot
        nn.init.uniform_(self.lin.weight, -bound, bound)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, adj_t, edge_weight=None):
        out = aggregate(adj_t, self.lin(x))
        return out if self.bias is None else out + self.bias


class GINConv(nn.Module):
    """h( A x + (1 + eps) x )"""

    def __init__(self, net: nn.Module, eps: float = 0.0, train_eps: bool = False, **_ignored):
        super().__init__()
        self.nn = net
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([eps]))
        else:
            self.register_buffer("eps", torch.tensor([eps]))

    def forward(self, x, adj_t, size=None):
        if FUSED_EPILOGUE and x.is_cuda and getattr(adj_t, "mul_fused", None) is not None:
            # (1 + eps) as the float32 value torch computes; read once per eps version (no per-call host sync)
            ver = self.eps._version
            if getattr(self, "_coeff_ver", None) != ver:
                self._coeff, self._coeff_ver = float((1 + self.eps).item()), ver
            return self.nn(aggregate(adj_t, x, residual=x, coeff=self._coeff))
        return self.nn(aggregate(adj_t, x) + (1 + self.eps) * x)


class SAGEConv(nn.Module):
    """lin_l( A x ) + lin_r( x )  (aggr = "add", the reference's default, pyg_sage_conv.py:72)"""

    def __init__(self, in_channels: int, out_channels: int, root_weight: bool = True, bias: bool = True,
                 normalize: bool = False, **_ignored):
        super().__init__()
        self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False) if root_weight else None
        self.normalize = normalize

    def forward(self, x, adj_t, size=None):
        out = self.lin_l(aggregate(adj_t, x))
        if self.lin_r is not None:
            out = out + self.lin_r(x)
        return nn.functional.normalize(out, p=2.0, dim=-1) if self.normalize else out
