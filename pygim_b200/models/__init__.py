"""The callers either side of the aggregation path: PyGim's quantised conv layers and 2-layer GNN stacks
(models/quantize.py, models/pyg_{gcn,gin,sage}_conv.py, models/models.py of the reference), restated without
torch_geometric so that the end-to-end inference configuration can run on the GPU box.  The dense Linear /
BatchNorm work stays plain torch; only `adj_t.mul(x_q)` goes through libbackend_pim.so."""
from .layers import GCNConv, GINConv, SAGEConv
from .nets import GCN, GIN, SAGE
from .quantize import symmetric_dequantize, symmetric_quantize

__all__ = ["GCNConv", "GINConv", "SAGEConv", "GCN", "GIN", "SAGE", "symmetric_quantize", "symmetric_dequantize"]
