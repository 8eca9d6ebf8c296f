import torch

from ..aggr import SumAggregation


class MessagePassing(torch.nn.Module):
    """The part of MessagePassing the reference's conv layers rely on: the constructor's aggregation bookkeeping.
    Their forward() calls message_and_aggregate directly (pyg_gcn_conv.py:121, pyg_gin_conv.py:80,
    pyg_sage_conv.py:131), so propagate() only has to route there."""

    def __init__(self, aggr="add", *, aggr_kwargs=None, flow="source_to_target", node_dim=-2, **_ignored):
        super().__init__()
        self.aggr = aggr
        self.aggr_module = SumAggregation()
        self.flow, self.node_dim, self.fuse = flow, node_dim, True

    def propagate(self, edge_index, size=None, **kwargs):
        return self.message_and_aggregate(edge_index, **kwargs)
