import torch

from . import aggr, conv, dense, inits  # noqa: F401
from .dense.linear import Linear


class MLP(torch.nn.Module):
    """torch_geometric.nn.MLP(channel_list) with its defaults: Linear -> BatchNorm -> ReLU per hidden layer, plain
    last Linear (norm='batch_norm', act='relu', plain_last=True, dropout=0)."""

    def __init__(self, channel_list, **_ignored):
        super().__init__()
        self.lins = torch.nn.ModuleList(Linear(a, b) for a, b in zip(channel_list[:-1], channel_list[1:]))
        self.norms = torch.nn.ModuleList(torch.nn.BatchNorm1d(c) for c in channel_list[1:-1])

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()
        for norm in self.norms:
            norm.reset_parameters()

    def forward(self, x):
        for lin, norm in zip(self.lins[:-1], self.norms):
            x = torch.relu(norm(lin(x)))
        return self.lins[-1](x)
