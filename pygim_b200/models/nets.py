"""GCN / SAGE / GIN stacks of the reference (models/models.py:12-131):
Linear -> BN -> ReLU -> dropout -> (conv -> BN -> ReLU -> dropout) x L -> Linear."""
from __future__ import annotations

import torch.nn.functional as F
from torch import nn

from .layers import GCNConv, GINConv, SAGEConv


def _gin_mlp(width: int) -> nn.Module:
    # torch_geometric.nn.MLP([w, w, w]) with its defaults: Linear, BatchNorm, ReLU, plain last Linear
    return nn.Sequential(nn.Linear(width, width), nn.BatchNorm1d(width), nn.ReLU(), nn.Linear(width, width))


class _Stack(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels, num_layers=2, dropout=0.5):
        super().__init__()
        self.ln1 = nn.Linear(in_channels, hidden_channels)
        self.bn0 = nn.BatchNorm1d(hidden_channels)
        self.convs = nn.ModuleList(self.make_conv(hidden_channels) for _ in range(num_layers))
        self.bns = nn.ModuleList(nn.BatchNorm1d(hidden_channels) for _ in range(num_layers))
        self.ln2 = nn.Linear(hidden_channels, out_channels)
        self.dropout = dropout

    def make_conv(self, width):
        raise NotImplementedError

    def forward(self, x, adj_t, edge_attr=None):
        x = F.dropout(F.relu(self.bn0(self.ln1(x))), p=self.dropout, training=self.training)
        for conv, bn in zip(self.convs, self.bns):
            x = F.dropout(F.relu(bn(conv(x, adj_t))), p=self.dropout, training=self.training)
        return self.ln2(x)


class GCN(_Stack):
    def make_conv(self, width):
        return GCNConv(width, width)


class SAGE(_Stack):
    def make_conv(self, width):
        return SAGEConv(width, width)


class GIN(_Stack):
    def make_conv(self, width):
        return GINConv(_gin_mlp(width))
