"""Dataset classes under the names the reference's drivers import (spmm_test.py:8, inference.py:8).  No dataset can
be downloaded here, so each serves the synthetic graph of the same SHAPE (pygim_b200.graphgen), scaled by the
PYGIM_SHIM_SCALE environment variable (default 0.05) to keep driver tests quick."""
import os

import torch

from pygim_b200 import graphgen


class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if hasattr(v, "to") and not isinstance(v, (int, float)):
                try:
                    self.__dict__[k] = v.to(device)
                except TypeError:
                    pass
        return self


class _Synthetic:
    shape, num_features, num_classes = "pubmed", 500, 3

    def __init__(self, root=None, name=None, transform=None, **_ignored):
        scale = float(os.environ.get("PYGIM_SHIM_SCALE", "0.05"))
        adj = graphgen.synthetic_adj(self.shape, seed=0, scale=scale)
        n = adj.size(0)
        g = torch.Generator().manual_seed(0)
        mask = torch.zeros(n, dtype=torch.bool)
        mask[torch.randperm(n, generator=g)[: max(1, n // 5)]] = True
        row, col, _ = adj.coo()
        self._data = Data(x=torch.randn(n, self.num_features, generator=g), y=torch.randint(0, self.num_classes, (n,), generator=g),
                          edge_index=torch.stack([col, row]), edge_attr=None, adj_t=adj, num_nodes=n,
                          train_mask=~mask, val_mask=mask, test_mask=mask)
        self.transform = transform

    def __getitem__(self, idx):
        return self.transform(self._data) if self.transform else self._data

    def __len__(self):
        return 1


class Planetoid(_Synthetic):
    shape, num_features, num_classes = "pubmed", 500, 3


class Reddit(_Synthetic):
    shape, num_features, num_classes = "reddit", 602, 41


class AmazonProducts(_Synthetic):
    shape, num_features, num_classes = "products", 200, 107
