import torch

from . import num_nodes  # noqa: F401


def add_remaining_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    raise NotImplementedError("not on the reference's aggregation path (GCNConv.forward never normalises)")


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    size = list(src.shape)
    size[dim] = int(dim_size if dim_size is not None else int(index.max()) + 1)
    out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.index_add_(dim, index, src)
