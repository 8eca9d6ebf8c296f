from pygim_b200.sparse_tensor import SparseTensor


class ToSparseTensor:
    """data.adj_t = the transposed, value-less adjacency (what the reference's drivers hand to prepare_pim_*)."""

    def __init__(self, remove_edge_index=True, **_ignored):
        self.remove_edge_index = remove_edge_index

    def __call__(self, data):
        if getattr(data, "adj_t", None) is None:
            row, col = data.edge_index
            data.adj_t = SparseTensor(row=col, col=row, value=None, sparse_sizes=(data.num_nodes, data.num_nodes))
        if self.remove_edge_index:
            data.edge_index = None
        return data


class NormalizeFeatures:
    def __call__(self, data):
        return data
