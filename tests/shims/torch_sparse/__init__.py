"""Stand-in for rusty1s/pytorch_sparse (unpinned in the reference's Libs/install_libs.sh:13): the data structure is
pygim_b200's, `matmul` is the CPU oracle's row-parallel CSR SpMM (the algorithm class of torch_sparse's spmm_sum)."""
import numpy as np
import torch

from pygim_b200.sparse_tensor import SparseTensor  # noqa: F401


def matmul(src, other, reduce: str = "sum"):
    assert reduce in ("sum", "add")
    from oracle import oracle as O
    rowptr, col, value = src.csr()
    x = other.detach().cpu().contiguous()
    v = None if value is None else value.to(x.dtype).cpu().numpy()
    y = O.spmm_csr_rowpar(rowptr.cpu().numpy().astype(np.int32), col.cpu().numpy().astype(np.int32), v, x.numpy(),
                          nthreads=O.max_threads())
    return torch.from_numpy(y).to(other.device)
