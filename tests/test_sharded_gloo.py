"""The N > 1 path on CPU: two gloo ranks row-shard the adjacency by nnz, compute their block and
all-gather the unequal row blocks.  The per-rank compute is injected (the CUDA plan cannot run here):
the oracle stands in as the local operator, so this checks partitioning + collective + assembly."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _OracleLocal:
    """Checker-side local operator: same mul(B, out=) surface as SparseTensorCOO."""

    def __init__(self, adj, args):
        from oracle import oracle as O
        self.O = O
        self.rowptr, self.col, _ = adj.csr()

    def mul(self, B, out=None):
        y = self.O.spmm_csr_rowpar(self.rowptr.numpy(), self.col.numpy(), None, B.numpy(), nthreads=1)
        out.copy_(torch.from_numpy(y))
        return out


class _OracleLocalOut(_OracleLocal):
    """Same checker, `mul(B)` returning a new tensor (the surface ColumnShardedSpMM uses)."""

    def mul(self, B, out=None):
        B = B.contiguous()
        return super().mul(B, out=torch.empty(self.rowptr.numel() - 1, B.size(1), dtype=B.dtype))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pygim_b200 import graphgen
        from pygim_b200.sharded import ShardedSpMM
        adj = graphgen.synthetic_adj("reddit", scale=0.01, seed=5)         # skewed rows => unequal blocks
        n = adj.size(0)
        x = graphgen.reference_features(n, 16, torch.float32, seed=1)
        args = types.SimpleNamespace(data_type=torch.float32, sp_format="CSR", hidden_size=16, sp_parts=1, ds_parts=1)
        op = ShardedSpMM(adj, args, make_local=_OracleLocal)
        out = op.mul(x)
        op2 = ShardedSpMM(adj, args, make_local=_OracleLocal, chunks=3)       # overlapped sub-block schedule
        out2 = op2.mul(x)
        from pygim_b200.sharded import ColumnShardedSpMM
        op3 = ColumnShardedSpMM(adj, args, make_local=_OracleLocalOut)        # feature-column sharding
        out3 = op3.mul(x)
        assert op3.col_splits[-1] == 16 and all(c % 4 == 0 for c in op3.col_splits)
        full = _OracleLocal(adj, args).mul(x, out=torch.empty(n, 16))
        nnz_local = int(op.local_adj.nnz())
        ret[rank] = (bool(torch.equal(out, full)) and bool(torch.equal(out2, full)) and bool(torch.equal(out3, full)),
                     op.splits, nnz_local, adj.nnz())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_spmm_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, splits, nnz_local, nnz = ret[r]
        assert ok, "rank %d assembled a wrong result" % r
        assert splits == ret[0][1] and len(set(np.diff(splits))) > 1       # same cuts everywhere, unequal rows
    shard_nnz = [ret[r][2] for r in range(world)]
    assert sum(shard_nnz) == ret[0][3]
    assert max(shard_nnz) <= 1.2 * ret[0][3] / world                       # balanced by nnz, not by rows
