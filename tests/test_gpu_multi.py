"""Two NCCL ranks on two GPUs: row-sharded SpMM + all-gather equals the oracle (skipped on 1-GPU boxes)."""
import os
import socket
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import oracle as O
        from pygim_b200 import graphgen
        from pygim_b200.backend_pim import pim_ops
        from pygim_b200.sharded import ShardedSpMM
        pim_ops.dpu_init_ranks(1)
        adj = graphgen.synthetic_adj("reddit", scale=0.01, seed=5)
        n = adj.size(0)
        rowptr, col, _ = adj.csr()
        ok = True
        for dtype, hidden in ((torch.float32, 64), (torch.int32, 32)):
            x = graphgen.reference_features(n, hidden, dtype, seed=1)
            args = types.SimpleNamespace(data_type=dtype, sp_format="CSR", hidden_size=hidden, sp_parts=1, ds_parts=1)
            want = O.spmm_csr_rowpar(rowptr.numpy(), col.numpy(), None, x.numpy())
            for kw in (dict(chunks=1), dict(chunks=3), dict(fused=True, use_multicast=False), dict(fused=True),
                       dict(fused=True, sync="barrier"), dict(fused=True, sync="flags", use_multicast=False)):
                op = ShardedSpMM(adj.to("cuda"), args, **kw)
                for _ in range(5):            # more calls than rotating buffers
                    out = op.mul(x.cuda())
                torch.cuda.synchronize()
                good = bool(np.array_equal(out.cpu().numpy(), want))
                if not good:
                    print("rank", rank, "mismatch", dtype, hidden, kw, flush=True)
                ok = ok and good
                op.free()
        from pygim_b200.sharded import ColumnShardedSpMM
        x = graphgen.reference_features(n, 48, torch.float32, seed=2)
        args = types.SimpleNamespace(data_type=torch.float32, sp_format="CSR", hidden_size=48, sp_parts=1, ds_parts=1)
        cop = ColumnShardedSpMM(adj.to("cuda"), args)
        out = cop.mul(x.cuda())
        torch.cuda.synchronize()
        ok = ok and bool(np.array_equal(out.cpu().numpy(), O.spmm_csr_rowpar(rowptr.numpy(), col.numpy(), None, x.numpy())))
        cop.free()
        ret[rank] = ok
        pim_ops.dpu_release()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_sharded_spmm_nccl():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
