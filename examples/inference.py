#!/usr/bin/env python
"""End-to-end 2-layer GCN / GIN / SAGE inference with GPU aggregation (the reference's inference.py, same
flags and `[DATA]` output protocol), on synthetic graphs of the named dataset SHAPES (no dataset download is
possible here).  `--version cpu` is not offered: there is no CPU fallback in this repository; the reference's CPU
aggregation is timed next to the GPU by `bench.py --workload inference` (its cpu_baseline leg)."""
import argparse
import datetime
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pygim_b200 import graphgen  # noqa: E402
from pygim_b200.backend_pim import pim_ops  # noqa: E402,F401  (registers torch.ops.pim_ops)
from pygim_b200.backend_pim.grande import prepare_pim_spmm_grande  # noqa: E402
from pygim_b200.backend_pim.spmm import TORCH_TYPES, prepare_pim_spmm  # noqa: E402
from pygim_b200.backend_pim.spmv import prepare_pim_spmv  # noqa: E402
from pygim_b200.models import GCN, GIN, SAGE  # noqa: E402

SHAPE_OF = {"Reddit": "reddit", "ogbn-arxiv": "arxiv", "ogbn-products": "products", "PubMed": "pubmed"}
CLASSES = {"Reddit": 41, "ogbn-arxiv": 40, "ogbn-products": 47, "PubMed": 3}
FEATURES = {"Reddit": 602, "ogbn-arxiv": 128, "ogbn-products": 100, "PubMed": 500}


@torch.no_grad()
def test(args, model, data):
    model.eval()
    if data["x"].is_cuda:
        torch.cuda.synchronize()
    st = datetime.datetime.now()
    y_pred = model(data["x"], data["adj_t"], data["edge_attr"])
    if y_pred.is_cuda:
        torch.cuda.synchronize()
    end = datetime.datetime.now()
    print("[DATA]infer_time(ms): ", (end - st).total_seconds() * 1000)
    return y_pred


def get_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", type=str, default="PubMed", choices=sorted(SHAPE_OF))
    p.add_argument("--datadir", type=str, default="./data")
    p.add_argument("--lr", type=float, default=0.01)
    p.add_argument("--version", type=str, default="spmm", choices=["spmm", "grande", "spmv", "cpu"])
    p.add_argument("--tune", type=bool, default=True)
    p.add_argument("--lib_path", type=str, default=None)
    p.add_argument("--model", type=str, default="gcn", choices=["gcn", "gin", "sage"])
    p.add_argument("--num_layers", type=int, default=2)
    p.add_argument("--hidden_size", type=int, default=128)
    p.add_argument("--data_type", type=str, default="INT32", choices=sorted(TORCH_TYPES))
    p.add_argument("--sp_format", type=str, default="COO", choices=["CSR", "COO"])
    p.add_argument("--sp_parts", type=int, default=1)
    p.add_argument("--ds_parts", type=int, default=1)
    p.add_argument("--repeat", type=int, default=3)
    p.add_argument("--nr_dpus", type=int, default=0)
    p.add_argument("--device", type=str, default="cuda" if torch.cuda.is_available() else "cpu",
                   help="where x and the dense layers live (the reference hard-codes 'cpu')")
    p.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic graph (tests)")
    args = p.parse_args(argv)
    print(args, flush=True)
    args.data_type = TORCH_TYPES[args.data_type]
    return args


def main(args):
    torch.manual_seed(0)
    adj = graphgen.synthetic_adj(SHAPE_OF[args.dataset], seed=0, scale=args.scale)
    n = adj.size(0)
    x = torch.randn(n, FEATURES[args.dataset])
    data = {"x": x.to(args.device), "edge_attr": None}
    if args.version != "cpu":
        if args.lib_path:
            torch.ops.load_library(args.lib_path)
        if args.nr_dpus == 0:
            if args.version == "grande":
                dpus_per_rank = torch.ops.pim_ops.dpu_init_ranks(args.sp_parts)
            else:
                torch.ops.pim_ops.dpu_init_ranks(args.sp_parts * args.ds_parts)
        else:
            torch.ops.pim_ops.dpu_init_dpus(args.nr_dpus)
        dev_adj = adj.to(args.device) if args.device != "cpu" else adj
        if args.version == "spmm":
            data["adj_t"] = prepare_pim_spmm(dev_adj, args)
        elif args.version == "spmv":
            data["adj_t"] = prepare_pim_spmv(dev_adj, args)
        else:
            data["adj_t"] = prepare_pim_spmm_grande(dev_adj, args, dpus_per_rank)
    else:
        raise NotImplementedError("no CPU fallback: `bench.py --workload inference` times the reference's CPU path")
    net = {"gcn": GCN, "gin": GIN, "sage": SAGE}[args.model]
    model = net(x.size(-1), args.hidden_size, CLASSES[args.dataset], args.num_layers).to(args.device)
    out = None
    for i in range(args.repeat):
        print("-------------------- Model={} nrl={} Repeat={}--------------------".format(
            args.model, args.num_layers, i), flush=True)
        out = test(args, model, data)
    if args.version != "cpu":
        torch.ops.pim_ops.dpu_release()
    return out


if __name__ == "__main__":
    main(get_args())
