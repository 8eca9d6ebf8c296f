#!/usr/bin/env python
"""SpMM micro-benchmark driver with the reference's flags and `[DATA]` stdout protocol (spmm_test.py), on
synthetic graphs of the named dataset shapes.  Prints `[DATA]pim_time_spmm(ms)` per repeat and, for host
operands, the five phase timers the reference's `spmm_pim_*` print (`[DATA]load_sparse_time` ...), so the
reference's `Experiment.parse_result` (utils/experiment.py:468-491) can consume the log unchanged.
`--version cpu` is not offered here (no CPU fallback in this repository; see bench.py --impl reference)."""
import argparse
import datetime
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pygim_b200 import graphgen  # noqa: E402
from pygim_b200.backend_pim import pim_ops  # noqa: E402
from pygim_b200.backend_pim.grande import prepare_pim_spmm_grande  # noqa: E402
from pygim_b200.backend_pim.spmm import TORCH_TYPES, prepare_pim_spmm  # noqa: E402
from pygim_b200.backend_pim.spmv import prepare_pim_spmv  # noqa: E402

SHAPE_OF = {"Reddit": "reddit", "ogbn-arxiv": "arxiv", "ogbn-products": "products", "PubMed": "pubmed"}


def spmm_test(adj, pim_adj_t, data_x, args):
    print("{} Dataset Info: Node({}), Edge({})".format(args.dataset, adj.size(1), adj.nnz()))
    data_x_pim = data_x.type(args.data_type)
    if data_x_pim.is_cuda:
        torch.cuda.synchronize()
    start_pim = datetime.datetime.now()
    res_lib = pim_adj_t.mul(data_x_pim)
    if res_lib.is_cuda:
        torch.cuda.synchronize()
    end_pim = datetime.datetime.now()
    print("[DATA]pim_time_spmm(ms): ", (end_pim - start_pim).total_seconds() * 1000, flush=True)
    if not data_x_pim.is_cuda and args.version != "spmv":
        for key, value in pim_ops.last_timers(pim_adj_t.sp_info_ptr).items():
            print("[DATA]%s: %f" % (key, value), flush=True)
    return res_lib


def get_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", type=str, default="PubMed", choices=sorted(SHAPE_OF))
    p.add_argument("--datadir", type=str, default="./data")
    p.add_argument("--lr", type=float, default=0.01)
    p.add_argument("--version", type=str, default="spmm", choices=["spmm", "grande", "spmv"])
    p.add_argument("--tune", type=bool, default=True)
    p.add_argument("--lib_path", type=str, default=None)
    p.add_argument("--hidden_size", type=int, default=256)
    p.add_argument("--data_type", type=str, default="INT32", choices=sorted(TORCH_TYPES))
    p.add_argument("--sp_format", type=str, default="COO", choices=["CSR", "COO"])
    p.add_argument("--sp_parts", type=int, default=1)
    p.add_argument("--ds_parts", type=int, default=1)
    p.add_argument("--repeat", type=int, default=3)
    p.add_argument("--nr_dpus", type=int, default=0)
    p.add_argument("--device", type=str, default="cpu", help="where data.x lives; the reference hard-codes 'cpu'")
    p.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic graph")
    args = p.parse_args(argv)
    print(args, flush=True)
    args.data_type = TORCH_TYPES[args.data_type]
    return args


def main(args):
    adj = graphgen.synthetic_adj(SHAPE_OF[args.dataset], seed=0, scale=args.scale)
    x = graphgen.reference_features(adj.size(1), args.hidden_size, args.data_type).to(args.device)   # spmm_test.py:70
    if args.lib_path:
        torch.ops.load_library(args.lib_path)
    if args.nr_dpus == 0:
        dpus_per_rank = torch.ops.pim_ops.dpu_init_ranks(args.sp_parts if args.version == "grande"
                                                         else args.sp_parts * args.ds_parts)
    else:
        torch.ops.pim_ops.dpu_init_dpus(args.nr_dpus)
        dpus_per_rank = [1] * args.sp_parts
    if args.version == "spmm":
        pim_adj_t = prepare_pim_spmm(adj, args)
    elif args.version == "spmv":
        pim_adj_t = prepare_pim_spmv(adj, args)
    else:
        pim_adj_t = prepare_pim_spmm_grande(adj, args, dpus_per_rank)
    res = None
    for i in range(args.repeat):
        print("-------------------- Model=spmm_test Repeat={}--------------------".format(i), flush=True)
        res = spmm_test(adj, pim_adj_t, x, args)
    torch.ops.pim_ops.dpu_release()
    return res


if __name__ == "__main__":
    main(get_args())
