#!/usr/bin/env python
"""bench.py - the aggregation hot path (sparse adjacency x dense features) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Reddit-shaped synthetic graph (232,965 nodes, 114,615,892 edges,
value-less adjacency => ones, features = randint(-8, 4) as spmm_test.py:70), FLT32 CSR SpMM, hidden-size
sweep 16/32/64/128.  One STEP = one pass of the hot path over the sweep (four SpMMs).

metric  : SpMM GFLOP/s = sum_H 2*nnz*H / time (whole job, all GPUs), inputs resident in HBM.
e2e     : the same step through the plugin with HOST (pinned) operands, H2D of B and D2H of C inside the timed
          region: at N = 1 the sweep is one call of the batch host entry point (pygim_spmm_run_many_host: one
          upload / compute / download pipeline), at N > 1 every rank uploads its row block of B, the blocks are
          all-gathered over NVLink and the rank's rows of C go back.
roofline: HBM bound for the dominant kernel instantiation (the 64-column-tile kernel: the H = 64 launch and the
          two H = 128 tile launches): algorithmic bytes 4(N+1) + 4 nnz + s nnz + s N H + s N H / its
          CUDA-event duration, against MEASURED_PEAKS.json's hbm_gbs; `gather` = the L2 -> SM rate against the
          measured ceiling of random row gathers.
N > 1   : the adjacency is row-sharded by nnz across ranks (one process per GPU), B replicated, every rank
          computes its C row block and stores every finished row into every GPU's copy of C from the kernel's
          epilogue (NVLink multimem / peer stores, in-kernel arrival flags; `--gather nccl` = one NCCL all-gather
          instead); the exchange is inside the timed region ("scaling": "strong"), `no_exchange` reports the
          step without it.
sub-records of the same line: `clustered` (block-model graph: natural / reordered / hot-cold tiles), `products`
          (configs[4], with the speed-up against one GPU of the same box at N > 1), `arxiv` (configs[0], CUDA-graph
          replay, whole-matrix parity; N = 1), `column_sharded` (N > 1), `selftest_multi` (N > 1), and
          `parity_all_ranks`: every rank checks sampled rows of EVERY rank's block against the oracle.
--impl reference : the reference's CPU path (`--version=cpu`: row-parallel CSR SpMM on the host cores,
          restated in oracle/spmm_oracle.c because torch_sparse is not installable) on a bounded sample.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL writes its banner ("NCCL version ...") to STDOUT at the VERSION/INFO levels; rank 0 must print one JSON line
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

HIDDEN_SWEEP = [16, 32, 64, 128]
SHAPE = "reddit"
METRIC = "spmm_gflops"
UNIT = "GFLOP/s"


def alg_bytes_csr(n_rows, n_cols, nnz, hidden, s=4, fmt="CSR"):
    """ALGORITHMIC bytes of one SpMM: row index (rowptr | rowind) + colind + values + B once +
    C once."""
    row_index = 4 * (n_rows + 1) if fmt == "CSR" else 4 * nnz
    return row_index + 4 * nnz + s * nnz + s * n_cols * hidden + s * n_rows * hidden


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        """Samples before this point (start-up, warm-up) are not reported."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t0 = getattr(self, "t0", 0.0)
        for ts, r in self.rows:
            if ts < t0:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def auto_ds_parts(n_cols, hidden, info, s=4, nnz=None):
    """Column tiling so that one B tile stays L2-resident next to the streaming A: see pygim_b200.utils.autotuner."""
    from pygim_b200.utils import autotuner
    return autotuner.choose_ds_parts(n_cols, hidden, s, info["l2_bytes"], nnz=nnz)


def make_args(hidden, dtype, fmt="CSR"):
    return types.SimpleNamespace(data_type=dtype, sp_format=fmt, hidden_size=hidden, sp_parts=1, ds_parts=1)


# ====================================================================================== CPU arms
def cpu_spmm_sample(O, clib, rowptr, col, x_by_h, threads, repeats=1):
    """Times the row-parallel CSR SpMM (the `--version=cpu` algorithm class) over the sweep on the given
    row block.  Returns (seconds for one sweep pass [best of repeats], flops of one pass)."""
    import numpy as np
    nnz = int(rowptr[-1])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for h, x in x_by_h.items():
            out = np.empty((rowptr.shape[0] - 1, h), dtype=x.dtype)
            O.spmm_csr_rowpar(rowptr, col, None, x, nthreads=threads, out=out, clib=clib)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, sum(2.0 * nnz * h for h in x_by_h)


def cpu_sample_graph(n, nnz, max_deg, sample_rows, seed=0):
    """The first `sample_rows` rows of the Reddit-shaped graph, generated on the host."""
    from pygim_b200 import graphgen
    rowptr, col = graphgen.synthetic_csr(n, nnz, max_deg, seed=seed, rows=(0, sample_rows))
    return rowptr.numpy().astype("int32"), col.numpy().astype("int32")


def workload_config(shape, clustered, dtype, fmt, sweep, n, nnz):
    """The `config` block: what is computed, identically worded in both arms (`--impl ours` / `--impl reference`).
    How OUR arm executes it (column tiling, sharding, fast paths) is the separate `plan` block."""
    return {"workload": "%s-shape%s %s %s SpMM, hidden sweep %s" % (shape, " (block-model communities)" if clustered else "",
                                                                    dtype, fmt, "/".join(map(str, sweep))),
            "nodes": n, "edges": nnz, "hidden_sweep": list(sweep), "format": fmt, "sp_parts": 1,
            "values": "all ones (value-less adjacency)", "features": "integers in [-8, 3] (spmm_test.py:70)",
            "l2": "inputs larger than L2 (A = %.0f MB is streamed every launch; nothing is cached between steps)"
                  % ((8.0 * nnz) / 1e6)}


def run_reference(a):
    """--impl reference: PyGim's CPU path timed on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import oracle as O
    from pygim_b200 import graphgen
    O.build()
    native = O.build_native()
    clib = O.lib(native) if native else O.lib()
    threads = O.max_threads()
    n, nnz, max_deg = graphgen.SHAPES[SHAPE]
    sample_rows = a.cpu_sample_rows or max(1024, n // 4)       # ~28 M nonzeros x the sweep per step
    rowptr, col = cpu_sample_graph(n, nnz, max_deg, sample_rows)
    x_by_h = {h: graphgen.reference_features(n, h, torch.float32, seed=h).numpy() for h in HIDDEN_SWEEP}
    for _ in range(a.warmup):
        cpu_spmm_sample(O, clib, rowptr, col, x_by_h, threads)
    t0 = time.perf_counter()
    flops = 0.0
    for _ in range(a.steps):
        _, f = cpu_spmm_sample(O, clib, rowptr, col, x_by_h, threads)
        flops += f
    dt = time.perf_counter() - t0
    value = flops / dt / 1e9
    sample = "rows [0,%d) of the Reddit-shaped graph (%d nnz), hidden sweep %s, per step" % (
        sample_rows, int(rowptr[-1]), HIDDEN_SWEEP)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / max(a.steps, 1) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(SHAPE, False, "FLT32", "CSR", HIDDEN_SWEEP, n, nnz),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "native_build": bool(native)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ====================================================================================== GPU arm
class SweepWorkload:
    """One graph shape, row-sharded over the ranks, with one plan per hidden size of the sweep."""

    def __init__(self, a, shape, dev, rank, world, clustered=False, reorder=None, sweep=None, tag=""):
        import torch
        from pygim_b200 import graphgen
        from pygim_b200.backend_pim import pim_ops
        from pygim_b200.backend_pim.spmm import TORCH_TYPES, SparseTensorCOO
        from pygim_b200.sharded import ShardedSpMM, shard_rows
        from pygim_b200.sparse_tensor import SparseTensor
        self.a, self.shape, self.dev, self.rank, self.world, self.tag = a, shape, dev, rank, world, tag
        self.clustered, self.reorder = clustered, reorder
        self.dtype = TORCH_TYPES[a.dtype]
        self.esize = torch.empty((), dtype=self.dtype).element_size()
        self.n, self.nnz, self.max_deg = graphgen.SHAPES[shape]
        self.sweep = list(sweep or a.hidden or HIDDEN_SWEEP)
        self.info = pim_ops.device_info()
        n, nnz, max_deg = self.n, self.nnz, self.max_deg
        # ---- this rank's row shard (nnz-balanced, the GPU-level partition_by_nnz_csr)
        self.deg = graphgen.degree_sequence(n, nnz, max_deg, n, seed=0)
        full_rowptr = torch.zeros(n + 1, dtype=torch.int64)
        torch.cumsum(self.deg, 0, out=full_rowptr[1:])
        self.splits = pim_ops.partition_rows_by_nnz(full_rowptr, world) if world > 1 else [0, n]
        self.r0, self.r1 = self.splits[rank], self.splits[rank + 1]
        self._full = None
        if clustered:
            # the block-model generator draws rows in one sequence: every rank generates the graph and keeps its block
            rowptr, col = graphgen.clustered_csr(n, nnz, max_deg, seed=0, device=str(dev), deg=self.deg)
            full = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(n, n), is_sorted=True)
            self._full = full
            adj = full if world == 1 else shard_rows(full, self.r0, self.r1)
        else:
            rowptr, col = graphgen.synthetic_csr(n, nnz, max_deg, seed=0, device=str(dev), rows=(self.r0, self.r1),
                                                 deg=self.deg)
            adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(self.r1 - self.r0, n), is_sorted=True)
        del rowptr, col
        self.adj_plain = adj                       # rows in natural order (parity checks read it)
        self.shard_nnz = int(adj.nnz())
        self.reorder_stats = None
        perm, hot = None, None
        if reorder:
            from pygim_b200 import reorder as R
            t0 = time.perf_counter()
            adj, perm, self.reorder_stats = R.reorder_rows(adj, reorder)
            groups = self.reorder_stats.pop("group_of_row", None)
            if reorder == "tiles":
                adj, hot = R.hot_cold_plan(adj, groups, hot_k=a.hot_k, super_nnz=a.tile_super_nnz)
                self.reorder_stats.update(hot_coverage=hot["coverage"], tile_supertickets=hot["supertickets"],
                                          hot_k=hot["hot_k"], seg_len=hot["seg_len"])
            del groups
            torch.cuda.synchronize()
            self.reorder_stats["seconds"] = time.perf_counter() - t0
        self.ds_parts = {h: (a.ds_parts if a.ds_parts > 0 else auto_ds_parts(n, h, self.info, self.esize, nnz=nnz))
                         for h in self.sweep}
        base = SparseTensorCOO(adj, dtype=self.dtype, format=a.format)
        base.row_perm = perm
        base.hot_plan = hot
        base.build_csr() if a.format == "CSR" else base.build_coo()
        self.plans = {}
        for h in self.sweep:     # one set of int32 CSR/COO arrays shared by the plans (one plan per hidden size)
            A = copy.copy(base)
            A.sp_info_ptr = None
            A.to_pim_group(h, self.ds_parts[h])
            self.plans[h] = A

        def args_for(h):
            ns = make_args(h, self.dtype, a.format)
            ns.ds_parts = self.ds_parts[h]
            return ns

        self.gather = a.gather
        self.ops = {h: ShardedSpMM(None, args_for(h), splits=self.splits, local_adj=adj,
                                   make_local=lambda _adj, _args, _h=h: self.plans[_h], chunks=1,
                                   fused=(world > 1 and a.gather == "fused"), use_multicast=not a.no_multicast,
                                   sync=a.sync, world=world, rank=rank) for h in self.sweep}
        for h in self.sweep:
            hdl = self.plans[h].sp_info_ptr
            if a.general_kernel:     # force the weighted kernels although the adjacency is value-less (all ones)
                pim_ops.plan_set_option(hdl, "unit_values", 0)
            if a.short_rows is not None:
                pim_ops.plan_set_option(hdl, "short_rows", a.short_rows)
            if a.no_l2_persist:
                pim_ops.plan_set_option(hdl, "l2_persist", 0)
            for kv in a.opt or []:
                k, v = kv.split("=")
                pim_ops.plan_set_option(hdl, k, int(v))
        self.x_dev = {h: graphgen.reference_features(n, h, self.dtype, seed=h, device=str(dev)) for h in self.sweep}
        self.c_full = {h: torch.empty((n, h), dtype=self.dtype, device=dev) for h in self.sweep}
        self.c_last = dict(self.c_full)

    def probe_fused(self):
        """The fused path needs NVLink peer mappings (symmetric memory); if this box cannot provide them every rank
        falls back to the NCCL all-gather together and the JSON line says so."""
        import torch
        import torch.distributed as dist
        if self.world == 1 or self.gather != "fused":
            return
        ok = torch.ones(1, device=self.dev)
        try:
            self.ops[self.sweep[0]].mul(self.x_dev[self.sweep[0]])
            torch.cuda.synchronize()
        except Exception as exc:      # noqa: BLE001
            ok.zero_()
            if self.rank == 0:
                print("bench.py: fused all-gather unavailable (%s); using NCCL" % str(exc)[:200], file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            self.gather = "nccl"
            for h in self.sweep:
                self.ops[h].fused = False

    def step(self, record=None, exchange=True):
        for i, h in enumerate(self.sweep):
            if record is not None:
                record[i][0].record()
            if not exchange:
                self.ops[h].mul(self.x_dev[h], out=self.c_full[h], gather=False)     # this rank's row block only
            elif self.world > 1 and self.gather == "fused":
                self.c_last[h] = self.ops[h].mul(self.x_dev[h])   # rows land in every peer's symmetric buffer
            else:
                self.ops[h].mul(self.x_dev[h], out=self.c_full[h])   # N > 1: SpMM of the row block + NCCL all-gather
                self.c_last[h] = self.c_full[h]
            if record is not None:
                record[i][1].record()

    def sync(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(self, steps, warmup, exchange=True):
        """(ms per step [max over ranks], per-hidden ms [max over ranks]) with CUDA events on the launching stream."""
        import torch
        import torch.distributed as dist
        for _ in range(warmup):
            self.step(exchange=exchange)
        self.sync()
        ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in self.sweep]
              for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync()
        t0.record()
        for k in range(steps):
            self.step(ev[k], exchange=exchange)
        t1.record()
        self.sync()
        ms = t0.elapsed_time(t1) / steps
        per_h = [sum(ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(steps)) / steps for i in range(len(self.sweep))]
        if self.world > 1:
            t = torch.tensor([ms] + per_h, dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, per_h = float(t[0]), [float(v) for v in t[1:]]
        return ms, per_h

    def graph_us(self, reps=10, replays=5):
        """Per-hidden device time (us) of one SpMM as a CUDA-graph replay of `reps` back-to-back calls: what a launch
        costs without the host's enqueue path - the figure that matters for graphs whose SpMM takes tens of us."""
        import torch
        assert self.world == 1
        out = []
        cur = torch.cuda.current_stream()
        for h in self.sweep:
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.ops[h].mul(self.x_dev[h], out=self.c_full[h])
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    self.ops[h].mul(self.x_dev[h], out=self.c_full[h])
            g.replay()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(replays):
                g.replay()
            t1.record()
            torch.cuda.synchronize()
            out.append(t0.elapsed_time(t1) / (reps * replays) * 1e3)
            del g
        return out

    def flops_step(self):
        return sum(2.0 * self.nnz * h for h in self.sweep)

    def block_rows(self, j):
        """(rowptr, col) of rank j's row block on this device - regenerated from the seed for foreign blocks."""
        from pygim_b200 import graphgen
        from pygim_b200.sharded import shard_rows
        if j == self.rank:
            rp, cl, _ = self.adj_plain.csr()
            return rp, cl
        if self._full is not None:
            sh = shard_rows(self._full, self.splits[j], self.splits[j + 1])
            rp, cl, _ = sh.csr()
            return rp, cl
        return graphgen.synthetic_csr(self.n, self.nnz, self.max_deg, seed=0, device=str(self.dev),
                                      rows=(self.splits[j], self.splits[j + 1]), deg=self.deg)

    def parity_all_ranks(self, O, results, x_host, rows_per_block=256):
        """EVERY rank checks sampled rows of EVERY rank's block of the gathered result - including each block's
        longest (segmented) row - against the oracle; the verdicts are AND-reduced.  This is the reference's
        exact-equality self-check (spmm_multigroup/mul_csr_multigroup.c:550-620) applied to what was just timed."""
        import numpy as np
        import torch
        import torch.distributed as dist
        ok, checked = True, 0
        for j in range(self.world):
            rp, cl = self.block_rows(j)
            nb = rp.numel() - 1
            if nb == 0:
                continue
            deg = rp[1:] - rp[:-1]
            g = torch.Generator().manual_seed(1000 + j)
            pick = torch.randperm(nb, generator=g)[: max(0, min(rows_per_block, nb) - 1)]
            pick = torch.unique(torch.cat([pick, torch.argmax(deg).cpu().view(1)])).to(rp.device)
            cnt = deg[pick]
            src = torch.repeat_interleave(rp[:-1][pick], cnt) + \
                (torch.arange(int(cnt.sum()), device=rp.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
            sub_col = cl[src].cpu().numpy().astype("int32")
            sub_rp = np.zeros(pick.numel() + 1, dtype="int32")
            np.cumsum(cnt.cpu().numpy(), out=sub_rp[1:])
            rows_glob = (pick + self.splits[j]).to(self.dev)
            for h in self.sweep:
                want = O.spmm_csr_rowpar(sub_rp, sub_col, None, x_host[h].numpy(), nthreads=O.max_threads())
                got = results[h][rows_glob].cpu().numpy()
                ok = ok and bool(np.array_equal(want, got))
            checked += int(pick.numel())
            del rp, cl
        if self.world > 1:
            t = torch.tensor([1.0 if ok else 0.0], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(float(t) == 1.0)
        return ok, checked

    def free(self):
        import torch
        for h in self.sweep:
            self.ops[h]._symm.clear()
            self.plans[h].free()
        self.x_dev.clear(); self.c_full.clear(); self.c_last.clear()
        self._full = None
        self.adj_plain = None
        torch.cuda.empty_cache()


def per_hidden_table(w, per_h_ms, peak):
    out = []
    for i, h in enumerate(w.sweep):
        # per-GPU algorithmic bytes of this rank's SpMM (shard of A and C, all of B), A counted once
        b = alg_bytes_csr(w.r1 - w.r0, w.n, w.shard_nnz, h, w.esize, w.a.format)
        out.append({"hidden": h, "kernel_ms": per_h_ms[i], "gflops": 2.0 * w.shard_nnz * h / per_h_ms[i] / 1e6,
                    "alg_gbs": b / per_h_ms[i] / 1e6, "frac_hbm": b / per_h_ms[i] / 1e6 / peak,
                    "gather_gbs": float(w.esize) * w.shard_nnz * h / per_h_ms[i] / 1e6,
                    "launches": w.ds_parts[h], "tile_cols": h // w.ds_parts[h]})
    return out


def selftest_multi(a, dev, rank, world):
    """Whole-matrix parity of every multi-GPU mode on a small Reddit-like graph (what tests/test_gpu_multi.py checks,
    run here because the driver's GPU-test box has one GPU): every rank compares the FULL gathered result."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from pygim_b200 import graphgen
    from pygim_b200.sharded import ColumnShardedSpMM, ShardedSpMM
    O.build()
    adj = graphgen.synthetic_adj("reddit", scale=0.01, seed=5)
    n = adj.size(0)
    rowptr, col, _ = adj.csr()
    ok, modes = True, []
    for dtype, hidden in ((torch.float32, 64), (torch.int32, 32)):
        x = graphgen.reference_features(n, hidden, dtype, seed=1)
        args = make_args(hidden, dtype, "CSR")
        want = O.spmm_csr_rowpar(rowptr.numpy(), col.numpy(), None, x.numpy())
        for kw in (dict(chunks=1), dict(chunks=3), dict(fused=True, sync="flags"),
                   dict(fused=True, sync="flags", use_multicast=False), dict(fused=True, sync="barrier")):
            try:
                op = ShardedSpMM(adj.to(str(dev)), args, **kw)
                for _ in range(5):
                    out = op.mul(x.to(dev))
                torch.cuda.synchronize()
                good = bool(np.array_equal(out.cpu().numpy(), want))
                op.free()
            except Exception as exc:       # noqa: BLE001
                good = False
                print("selftest-multi rank %d: %s failed: %s" % (rank, kw, str(exc)[:300]), file=sys.stderr)
            ok = ok and good
            if dtype == torch.float32:
                modes.append("+".join("%s=%s" % kv for kv in kw.items()))
    x = graphgen.reference_features(n, 48, torch.float32, seed=2)
    cop = ColumnShardedSpMM(adj.to(str(dev)), make_args(48, torch.float32, "CSR"))
    out = cop.mul(x.to(dev))
    torch.cuda.synchronize()
    ok = ok and bool(np.array_equal(out.cpu().numpy(), O.spmm_csr_rowpar(rowptr.numpy(), col.numpy(), None, x.numpy())))
    cop.free()
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(float(t) == 1.0), modes + ["column-sharded"]


def run_ours(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the aggregation path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from pygim_b200.backend_pim import pim_ops
    pim_ops.dpu_init_ranks(1)
    steps, warmup = a.steps, max(a.warmup, 3)

    selftest = None
    if world > 1 and not a.no_selftest:
        selftest = selftest_multi(a, dev, rank, world)

    w = SweepWorkload(a, a.shape, dev, rank, world, clustered=a.clustered, reorder=a.reorder)
    w.probe_fused()
    sweep, n, nnz, esize = w.sweep, w.n, w.nnz, w.esize
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.1 s to deliver its first sample: start before the warm-up
    for _ in range(warmup):
        w.step()
    w.sync()
    sampler.mark()               # report only samples taken from here (timed region + e2e region) on
    ms_step, per_h_ms = w.timed(steps, 0)
    launches_per_step = sum(pim_ops.last_launches(w.plans[h].sp_info_ptr) for h in sweep)
    if world > 1:
        launches_per_step += len(sweep) if (w.gather == "fused" and a.sync == "flags") else 0     # the one-warp waits
    value = w.flops_step() / (ms_step * 1e-3) / 1e9
    # the same step WITHOUT the exchange (report both)
    no_exchange = None
    if world > 1:
        ms_nx, per_h_nx = w.timed(steps, 2, exchange=False)
        no_exchange = {"value": w.flops_step() / (ms_nx * 1e-3) / 1e9, "ms_per_step": ms_nx, "per_hidden_ms": per_h_nx}
        w.step()                 # c_last holds gathered results again (parity below)
        w.sync()

    # ---- e2e: the same step through the public API with pinned HOST operands (H2D + D2H inside the timing)
    x_host = {h: w.x_dev[h].cpu().pin_memory() for h in sweep}
    e2e = None
    if not a.no_e2e:
        if world == 1:
            c_host = {h: torch.empty((n, h), dtype=w.dtype).pin_memory() for h in sweep}
            handles = [w.plans[h].sp_info_ptr for h in sweep]

            def step_host():
                if a.e2e_mode == "pipelined":     # one software pipeline over the sweep's four calls
                    pim_ops.spmm_run_dense_many(handles, [x_host[h] for h in sweep], [c_host[h] for h in sweep])
                else:                             # four independent synchronous calls (host entry point)
                    for h in sweep:
                        w.plans[h].mul(x_host[h], out=c_host[h])
            h2d = sum(n * h * esize for h in sweep)
            d2h = h2d
        else:
            # every rank uploads only ITS 1/N row block of B over PCIe; the blocks are all-gathered over NVLink into
            # the full operand, the SpMM runs from device memory and only this rank's rows of C go back to the host
            # equal blocks of `pad` rows (the last one shorter): the all-gather lands every block in place in a padded
            # [world * pad, H] operand whose first n rows ARE the full B - no un-padding copies
            pad = -(-n // world)
            blk = [min(n, r * pad) for r in range(world + 1)]
            b0, b1 = blk[rank], blk[rank + 1]
            xb_host = {h: x_host[h][b0:b1].clone().pin_memory() for h in sweep}
            xg = {h: torch.empty((world * pad, h), dtype=w.dtype, device=dev) for h in sweep}
            x_full = {h: xg[h][:n] for h in sweep}
            c_host = {h: torch.empty((w.r1 - w.r0, h), dtype=w.dtype).pin_memory() for h in sweep}
            c_loc = {h: torch.empty((n, h), dtype=w.dtype, device=dev) for h in sweep}

            s_in, s_ag, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            order = sorted(sweep, reverse=True)          # widest operand first: its chain is the longest

            def step_host():
                # four streams: the uploads, the NVLink all-gathers of the NEXT operands and the downloads of finished
                # results all overlap the SpMM of the current operand
                cur = torch.cuda.current_stream(dev)
                for s_ in (s_in, s_ag, s_out):
                    s_.wait_stream(cur)
                up, ag = {}, {}
                with torch.cuda.stream(s_in):
                    for h in order:
                        xg[h][b0:b1].copy_(xb_host[h], non_blocking=True)
                        up[h] = torch.cuda.Event()
                        up[h].record(s_in)
                with torch.cuda.stream(s_ag):
                    for h in order:             # same order on every rank
                        s_ag.wait_event(up[h])
                        dist.all_gather_into_tensor(xg[h], xg[h][rank * pad:(rank + 1) * pad])
                        ag[h] = torch.cuda.Event()
                        ag[h].record(s_ag)
                for h in order:
                    cur.wait_event(ag[h])
                    w.ops[h].mul(x_full[h], out=c_loc[h], gather=False)
                    done = torch.cuda.Event()
                    done.record(cur)
                    s_out.wait_event(done)
                    with torch.cuda.stream(s_out):
                        c_host[h].copy_(c_loc[h][w.r0:w.r1], non_blocking=True)
                cur.wait_stream(s_out)
                torch.cuda.synchronize()
            h2d = sum((b1 - b0) * h * esize for h in sweep)
            d2h = sum((w.r1 - w.r0) * h * esize for h in sweep)
        e2e_steps = max(3, min(steps, 10))
        for _ in range(2):
            step_host()
        w.sync()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        w.sync()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            e2e_s, h2d, d2h = float(tm[0]), int(t[1]), int(t[2])
        e2e = {"value": w.flops_step() * e2e_steps / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
               "mode": (a.e2e_mode if world == 1 else "row block of B per rank over PCIe + NVLink all-gather (in place, "
                        "overlapping the previous operand's SpMM), local rows of C back (bytes are totals over ranks)")}
        if world == 1:   # per call: exposed upload, kernel window, exposed download tail (pygim_last_timers)
            e2e["phases_ms"] = {str(h): pim_ops.last_timers(w.plans[h].sp_info_ptr) for h in sweep}
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity of what was just timed: every rank, sampled rows of every rank's block, against the oracle
    parity, parity_rows, parity_e2e = None, 0, None
    if not a.no_check:
        import numpy as np
        from oracle import oracle as O
        O.build()
        parity, parity_rows = w.parity_all_ranks(O, w.c_last, x_host)
        if e2e is not None:      # the host-path results: this rank's first rows
            rows_chk = min(256, w.r1 - w.r0)
            rp, cl, _ = w.adj_plain.csr()
            rp_h = rp[: rows_chk + 1].cpu().numpy().astype("int32")
            cl_h = cl[: int(rp_h[-1])].cpu().numpy().astype("int32")
            parity_e2e = True
            for h in sweep:
                want = O.spmm_csr_rowpar(rp_h, cl_h, None, x_host[h].numpy(), nthreads=O.max_threads())
                off = w.r0 if world == 1 else 0
                parity_e2e = parity_e2e and bool(np.array_equal(want, c_host[h][off:off + rows_chk].numpy()))
            if world > 1:
                t = torch.tensor([1.0 if parity_e2e else 0.0], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                parity_e2e = bool(float(t) == 1.0)

    peak, peak_kind = measured_peaks()
    per_hidden = per_hidden_table(w, per_h_ms, peak)
    shard_nnz, r0, r1, ds_parts = w.shard_nnz, w.r0, w.r1, w.ds_parts
    adj_for_cpu = w.adj_plain if (world == 1 and not a.no_cpu) else None
    if world > 1:
        x_full = xg = c_loc = None
    if adj_for_cpu is None:
        w.free()

    # ---- sub-records (same process, same box): the clustered graph with prepare-time reordering, products-shape
    clustered_rec = products_rec = arxiv_rec = column_rec = None
    if a.shape == "reddit" and not a.clustered and a.dtype == "FLT32" and a.format == "CSR" and not a.hidden:
        if world == 1 and not a.no_clustered:
            clustered_rec = run_sub_workload(a, "reddit", dev, rank, world, peak, clustered=True)
        if not a.no_products:
            products_rec = run_sub_workload(a, "products", dev, rank, world, peak, clustered=False)
        if world == 1 and not a.no_arxiv:
            arxiv_rec = run_small_graph(a, "arxiv", dev, peak)
        if world > 1 and not a.no_column_sharded:
            column_rec = run_column_sharded(a, dev, rank, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of the step)
    # The dominant KERNEL is the template instantiation with the largest share of the step.  An instantiation
    # is fixed by the dense tile a launch handles (G = lanes per tile row), so hidden sizes that run as
    # several column tiles (H = 128 -> two 64-column launches) are launches of the SAME kernel as H = 64.
    # Per launch the algorithmic bytes are A + that tile of B + that tile of C (SURVEY.md 8d with H = tile).
    by_kernel = {}
    for i, h in enumerate(sweep):
        wd = h // ds_parts[h]
        k = by_kernel.setdefault(wd, {"ms": 0.0, "launches": 0, "bytes": 0.0, "hidden": []})
        k["ms"] += per_h_ms[i]
        k["launches"] += ds_parts[h]
        k["bytes"] += ds_parts[h] * alg_bytes_csr(r1 - r0, n, shard_nnz, wd, esize, a.format)
        k["hidden"].append(h)
    dom_w = max(by_kernel, key=lambda wd: by_kernel[wd]["ms"])
    dk = by_kernel[dom_w]
    lanes = max(1, min(32, dom_w * esize // 16))
    roof = {"bound": "hbm",
            "kernel": "csr_spmm_kernel<%s, G=%d> (%d-column tiles; hidden %s)%s"
                      % (a.dtype, lanes, dom_w, "/".join(map(str, dk["hidden"])),
                         " over the row pointer derived from the sorted COO stream" if a.format == "COO" else ""),
            "achieved": dk["bytes"] / dk["ms"] / 1e6, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
            "frac": dk["bytes"] / dk["ms"] / 1e6 / peak, "traffic": None,
            "alg_bytes_per_launch": dk["bytes"] / dk["launches"], "launches_per_step": dk["launches"],
            "avg_launch_ms": dk["ms"] / dk["launches"], "share_of_step": dk["ms"] / sum(per_h_ms),
            "sweep_achieved": sum(alg_bytes_csr(r1 - r0, n, shard_nnz, h, esize, a.format) for h in sweep)
            / sum(per_h_ms) / 1e6,
            "note": "uniformly random Reddit-shape is bound by L2 -> SM gather traffic (s*nnz*H bytes leave L2 per "
                    "launch), see DESIGN.md 4.3; the `clustered` sub-record is the same shape with community structure"}
    # secondary bound: the measured ceiling of random row gathers (tools/l2_gather_probe)
    try:
        with open(os.path.join(ROOT, "profiles", "gather_ceiling.json")) as f:
            ceil = json.load(f)
        row_bytes = dom_w * esize
        key = str(min((64, 128, 256, 512), key=lambda b: abs(b - row_bytes)))
        resident = n * row_bytes <= 0.46 * w.info["l2_bytes"]
        peak_g = ceil["l2_resident_tbs" if resident else "hbm_served_tbs"][key]
        ach_g = float(esize) * shard_nnz * sum(h for h in dk["hidden"]) / dk["ms"] / 1e9
        roof["gather"] = {"bound": "l2-gather" if resident else "hbm-gather", "achieved_tbs": ach_g,
                          "ceiling_tbs": peak_g, "frac": ach_g / peak_g, "row_bytes": row_bytes,
                          "what": "s*nnz*H bytes gathered per launch / time, against the measured ceiling of random "
                                  "row gathers (profiles/r01_l2_gather_probe.txt)"}
    except Exception:
        pass
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path) and (a.shape, a.dtype, a.format, world, a.clustered) == ("reddit", "FLT32", "CSR", 1, False):
        try:
            with open(traffic_path) as f:
                tj = json.load(f)
            roof["traffic"] = tj.get("tile_%d" % dom_w)
            roof["traffic_source"] = tj.get("source", "profiles/traffic.json (one ncu --set full capture of this kernel; "
                                                      "not measured in this run)")
        except Exception:
            pass

    # ---- CPU baseline beside it (rank 0, N == 1 only): bounded sample of the same workload
    cpu = None
    if adj_for_cpu is not None:
        from oracle import oracle as O
        O.build()
        native = O.build_native()
        clib = O.lib(native) if native else O.lib()
        threads = O.max_threads()
        sample_rows = min(a.cpu_sample_rows or n, r1 - r0)         # default: the whole workload
        rp, cl, _ = adj_for_cpu.csr()
        rp_h = rp[: sample_rows + 1].cpu().numpy().astype("int32")
        cl_h = cl[: int(rp_h[-1])].cpu().numpy().astype("int32")
        xs = {h: x_host[h].numpy() for h in sweep}
        secs, fl = cpu_spmm_sample(O, clib, rp_h, cl_h, xs, threads, repeats=3)
        cpu = {"value": fl / secs / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "rows [0,%d) of the same graph (%d nnz), hidden sweep %s, best of 3 passes (%.2f s each)"
                         % (sample_rows, int(rp_h[-1]), sweep, secs), "native_build": bool(native)}
        w.free()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {"FLT32": "f32", "DBL64": "f64", "INT8": "i8", "INT16": "i16", "INT32": "i32", "INT64": "i64"}[a.dtype],
        "data": "synthetic",
        "config": workload_config(a.shape, a.clustered, a.dtype, a.format, sweep, n, nnz),
        "plan": {"values": ("general weighted kernel forced" if a.general_kernel
                            else "unit-value fast path: value stream not read, results bit-identical"),
                 "ds_parts": {str(h): ds_parts[h] for h in sweep},
                 "reorder": a.reorder or "none",
                 "sharding": ("rows by nnz over %d GPUs, B replicated, all-gather of C %s, inside the timing"
                              % (world, ("fused into the kernel epilogue (NVLink %s stores, %s)"
                                         % ("per-peer" if a.no_multicast else "multimem",
                                            "in-kernel arrival flags, no barrier" if a.sync == "flags" else "one barrier per call"))
                                 if w.gather == "fused" else "by NCCL"))
                 if world > 1 else "single GPU",
                 "l2_bytes": w.info["l2_bytes"]},
        "e2e": e2e, "gpu_launches": launches_per_step * steps,
        "roofline": roof, "per_hidden": per_hidden, "cpu_baseline": cpu, "clocks": clocks,
        "parity_all_ranks": parity, "parity_rows_checked_per_rank": parity_rows, "parity_e2e": parity_e2e,
        "no_exchange": no_exchange, "selftest_multi": None if selftest is None else {"ok": selftest[0], "modes": selftest[1]},
        "clustered": clustered_rec, "products": products_rec, "arxiv": arxiv_rec, "column_sharded": column_rec,
        "reorder_stats": w.reorder_stats,
        "lib": os.path.relpath(__import__("pygim_b200._lib", fromlist=["x"]).loaded_path() or "", ROOT),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_column_sharded(a, dev, rank, world):
    """SURVEY.md 8(e) / north_star "optionally column-sharded": the FEATURE COLUMNS dealt over the GPUs (every rank
    streams all of A and computes all rows of its H / N columns), with and without the all-gather of the column
    blocks, for the wide end of the sweep - beside the row-sharded default of the main line."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from pygim_b200 import graphgen
    from pygim_b200.backend_pim.spmm import TORCH_TYPES
    from pygim_b200.sharded import ColumnShardedSpMM
    from pygim_b200.sparse_tensor import SparseTensor
    n, nnz, max_deg = graphgen.SHAPES["reddit"]
    rowptr, col = graphgen.synthetic_csr(n, nnz, max_deg, seed=0, device=str(dev))
    adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(n, n), is_sorted=True)
    dtype = TORCH_TYPES[a.dtype]
    rec = {"per_hidden": []}
    ok = True
    for h in (64, 128):
        op = ColumnShardedSpMM(adj, make_args(h, dtype, a.format))
        x = graphgen.reference_features(n, h, dtype, seed=h, device=str(dev))
        times = {}
        for gather in (True, False):
            for _ in range(3):
                out = op.mul(x, gather=gather)
            torch.cuda.synchronize()
            dist.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10):
                out = op.mul(x, gather=gather)
            t1.record()
            torch.cuda.synchronize()
            t = torch.tensor([t0.elapsed_time(t1) / 10], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times[gather] = float(t)
        if not a.no_check:      # this rank's column block, first 512 rows, against the oracle
            O.build()
            rp = rowptr[:513].cpu().numpy().astype("int32")
            cl = col[: int(rp[-1])].cpu().numpy().astype("int32")
            want = O.spmm_csr_rowpar(rp, cl, None, x.cpu().numpy())
            ok = ok and bool(np.array_equal(out[:512].cpu().numpy(), want[:, op.c0:op.c1]))
        rec["per_hidden"].append({"hidden": h, "columns_per_gpu": op.c1 - op.c0, "ms_with_all_gather": times[True],
                                  "ms_without": times[False], "gflops_with_all_gather": 2.0 * nnz * h / times[True] / 1e6,
                                  "gflops_without": 2.0 * nnz * h / times[False] / 1e6})
        op.free()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    rec["parity_all_ranks"] = bool(float(flag) == 1.0) if not a.no_check else None
    del adj, rowptr, col
    torch.cuda.empty_cache()
    return rec


def run_small_graph(a, shape, dev, peak):
    """arxiv-shape (BASELINE.json configs[0], the reference's own CPU-runnable case): an SpMM takes tens of
    microseconds, so it is timed both call by call (CUDA events, host enqueue included) and as a CUDA-graph replay,
    and EVERY element is compared with the oracle."""
    import torch
    from oracle import oracle as O
    w = SweepWorkload(a, shape, dev, 0, 1)
    ms, per_h = w.timed(max(5, min(a.steps, 10)), 3)
    us = w.graph_us()
    rec = {"config": {"workload": "%s-shape FLT32 CSR SpMM, hidden sweep %s" % (shape, "/".join(map(str, w.sweep))),
                      "nodes": w.n, "edges": w.nnz},
           "value": w.flops_step() / (sum(us) * 1e-6) / 1e9, "unit": UNIT, "timing": "CUDA-graph replay",
           "per_hidden": [{"hidden": h, "graph_us": u, "per_call_events_us": p * 1e3, "gflops": 2.0 * w.nnz * h / (u * 1e-6) / 1e9,
                           "frac_hbm": alg_bytes_csr(w.n, w.n, w.nnz, h, w.esize, a.format) / (u * 1e-6) / 1e9 / peak}
                          for h, u, p in zip(w.sweep, us, per_h)]}
    if not a.no_check:
        O.build()
        rp, cl, _ = w.adj_plain.csr()
        rp, cl = rp.cpu().numpy().astype("int32"), cl.cpu().numpy().astype("int32")
        bad = 0
        for h in w.sweep:
            got = w.ops[h].mul(w.x_dev[h], out=w.c_full[h])
            torch.cuda.synchronize()
            want = O.spmm_csr_rowpar(rp, cl, None, w.x_dev[h].cpu().numpy(), nthreads=O.max_threads())
            bad += int((got.cpu().numpy() != want).sum())
        rec["parity_whole_matrix"] = bad == 0
    w.free()
    return rec


def run_sub_workload(a, shape, dev, rank, world, peak, clustered):
    """A second workload in the same process (same box, same clocks): the block-model Reddit-shape graph with
    prepare-time reordering (N = 1), or products-shape (BASELINE.json configs[4]) with and without the exchange and -
    at N > 1 - rank 0 alone on the whole graph as the single-GPU reference of the speed-up."""
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    steps, warmup = max(5, min(a.steps, 10)), 3
    rec = {}
    variants = [("natural_order", None), ("reordered_cluster", "cluster"), ("reordered_tiles", "tiles")] if clustered \
        else [("sharded", None)]
    for name, reorder in variants:
        w = SweepWorkload(a, shape, dev, rank, world, clustered=clustered, reorder=reorder)
        w.probe_fused()
        ms, per_h = w.timed(steps, warmup)
        r = {"value": w.flops_step() / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms,
             "per_hidden": per_hidden_table(w, per_h, peak), "reorder_stats": w.reorder_stats}
        if world > 1:
            ms_nx, per_nx = w.timed(steps, 2, exchange=False)
            r["no_exchange"] = {"value": w.flops_step() / (ms_nx * 1e-3) / 1e9, "ms_per_step": ms_nx, "per_hidden_ms": per_nx}
            # NVLink bytes every GPU must RECEIVE per step with a replicated result, against the measured 770 GB/s
            ingress = sum((w.n - (w.r1 - w.r0)) * h * w.esize for h in w.sweep)
            r["exchange_floor_ms"] = ingress / 770e9 * 1e3
            w.step()
            w.sync()
        if not a.no_check:
            O.build()
            x_host = {h: w.x_dev[h].cpu() for h in w.sweep}
            r["parity_all_ranks"], r["parity_rows_checked_per_rank"] = w.parity_all_ranks(O, w.c_last, x_host)
        rec[name] = r
        w.free()
    if not clustered and world > 1:
        # single-GPU reference of the speed-up: rank 0 alone on the whole graph, the other ranks wait
        if rank == 0:
            w1 = SweepWorkload(a, shape, dev, 0, 1)
            ms1, _ = w1.timed(steps, warmup)
            rec["n1_same_box"] = {"value": w1.flops_step() / (ms1 * 1e-3) / 1e9, "ms_per_step": ms1}
            rec["speedup_with_exchange"] = rec["sharded"]["value"] / rec["n1_same_box"]["value"]
            rec["speedup_without_exchange"] = rec["sharded"]["no_exchange"]["value"] / rec["n1_same_box"]["value"]
            w1.free()
        dist.barrier()
    rec["config"] = {"workload": "%s-shape%s FLT32 CSR SpMM, hidden sweep %s" % (
        shape, " with block-model communities (1024 nodes, 70 %% of a row's edges inside)" if clustered else "",
        "/".join(map(str, HIDDEN_SWEEP))), "n_gpus": world}
    return rec


# ====================================================================================== inference workload
def run_inference(a):
    """BASELINE.json configs[3]: 2-layer GCN / GIN / SAGE end to end on the Reddit-shaped graph, hidden 128 -
    GPU aggregation (libbackend_pim.so) + torch Linear/BatchNorm on the GPU.  The reference's `--version=cpu`
    run (aggregation = row-parallel CSR SpMM on the host cores, dense layers = torch CPU) is timed beside it."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from pygim_b200 import graphgen
    from pygim_b200.backend_pim import pim_ops
    from pygim_b200.backend_pim.spmm import TORCH_TYPES, prepare_pim_spmm
    from pygim_b200.models import GCN, GIN, SAGE
    torch.cuda.set_device(0)
    dtype = TORCH_TYPES[a.dtype]
    hidden = a.hidden[0] if a.hidden else 128
    n, nnz, max_deg = graphgen.SHAPES[a.shape]
    rowptr, col = graphgen.synthetic_csr(n, nnz, max_deg, seed=0, device="cuda")
    from pygim_b200.sparse_tensor import SparseTensor
    adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(n, n), is_sorted=True)
    pim_ops.dpu_init_ranks(1)
    info = pim_ops.device_info()
    ns = make_args(hidden, dtype, a.format)
    ns.ds_parts = a.ds_parts if a.ds_parts > 0 else auto_ds_parts(n, hidden, info, torch.empty((), dtype=dtype).element_size())
    A = prepare_pim_spmm(adj, ns)
    feats, classes = 602, 41                                      # Reddit's feature / class counts
    torch.manual_seed(0)
    x = torch.randn(n, feats, device="cuda")
    from pygim_b200.models import layers
    results = {}
    for name, net in (("gcn", GCN), ("gin", GIN), ("sage", SAGE)):
        model = net(feats, hidden, classes, 2).cuda().eval()
        rec = {}
        outs = {}
        for key, fused in (("gpu_infer_ms", True), ("gpu_infer_ms_unfused_epilogue", False)):
            # fused: quantise = 2 kernels, de-quantise (+ GIN's (1+eps) x) inside the SpMM's row store;
            # unfused: the reference's torch expressions around A.mul (models/quantize.py:20-42)
            layers.FUSED_EPILOGUE = fused
            with torch.no_grad():
                for _ in range(max(a.warmup, 3)):
                    model(x, A)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    y = model(x, A)
                e1.record()
                torch.cuda.synchronize()
            rec[key] = e0.elapsed_time(e1) / a.steps
            outs[key] = y
        layers.FUSED_EPILOGUE = True
        rec["finite"] = bool(torch.isfinite(y).all())
        rec["fused_equals_unfused"] = bool(torch.equal(outs["gpu_infer_ms"], outs["gpu_infer_ms_unfused_epilogue"]))
        results[name] = rec
    cpu = None
    if not a.no_cpu:
        from oracle import oracle as O
        O.build()
        native = O.build_native()
        clib = O.lib(native) if native else O.lib()
        threads = O.max_threads()
        rp_h, cl_h = rowptr.cpu().numpy().astype("int32"), col.cpu().numpy().astype("int32")

        class CpuAdj:      # the torch_sparse.matmul branch of the reference's conv layers (float, 2^19 grid)
            dtype = torch.float

            def mul(self, xq):
                return torch.from_numpy(O.spmm_csr_rowpar(rp_h, cl_h, None, xq.numpy(), nthreads=threads, clib=clib))

        torch.set_num_threads(threads)
        xc = x.cpu()
        cpu = {"cores": threads, "kind": "port", "unit": "ms",
               "sample": "one full forward pass per model (whole Reddit-shaped graph), after one warm-up pass"}
        for name, net in (("gcn", GCN), ("gin", GIN), ("sage", SAGE)):
            model = net(feats, hidden, classes, 2).eval()
            with torch.no_grad():
                model(xc, CpuAdj())
                t0 = time.perf_counter()
                model(xc, CpuAdj())
                cpu[name] = (time.perf_counter() - t0) * 1e3
    total = sum(r["gpu_infer_ms"] for r in results.values())
    line = {"metric": "infer_ms_gcn+gin+sage", "value": total, "unit": "ms", "n_gpus": 1, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": total, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": "inference.py 2-layer GCN/GIN/SAGE, %s-shape, hidden %d, %s %s aggregation"
                                   % (a.shape, hidden, a.dtype, a.format), "nodes": n, "edges": nnz},
            "per_model": results, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default=SHAPE, choices=["reddit", "products", "arxiv"])
    ap.add_argument("--hidden", type=int, nargs="*", default=None, help="override the hidden sweep")
    ap.add_argument("--dtype", default="FLT32", choices=["INT8", "INT16", "INT32", "INT64", "FLT32", "DBL64"])
    ap.add_argument("--format", default="CSR", choices=["CSR", "COO"])
    ap.add_argument("--cpu-sample-rows", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs: one untimed-quality e2e pass only")
    ap.add_argument("--ds-parts", type=int, default=0, help="dense column parts per launch group; 0 = automatic")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N > 1: all-gather fused into the kernel epilogue (peer stores) or a separate NCCL collective")
    ap.add_argument("--no-multicast", action="store_true", help="fused gather: per-peer stores instead of multimem.st")
    ap.add_argument("--chunks", type=int, default=1, help="N > 1: sub-blocks per rank (all-gather/compute overlap)")
    ap.add_argument("--short-rows", type=int, default=None, choices=[0, 1, 2, 3, 4],
                    help="force the CSR instantiation: 0 deep unroll, 1 high occupancy, 2 streamed row tickets")
    ap.add_argument("--no-l2-persist", action="store_true",
                    help="do not put the access-policy window (persisting L2) over the dense tile")
    ap.add_argument("--general-kernel", action="store_true",
                    help="do not use the unit-value fast path (the adjacency of the benchmark is value-less => ones)")
    ap.add_argument("--sync", default="flags", choices=["flags", "barrier"],
                    help="fused gather: in-kernel arrival flags (no barrier) or one symmetric-memory barrier per call")
    ap.add_argument("--clustered", action="store_true",
                    help="headline graph with block-model communities (same N, nnz, degrees) instead of uniform columns")
    ap.add_argument("--reorder", default=None, choices=["cluster", "tiles", "degree"],
                    help="prepare-time row reordering (tiles = cluster + hot/cold shared-memory tiles)")
    ap.add_argument("--hot-k", type=int, default=1280, help="tile rows of the hot/cold plan")
    ap.add_argument("--tile-super-nnz", type=int, default=65536, help="nonzeros per superticket of the hot/cold plan")
    ap.add_argument("--opt", action="append", help="plan option key=value (pygim_plan_set_option), repeatable")
    ap.add_argument("--e2e-mode", default="pipelined", choices=["pipelined", "per-call"],
                    help="N = 1 host-operand step: one pipeline over the sweep (spmm_run_dense_many) or four host calls")
    ap.add_argument("--no-selftest", action="store_true", help="N > 1: skip the whole-matrix multi-GPU self-test")
    ap.add_argument("--no-clustered", action="store_true", help="skip the clustered-graph sub-record (N = 1)")
    ap.add_argument("--no-products", action="store_true", help="skip the products-shape sub-record")
    ap.add_argument("--no-arxiv", action="store_true", help="skip the arxiv-shape sub-record (N = 1)")
    ap.add_argument("--no-column-sharded", action="store_true", help="skip the column-sharded sub-record (N > 1)")
    ap.add_argument("--workload", default="spmm", choices=["spmm", "inference"],
                    help="spmm = the headline hidden sweep; inference = 2-layer GCN/GIN/SAGE end to end (configs[3])")
    a = ap.parse_args()
    if a.workload == "inference" and a.impl == "ours":
        run_inference(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
