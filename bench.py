#!/usr/bin/env python
"""bench.py - the aggregation hot path (sparse adjacency x dense features) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Reddit-shaped synthetic graph (232,965 nodes, 114,615,892 edges,
value-less adjacency => ones, features = randint(-8, 4) as spmm_test.py:70), FLT32 CSR SpMM, hidden-size
sweep 16/32/64/128.  One STEP = one pass of the hot path over the sweep (four SpMMs).

metric  : SpMM GFLOP/s = sum_H 2*nnz*H / time (whole job, all GPUs), inputs resident in HBM.
e2e     : the same step through the public API `prepare_pim_spmm(...).mul(x)` with HOST (pinned) operands:
          H2D of B and D2H of C inside the timed region.
roofline: HBM bound for the dominant kernel (the H=128 launch): algorithmic bytes
          4(N+1) + 4 nnz + s nnz + s N H + s N H / its CUDA-event duration, against
          MEASURED_PEAKS.json's hbm_gbs.
N > 1   : the adjacency is row-sharded by nnz across ranks (one process per GPU), B replicated, every
          rank computes its C row block and the blocks are all-gathered over NCCL; the collective is inside
          the timed region ("scaling": "strong").
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


def auto_ds_parts(n_cols, hidden, info, s=4):
    """Column tiling so that one B tile stays L2-resident next to the streaming A: see pygim_b200.utils.autotuner."""
    from pygim_b200.utils import autotuner
    return autotuner.choose_ds_parts(n_cols, hidden, s, info["l2_bytes"])


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
        "config": {"workload": "reddit-shape FLT32 CSR SpMM, hidden sweep 16/32/64/128", "nodes": n, "edges": nnz,
                   "hidden_sweep": HIDDEN_SWEEP, "format": "CSR"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "native_build": bool(native)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ====================================================================================== GPU arm
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

    from pygim_b200 import graphgen
    from pygim_b200.backend_pim import pim_ops
    from pygim_b200.backend_pim.spmm import SparseTensorCOO
    from pygim_b200.sparse_tensor import SparseTensor

    from pygim_b200.backend_pim.spmm import TORCH_TYPES
    dtype = TORCH_TYPES[a.dtype]
    esize = torch.empty((), dtype=dtype).element_size()
    n, nnz, max_deg = graphgen.SHAPES[a.shape]
    sweep = a.hidden if a.hidden else HIDDEN_SWEEP
    pim_ops.dpu_init_ranks(1)
    info = pim_ops.device_info()

    # ---- this rank's row shard (nnz-balanced, the GPU-level partition_by_nnz_csr)
    deg = graphgen.degree_sequence(n, nnz, max_deg, n, seed=0)
    full_rowptr = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(deg, 0, out=full_rowptr[1:])
    splits = pim_ops.partition_rows_by_nnz(full_rowptr, world) if world > 1 else [0, n]
    r0, r1 = splits[rank], splits[rank + 1]
    rowptr, col = graphgen.synthetic_csr(n, nnz, max_deg, seed=0, device=str(dev), rows=(r0, r1), deg=deg)
    shard_nnz = int(col.numel())
    adj = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(r1 - r0, n), is_sorted=True)
    del rowptr, col
    from pygim_b200.sharded import ShardedSpMM
    ds_parts = {h: (a.ds_parts if a.ds_parts > 0 else auto_ds_parts(n, h, info, esize)) for h in sweep}

    def args_for(h):
        ns = make_args(h, dtype, a.format)
        ns.ds_parts = ds_parts[h]
        return ns

    if world == 1:
        # one set of int32 CSR/COO arrays shared by the four plans (one plan per hidden size)
        base = SparseTensorCOO(adj, dtype=dtype, format=a.format)
        base.build_csr() if a.format == "CSR" else base.build_coo()
        plans = {}
        for h in sweep:
            A = copy.copy(base)
            A.sp_info_ptr = None
            A.to_pim_group(h, ds_parts[h])
            plans[h] = A
        ops = {h: ShardedSpMM(None, args_for(h), splits=splits, local_adj=adj,
                              make_local=lambda _adj, _args, _h=h: plans[_h]) for h in sweep}
    else:
        ops = {h: ShardedSpMM(None, args_for(h), splits=splits, local_adj=adj, chunks=a.chunks,
                              fused=(a.gather == "fused"), use_multicast=not a.no_multicast) for h in sweep}
        plans = {h: ops[h].locals[0] for h in sweep}
        if a.gather == "fused":
            # the fused path needs NVLink peer mappings (symmetric memory); if this box cannot provide them, every
            # rank falls back to the NCCL all-gather together and the JSON line says so
            ok = torch.ones(1, device=dev)
            try:
                probe = torch.zeros((n, sweep[0]), dtype=dtype, device=dev)
                ops[sweep[0]].mul(probe)
                torch.cuda.synchronize()
            except Exception as exc:      # noqa: BLE001
                ok.zero_()
                if rank == 0:
                    print("bench.py: fused all-gather unavailable (%s); using NCCL" % str(exc)[:200], file=sys.stderr)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok) == 0.0:
                a.gather = "nccl"
                for h in sweep:
                    ops[h].fused = False
    if a.general_kernel:     # force the weighted kernels although the adjacency is value-less (all ones)
        for h in sweep:
            for op in ops[h].locals:
                pim_ops.plan_set_option(op.sp_info_ptr, "unit_values", 0)
    if a.short_rows is not None:
        for h in sweep:
            for op in ops[h].locals:
                pim_ops.plan_set_option(op.sp_info_ptr, "short_rows", a.short_rows)
    if a.no_l2_persist:
        for h in sweep:
            for op in ops[h].locals:
                pim_ops.plan_set_option(op.sp_info_ptr, "l2_persist", 0)
    x_dev = {h: graphgen.reference_features(n, h, dtype, seed=h, device=str(dev)) for h in sweep}
    # full outputs (every rank ends with all rows, ready for the next layer)
    c_full = {h: torch.empty((n, h), dtype=dtype, device=dev) for h in sweep}

    c_last = dict(c_full)

    def step_device(record=None):
        for i, h in enumerate(sweep):
            if record is not None:
                record[i][0].record()
            if world > 1 and a.gather == "fused":
                c_last[h] = ops[h].mul(x_dev[h])       # rows land in every peer's symmetric buffer (no collective)
            else:
                ops[h].mul(x_dev[h], out=c_full[h])    # N > 1: SpMM of the row block(s) + NCCL all-gather
            if record is not None:
                record[i][1].record()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.1 s to deliver its first sample: start before the warm-up
    for _ in range(max(a.warmup, 3)):
        step_device()
    sync()
    sampler.mark()               # report only samples taken from here (timed region + e2e region) on
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in sweep]
          for _ in range(a.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t_begin.record()
    for k in range(a.steps):
        step_device(ev[k])
    t_end.record()
    sync()
    elapsed_ms = t_begin.elapsed_time(t_end)
    launches_per_step = sum(pim_ops.last_launches(op.sp_info_ptr) for h in sweep for op in ops[h].locals)
    per_h_ms = [sum(ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(a.steps)) / a.steps
                for i in range(len(sweep))]
    if world > 1:
        t = torch.tensor([elapsed_ms] + per_h_ms, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, per_h_ms = float(t[0]), [float(v) for v in t[1:]]
    flops_step = sum(2.0 * nnz * h for h in sweep)
    value = flops_step * a.steps / (elapsed_ms * 1e-3) / 1e9

    # ---- e2e: same step through the public API with pinned HOST operands (H2D + D2H inside the timing)
    x_host = {h: x_dev[h].cpu().pin_memory() for h in sweep}
    c_host = {h: torch.empty((r1 - r0, h), dtype=dtype).pin_memory() for h in sweep}

    def step_host():
        for h in sweep:
            if world == 1:
                plans[h].mul(x_host[h], out=c_host[h])
            else:   # each rank's row block(s) through the host entry point; no collective on host results
                blk = ops[h]
                for k, op in enumerate(blk.locals):
                    op.mul(x_host[h], out=c_host[h][blk.sub[k]:blk.sub[k + 1]])

    e2e_steps = max(3, min(a.steps, 10)) if not a.no_e2e else 1
    for _ in range(2 if not a.no_e2e else 0):
        step_host()
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    sync()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = flops_step * e2e_steps / e2e_s / 1e9
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(n * h * esize for h in sweep)
    d2h = sum((r1 - r0) * h * esize for h in sweep)
    timers = {h: pim_ops.last_timers(plans[h].sp_info_ptr) for h in sweep}

    # ---- parity spot check of what was just timed (rank 0, first rows, against the oracle)
    parity = None
    if rank == 0 and not a.no_check:
        import numpy as np
        from oracle import oracle as O
        O.build()
        rows_chk = min(256, r1 - r0)
        rp, cl, _ = adj.csr()
        rp_h = rp[: rows_chk + 1].cpu().numpy().astype("int32")
        cl_h = cl[: int(rp_h[-1])].cpu().numpy().astype("int32")
        parity = True
        for h in sweep:
            want = O.spmm_csr_rowpar(rp_h, cl_h, None, x_host[h].numpy(), nthreads=O.max_threads())
            parity = parity and bool(np.array_equal(want, c_last[h][r0:r0 + rows_chk].cpu().numpy())) \
                and bool(np.array_equal(want, c_host[h][:rows_chk].numpy()))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of the step)
    peak, peak_kind = measured_peaks()
    per_hidden = []
    for i, h in enumerate(sweep):
        # per-GPU algorithmic bytes of this rank's SpMM (shard of A and C, all of B), A counted once
        b = alg_bytes_csr(r1 - r0, n, shard_nnz, h, esize, a.format)
        per_hidden.append({"hidden": h, "kernel_ms": per_h_ms[i], "gflops": 2.0 * shard_nnz * h / per_h_ms[i] / 1e6,
                           "alg_gbs": b / per_h_ms[i] / 1e6, "frac_hbm": b / per_h_ms[i] / 1e6 / peak,
                           "gather_gbs": float(esize) * shard_nnz * h / per_h_ms[i] / 1e6,
                           "launches": ds_parts[h], "tile_cols": h // ds_parts[h]})
    # The dominant KERNEL is the template instantiation with the largest share of the step.  An instantiation
    # is fixed by the dense tile a launch handles (G = lanes per tile row), so hidden sizes that run as
    # several column tiles (H = 128 -> two 64-column launches) are launches of the SAME kernel as H = 64.
    # Per launch the algorithmic bytes are A + that tile of B + that tile of C (SURVEY.md 8d with H = tile).
    by_kernel = {}
    for i, h in enumerate(sweep):
        w = h // ds_parts[h]
        k = by_kernel.setdefault(w, {"ms": 0.0, "launches": 0, "bytes": 0.0, "hidden": []})
        k["ms"] += per_h_ms[i]
        k["launches"] += ds_parts[h]
        k["bytes"] += ds_parts[h] * alg_bytes_csr(r1 - r0, n, shard_nnz, w, esize, a.format)
        k["hidden"].append(h)
    dom_w = max(by_kernel, key=lambda w: by_kernel[w]["ms"])
    dk = by_kernel[dom_w]
    lanes = max(1, min(32, dom_w * esize // 16))
    roof = {"bound": "hbm",
            "kernel": "%s_spmm_kernel<%s, G=%d> (%d-column tiles; hidden %s)"
                      % (a.format.lower(), a.dtype, lanes, dom_w, "/".join(map(str, dk["hidden"]))),
            "achieved": dk["bytes"] / dk["ms"] / 1e6, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
            "frac": dk["bytes"] / dk["ms"] / 1e6 / peak, "traffic": None,
            "alg_bytes_per_launch": dk["bytes"] / dk["launches"], "launches_per_step": dk["launches"],
            "avg_launch_ms": dk["ms"] / dk["launches"], "share_of_step": dk["ms"] / sum(per_h_ms),
            "sweep_achieved": sum(alg_bytes_csr(r1 - r0, n, shard_nnz, h, esize, a.format) for h in sweep)
            / sum(per_h_ms) / 1e6,
            "note": "Reddit-shape is L2-gather bound (s*nnz*H bytes leave L2 per launch), see DESIGN.md 4.3"}
    # secondary bound: the measured ceiling of random row gathers (tools/l2_gather_probe)
    try:
        with open(os.path.join(ROOT, "profiles", "gather_ceiling.json")) as f:
            ceil = json.load(f)
        row_bytes = dom_w * esize
        key = str(min((64, 128, 256, 512), key=lambda b: abs(b - row_bytes)))
        resident = n * row_bytes <= 0.46 * info["l2_bytes"]
        peak_g = ceil["l2_resident_tbs" if resident else "hbm_served_tbs"][key]
        ach_g = float(esize) * shard_nnz * sum(h for h in dk["hidden"]) / dk["ms"] / 1e9
        roof["gather"] = {"bound": "l2-gather" if resident else "hbm-gather", "achieved_tbs": ach_g,
                          "ceiling_tbs": peak_g, "frac": ach_g / peak_g, "row_bytes": row_bytes,
                          "what": "s*nnz*H bytes gathered per launch / time, against the measured ceiling of random "
                                  "row gathers (profiles/r01_l2_gather_probe.txt)"}
    except Exception:
        pass
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path) and (a.shape, a.dtype, a.format, world) == ("reddit", "FLT32", "CSR", 1):
        try:
            with open(traffic_path) as f:
                roof["traffic"] = json.load(f).get("tile_%d" % dom_w)
        except Exception:
            pass

    # ---- CPU baseline beside it (rank 0, N == 1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not a.no_cpu:
        from oracle import oracle as O
        O.build()
        native = O.build_native()
        clib = O.lib(native) if native else O.lib()
        threads = O.max_threads()
        sample_rows = min(a.cpu_sample_rows or n, r1 - r0)         # default: the whole workload
        rp, cl, _ = adj.csr()
        rp_h = rp[: sample_rows + 1].cpu().numpy().astype("int32")
        cl_h = cl[: int(rp_h[-1])].cpu().numpy().astype("int32")
        xs = {h: x_host[h].numpy() for h in sweep}
        secs, fl = cpu_spmm_sample(O, clib, rp_h, cl_h, xs, threads, repeats=3)
        cpu = {"value": fl / secs / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "rows [0,%d) of the same graph (%d nnz), hidden sweep %s, best of 3 passes (%.2f s each)"
                         % (sample_rows, int(rp_h[-1]), sweep, secs), "native_build": bool(native)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {"FLT32": "f32", "DBL64": "f64", "INT8": "i8", "INT16": "i16", "INT32": "i32", "INT64": "i64"}[a.dtype],
        "data": "synthetic",
        "config": {"workload": "%s-shape %s %s SpMM, hidden sweep %s" % (a.shape, a.dtype, a.format,
                                                                         "/".join(map(str, sweep))),
                   "nodes": n, "edges": nnz, "hidden_sweep": sweep, "format": a.format, "sp_parts": 1,
                   "values": "all ones (value-less adjacency); " + ("general weighted kernel forced" if a.general_kernel
                             else "unit-value fast path: value stream not read, results bit-identical"),
                   "ds_parts": {str(h): ds_parts[h] for h in sweep},
                   "sharding": ("rows by nnz over %d GPUs, B replicated, all-gather of C %s, inside the timing"
                                % (world, "fused into the kernel epilogue (NVLink peer stores)" if a.gather == "fused"
                                   else "by NCCL (%d sub-blocks per rank)" % a.chunks))
                   if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (A = %.0f MB streams through a %.0f MB L2 every launch)"
                         % ((8.0 * nnz) / 1e6, info["l2_bytes"] / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                "phases_ms": {str(h): timers[h] for h in sweep}},
        "gpu_launches": launches_per_step * a.steps,
        "roofline": roof, "per_hidden": per_hidden, "cpu_baseline": cpu, "clocks": clocks,
        "parity_spot_check": parity,
        "lib": os.path.relpath(__import__("pygim_b200._lib", fromlist=["x"]).loaded_path() or "", ROOT),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    results = {}
    for name, net in (("gcn", GCN), ("gin", GIN), ("sage", SAGE)):
        model = net(feats, hidden, classes, 2).cuda().eval()
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
        results[name] = {"gpu_infer_ms": e0.elapsed_time(e1) / a.steps, "finite": bool(torch.isfinite(y).all())}
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
    ap.add_argument("--short-rows", type=int, default=None, choices=[0, 1, 2],
                    help="force the CSR instantiation: 0 deep unroll, 1 high occupancy, 2 streamed row tickets")
    ap.add_argument("--no-l2-persist", action="store_true",
                    help="do not put the access-policy window (persisting L2) over the dense tile")
    ap.add_argument("--general-kernel", action="store_true",
                    help="do not use the unit-value fast path (the adjacency of the benchmark is value-less => ones)")
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
