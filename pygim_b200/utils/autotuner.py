"""Retargeted autotuner: picks the (sp_parts, ds_parts) split and the kernel scheduling parameters from graph
statistics and a small B200 cost model.

The reference's `autotune(datadir, dataset, hidden_size, split_set, blnc_set)` (utils/autotuner.py:263-343)
scores `load/HOST_DPU_BW + merge + max_nnz_per_dpu/FMA_THROUGHPUT + retrieve/DPU_HOST_BW` with UPMEM
micro-benchmark tables (:23-89) over `sp_ds_set = [(1,32),(2,16)]` x balance and returns
`[sp_parts, ds_parts, balance, balance_tsklt, None]`.  Here the same shape of answer comes from:

* graph statistics (rows, nnz, mean / max / CV of the row degree) - `GraphStats`;
* measured B200 constants (`DeviceModel`): HBM copy bandwidth, the L2->SM gather rate as a function of the
  gathered row's bytes, the HBM random-row efficiency when B does not fit L2, the L2-resident capacity;
* an analytic time per (sp_parts, ds_parts): every dense tile re-streams A from HBM and gathers nnz rows of
  `w*s` bytes either out of L2 (tile resident) or out of HBM; sparse parts >= 1 add a read-modify-write of C.

`choose_ds_parts` is the closed-form special case used by the front-ends when the caller passes ds_parts=0.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from .space import For, Space, Table

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@dataclass
class DeviceModel:
    """B200 constants, measured with bench.py / ncu in round 1 (profiles/r01_*.json)."""
    hbm_gbs: float = 6541.0
    l2_bytes: int = 126 * 2 ** 20
    l2_resident_fraction: float = 0.46          # read-shared data is held in both die-local L2 halves
    sm_count: int = 148
    launch_us: float = 4.0
    # a row is a dependent chain (index load -> gather -> shuffle tree -> store) that one warp walks alone:
    # short-row graphs are bound by rows x row_latency / resident warps (measured on products-shape, 16-byte rows)
    row_latency_us: float = 1.3          # with streamed row tickets (1.9 without)
    resident_warps_per_sm: int = 32
    # gather rate out of L2 (TB/s of gathered payload) vs bytes per gathered row
    l2_gather_tbs: Dict[int, float] = field(default_factory=lambda: {16: 3.2, 32: 6.5, 64: 13.0, 128: 19.0,
                                                                     256: 19.8, 512: 19.5})
    # fraction of hbm_gbs reached when the rows are gathered from HBM (B >> L2)
    hbm_gather_eff: Dict[int, float] = field(default_factory=lambda: {16: 0.10, 32: 0.20, 64: 0.55, 128: 0.85,
                                                                      256: 0.97, 512: 1.00})

    @classmethod
    def from_environment(cls, info: Optional[dict] = None) -> "DeviceModel":
        m = cls()
        path = os.path.join(_ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(path):
            try:
                with open(path) as f:
                    m.hbm_gbs = float(json.load(f)["hbm_gbs"])
            except Exception:
                pass
        if info:
            m.l2_bytes = int(info.get("l2_bytes", m.l2_bytes))
            m.sm_count = int(info.get("sm_count", m.sm_count))
        return m

    @staticmethod
    def _interp(table: Dict[int, float], row_bytes: float) -> float:
        keys = sorted(table)
        if row_bytes <= keys[0]:
            return table[keys[0]] * row_bytes / keys[0]
        if row_bytes >= keys[-1]:
            return table[keys[-1]]
        for lo, hi in zip(keys, keys[1:]):
            if lo <= row_bytes <= hi:
                t = (math.log2(row_bytes) - math.log2(lo)) / (math.log2(hi) - math.log2(lo))
                return table[lo] + t * (table[hi] - table[lo])
        return table[keys[-1]]


@dataclass
class GraphStats:
    nrows: int
    ncols: int
    nnz: int
    mean_degree: float
    max_degree: int
    cv_degree: float          # coefficient of variation (degree skew)
    empty_rows: int

    @classmethod
    def from_rowptr(cls, rowptr, ncols: Optional[int] = None) -> "GraphStats":
        import torch
        rp = rowptr.to(torch.int64).cpu()
        deg = (rp[1:] - rp[:-1]).to(torch.float64)
        n = int(deg.numel())
        mean = float(deg.mean()) if n else 0.0
        std = float(deg.std(unbiased=False)) if n else 0.0
        return cls(n, int(ncols if ncols is not None else n), int(rp[-1]) if n else 0, mean,
                   int(deg.max()) if n else 0, std / mean if mean > 0 else 0.0, int((deg == 0).sum()))


def choose_ds_parts(n_cols: int, hidden: int, elem_size: int, l2_bytes: int, l2_fraction: float = 0.46,
                    nnz: Optional[int] = None) -> int:
    """Smallest number (1, 2 or 4) of equal column tiles for which one B tile (n_cols x hidden/ds x elem_size)
    fits in `l2_fraction` of L2.  B200's L2 is two die-local halves and read-shared data ends up in both, so the
    budget for the resident tile is well below the nominal 126 MB; the streaming A/C traffic needs room too.
    Tiling buys L2 hits for rows that are gathered MANY times; when a feature row is gathered only a few times per
    launch (`nnz` / `n_cols` < 32: citation-shaped graphs) there is nothing to keep resident and every extra tile
    only repeats the launches (arxiv-shape H = 128: 136 us as two tiles, 105 us as one)."""
    if nnz is not None and nnz < 32 * max(n_cols, 1):
        return 1
    budget = l2_bytes * l2_fraction
    for ds in (1, 2, 4):
        if hidden % ds:
            continue
        w = hidden // ds
        if ds > 1 and (w * elem_size) % 128:
            continue      # a tile row must stay whole 128-byte lines, or the gather wastes sectors
        if n_cols * w * elem_size <= budget:
            return ds
    # B is far larger than L2 (ogbn-products-shape): the gather is served by HBM whatever the tiling, and every
    # extra tile re-streams A and shortens the gathered rows (measured: 5.6 ms at ds=1, 12 ms at ds=2).
    return 1


def predict_ms(stats: GraphStats, hidden: int, elem_size: int, sp_parts: int, ds_parts: int,
               dev: Optional[DeviceModel] = None, fmt: str = "CSR") -> float:
    """Analytic time of one SpMM under the (sp_parts, ds_parts) tiling."""
    dev = dev or DeviceModel()
    w = -(-hidden // ds_parts)                              # ceil: widest dense tile
    row_bytes = w * elem_size
    idx_bytes = (4 if fmt == "CSR" else 8) + elem_size      # per nonzero: (rowind) + colind + value
    cols_per_part = -(-stats.ncols // sp_parts)
    tile_bytes = cols_per_part * row_bytes
    resident = tile_bytes <= dev.l2_bytes * dev.l2_resident_fraction
    total_us = 0.0
    for sp in range(sp_parts):
        nnz = stats.nnz / sp_parts
        for _ in range(ds_parts):
            stream_us = (nnz * idx_bytes + 4 * stats.nrows + tile_bytes * (1 if resident else 0)
                         + stats.nrows * row_bytes * (1 if sp == 0 else 3)) / (dev.hbm_gbs * 1e3)
            if resident:
                gather_us = nnz * row_bytes / (DeviceModel._interp(dev.l2_gather_tbs, row_bytes) * 1e6)
            else:
                gather_us = nnz * row_bytes / (dev.hbm_gbs * 1e3 * DeviceModel._interp(dev.hbm_gather_eff, row_bytes))
                stream_us += gather_us                       # the gather itself is HBM traffic
                gather_us = 0.0
            chain_us = stats.nrows * dev.row_latency_us / (dev.sm_count * dev.resident_warps_per_sm)
            total_us += max(stream_us, gather_us, chain_us) + dev.launch_us
    return total_us / 1e3


def default_space(hidden: int) -> Space:
    """(sp_parts, ds_parts) candidates: the reference's pairs rescaled to what a GPU needs (utils/autotuner.py:259-263
    uses [(1,32),(2,16)] because one UPMEM rank holds one B slice)."""
    return For("sp_parts", [1, 2, 4]) * For("ds_parts", [d for d in (1, 2, 4, 8) if hidden % d == 0])


def autotune(adj_or_stats, hidden_size, split_set=None,
             blnc_set: Sequence = (0, 2), elem_size: int = 4, fmt: str = "CSR",
             dev: Optional[DeviceModel] = None) -> List:
    """Returns `[sp_parts, ds_parts, balance, balance_tsklt, extra]` like the reference (utils/autotuner.py:338).

    `adj_or_stats`: a SparseTensor-like object (`.csr()`, `.size(1)`) or a GraphStats.  `split_set`: explicit
    (sp, ds) pairs (the reference's argument) or None for `default_space`.  On the GPU both balancing levels are
    nnz-based (long rows are cut into segments, work is ticketed), so `balance`/`balance_tsklt` are "nnz" unless
    the graph has no skew at all, where plain row tickets ("row") suffice.  `extra` carries the scheduling
    parameters and the predicted time."""
    if isinstance(adj_or_stats, str):
        # the reference's positional form: autotune(datadir, dataset, hidden_size, split_set, blnc_set)
        # (utils/autotuner.py:263; called as autotune(data_root, dataset, dense_size, sp_ds_set) at
        # utils/experiment.py:400)
        datadir, dataset, hidden_size = adj_or_stats, hidden_size, split_set
        split_set = blnc_set if blnc_set and isinstance(blnc_set[0], (tuple, list)) else None
        stats = load_dataset_stats(datadir, str(dataset))
    elif isinstance(adj_or_stats, GraphStats):
        stats = adj_or_stats
    else:
        stats = GraphStats.from_rowptr(adj_or_stats.csr()[0], adj_or_stats.size(1))
    dev = dev or DeviceModel.from_environment()
    space: Space = Table(["sp_parts", "ds_parts"], split_set) if split_set else default_space(hidden_size)
    best, best_ms = None, float("inf")
    for cfg in space.iter_dict():
        ms = predict_ms(stats, hidden_size, elem_size, cfg["sp_parts"], cfg["ds_parts"], dev, fmt)
        if ms < best_ms:
            best, best_ms = cfg, ms
    slots = dev.sm_count * 64
    seg_len = 512
    while seg_len < stats.nnz / max(1, slots * 8) and seg_len < 4096:
        seg_len *= 2
    skewed = stats.max_degree > seg_len or stats.cv_degree > 0.5
    balance = "nnz" if skewed else "row"
    extra = {"predicted_ms": best_ms, "seg_len": seg_len,
             "rows_per_ticket": int(max(1, min(31, 256 // max(1.0, stats.mean_degree)))),
             "options": kernel_options(stats, hidden_size, elem_size),
             "kernel": "csr-light (64 regs)" if stats.mean_degree < 96 else "csr-deep (128 regs) + segments"}
    return [best["sp_parts"], best["ds_parts"], balance, "nnz", extra]


# ====================================================================================== plan tuning (in the loop)
# The reference consumes autotune() where the experiment is assembled (utils/experiment.py:398-401: the tuned
# [sp_parts, ds_parts, balance, balance_tsklt] become the CLI of the run).  Here the consumer is
# prepare_pim_spmm(..., args.tune / ds_parts == 0): the pick becomes the plan's column tiling and its kernel
# options, and is persisted next to the dataset so the next run skips the search.
TUNE_FILE = "pygim_b200_tune.json"


def kernel_options(stats: GraphStats, hidden: int, elem_size: int, reordered: bool = False) -> Dict[str, int]:
    """Kernel variant + scheduling options (pygim_plan_set_option keys) from degree skew, nnz/row and row bytes."""
    opts: Dict[str, int] = {}
    short = stats.mean_degree < 96
    # kernel family: deep (128 registers, 16 gathers in flight) for long rows, light (64 registers, twice the warps)
    # for short rows; measured: Reddit-shape 9290 vs 8500 GFLOP/s, products-shape 2314 vs 2656
    # very short rows (citation graphs, mean degree < 12): the two-launch family - tiny rows by lane groups, the rest as
    # pieces (arxiv-shape H = 32: 36 vs 70 us)
    opts["short_rows"] = (4 if stats.mean_degree < 12 else 3) if short else 0
    # work items: ~256 nonzeros, fewer on small graphs so that every resident warp still gets about four items
    # (the library's own default, csrc/backend_pim.cu::build_plan_range)
    want, item = stats.nnz // (4 * 148 * 32), 64
    while item * 2 <= want and item < 256:
        item *= 2
    opts["item_nnz"] = item
    # (reordered plans: the library itself schedules SM-affine supertickets when the dense rows are >= 256 bytes.
    # Capping the lanes per row at 8 - one 128-byte L1 line per gathered row, so a community's rows fit the L1 -
    # was measured on the clustered Reddit-shape graph and LOSES: 1.79 vs 1.47 ms at H = 64.)
    return opts


def candidate_options(stats: GraphStats, hidden: int, elem_size: int, reordered: bool = False) -> List[Dict[str, int]]:
    """The (small) space a measured search walks: the analytic pick first, then its neighbours."""
    base = kernel_options(stats, hidden, elem_size, reordered)
    out = [dict(base)]
    for sr in (0, 2, 3, 4):
        if sr != base["short_rows"]:
            out.append({**base, "short_rows": sr})
    for item in (64, 128, 256, 512):
        if item != base["item_nnz"]:
            out.append({**base, "item_nnz": item})
    if hidden * elem_size > 128:
        out.append({**base, "max_g": 8} if "max_g" not in base else {k: v for k, v in base.items() if k != "max_g"})
    out.append({**base, "cta_threads": 1024})
    return out


def _graph_key(stats: GraphStats, hidden: int, dtype, fmt: str, reordered: bool, dataset: Optional[str]) -> str:
    name = dataset or "graph"
    return "%s|n=%d|m=%d|nnz=%d|maxdeg=%d|H=%d|%s|%s|reord=%d" % (
        name, stats.nrows, stats.ncols, stats.nnz, stats.max_degree, hidden, str(dtype).replace("torch.", ""), fmt,
        int(reordered))


def _load_cache(cache_dir: Optional[str]) -> dict:
    if not cache_dir:
        return {}
    try:
        with open(os.path.join(cache_dir, TUNE_FILE)) as f:
            return json.load(f)
    except Exception:
        return {}


def _store_cache(cache_dir: Optional[str], cache: dict) -> None:
    if not cache_dir:
        return
    try:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = os.path.join(cache_dir, TUNE_FILE + ".tmp")
        with open(tmp, "w") as f:
            json.dump(cache, f, indent=1, sort_keys=True)
        os.replace(tmp, os.path.join(cache_dir, TUNE_FILE))
    except OSError:
        pass


def measure_options(A, x, candidates: Sequence[Dict[str, int]], repeats: int = 5) -> List[float]:
    """Median CUDA-event time (ms) of A.mul(x) under each option set; the plan's options are restored to automatic."""
    import torch
    from ..backend_pim import pim_ops
    keys = sorted({k for c in candidates for k in c})

    def one(cand):
        for k in keys:
            pim_ops.plan_set_option(A.sp_info_ptr, k, cand.get(k, -1))
        for _ in range(3):
            A.mul(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(repeats):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            A.mul(x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    one(candidates[0])                                   # clocks and caches settle before anything is recorded
    times = [one(c) for c in candidates]
    times = [min(t, one(c)) for t, c in zip(times, candidates)]      # second pass in the same order: drift cancels
    for k in keys:
        pim_ops.plan_set_option(A.sp_info_ptr, k, -1)
    return times


def tune_plan(adj, hidden_size: int, dtype=None, fmt: str = "CSR", cache_dir: Optional[str] = None,
              dataset: Optional[str] = None, reordered: bool = False, dev: Optional[DeviceModel] = None) -> dict:
    """{"sp_parts", "ds_parts", "options", "predicted_ms", "source"} for one plan; persisted in
    `cache_dir/pygim_b200_tune.json` (the reference keeps its tuning inputs next to --datadir as well,
    utils/autotuner.py:386-419)."""
    import torch
    elem = torch.empty((), dtype=dtype or torch.float32).element_size()
    stats = adj if isinstance(adj, GraphStats) else GraphStats.from_rowptr(adj.csr()[0], adj.size(1))
    key = _graph_key(stats, hidden_size, dtype, fmt, reordered, dataset)
    cache = _load_cache(cache_dir)
    if key in cache:
        hit = dict(cache[key])
        hit["source"] = "cache"
        return hit
    if dev is None:
        info = None
        try:
            from ..backend_pim import pim_ops
            info = pim_ops.device_info()
        except Exception:
            pass
        dev = DeviceModel.from_environment(info)
    ds = choose_ds_parts(stats.ncols, hidden_size, elem, dev.l2_bytes, dev.l2_resident_fraction, nnz=stats.nnz)
    choice = {"sp_parts": 1, "ds_parts": ds, "options": kernel_options(stats, hidden_size, elem, reordered),
              "predicted_ms": predict_ms(stats, hidden_size, elem, 1, ds, dev, fmt), "source": "model"}
    cache[key] = {k: v for k, v in choice.items() if k != "source"}
    _store_cache(cache_dir, cache)
    return choice


def load_dataset_stats(datadir: str, dataset: str) -> GraphStats:
    """Graph statistics of a dataset directory as the reference's tuner reads it (utils/autotuner.py:386-419 loads
    `<datadir>/<dataset>` through PyG).  No dataset can be downloaded here: a `<dataset>.pt` / `<dataset>.npz` file
    holding rowptr (and ncols) is read if present, else the named synthetic shape is used."""
    import torch
    base = os.path.join(datadir, dataset)
    for ext in (".pt", ".npz"):
        if os.path.exists(base + ext):
            if ext == ".pt":
                blob = torch.load(base + ext)
                return GraphStats.from_rowptr(blob["rowptr"], int(blob.get("ncols", blob["rowptr"].numel() - 1)))
            import numpy as np
            blob = np.load(base + ext)
            return GraphStats.from_rowptr(torch.from_numpy(blob["rowptr"]), int(blob["ncols"]) if "ncols" in blob else None)
    from .. import graphgen
    shape = {"ogbn-arxiv": "arxiv", "Reddit": "reddit", "reddit": "reddit", "ogbn-products": "products",
             "PubMed": "pubmed"}.get(dataset, dataset)
    if shape not in graphgen.SHAPES:
        raise FileNotFoundError("no %s.pt/.npz under %s and no synthetic shape of that name" % (dataset, datadir))
    n, nnz, max_deg = graphgen.SHAPES[shape]
    deg = graphgen.degree_sequence(n, nnz, max_deg, n, seed=0)
    rp = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(deg, 0, out=rp[1:])
    return GraphStats.from_rowptr(rp, n)
