"""Retargeted autotuner (utils/autotuner.py of the reference): picks the (sp_parts, ds_parts) split and
kernel parameters from graph statistics instead of UPMEM bandwidth tables.  [first cut: column tiling]"""
from __future__ import annotations


def choose_ds_parts(n_cols: int, hidden: int, elem_size: int, l2_bytes: int, l2_fraction: float = 0.46) -> int:
    """Smallest number (1, 2 or 4) of equal column tiles for which one B tile (n_cols x hidden/ds x elem_size) fits in
    `l2_fraction` of L2.  B200's L2 is two die-local halves and read-shared data ends up in both, so the
    budget for the resident tile is well below the nominal 126 MB; the streaming A/C traffic needs room too.
    Tiles keep 16-byte rows (hidden/ds * elem_size % 16 == 0) so the vector kernels stay usable."""
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
