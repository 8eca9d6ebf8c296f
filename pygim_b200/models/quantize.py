"""Symmetric quantisation around the aggregation (models/quantize.py:20-42 of the reference).

The grid depends on the aggregation dtype: int8 -> |x_q| <= 16, int16 -> <= 512, int32 -> <= 2^19; any other
dtype (FLT32 / DBL64) rounds to the +-2^19 grid but stays torch.float.  All of it is elementwise torch work on
whatever device `v` lives on."""
from __future__ import annotations

import torch

_GRID_BITS = {torch.int8: 5, torch.int16: 10, torch.int32: 20}


def symmetric_quantize(v: torch.Tensor, dtype: torch.dtype = torch.int32):
    """Returns (scale, round(v / scale) as `dtype`) with scale = 2 * max|v| / 2^bits."""
    bits = _GRID_BITS.get(dtype)
    if bits is None:
        bits, dtype = 20, torch.float
    scale = v.abs().max() * 2 / float(2 ** bits)
    q = torch.round(v / scale)
    q = q.clone() if q.dtype == dtype else q.to(dtype)
    return scale, q


def symmetric_dequantize(out: torch.Tensor, scale_edge, scale_x) -> torch.Tensor:
    return out * (scale_edge * scale_x)
