"""Generates tests/golden/*.npz from the REFERENCE's own host oracles.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

Every case stores seeded inputs together with the output of the reference's scalar host loops
compiled in place by oracle/build_ref.sh (spmm_host_coo: spmm_default/spmm_mul_coo.c:40-51;
spmm_host_csr with values: spmm_grande/spmm_mul_csr.c:119-136; spmm_host: spmv_sparseP/spmv_mul_coo.c:92-103;
add_2D: spmm_default/spmm_mul_csr.c).  The fixtures travel to the GPU box, where /root/reference does
not exist; tests/test_golden.py checks the oracle restatement and the CUDA path against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

DTYPES = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64]


def random_sparse(rng, n, m, density, dtype, full_range=False, empty_rows=(), long_row=None):
    mask = rng.random((n, m)) < density
    for r in empty_rows:
        mask[r] = False
    if long_row is not None:
        mask[long_row] = rng.random(m) < 0.85
    row, col = np.nonzero(mask)                      # row-major sorted, unique
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        lo, hi = (max(info.min, -2 ** 31), min(info.max, 2 ** 31 - 1)) if full_range else (-7, 8)
        val = rng.integers(lo, hi, row.shape[0]).astype(dtype)
    else:
        val = rng.integers(-7, 8, row.shape[0]).astype(dtype)       # integer-valued => order-independent sums
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(row, minlength=n), out=rowptr[1:])
    return row.astype(np.int32), col.astype(np.int32), val, rowptr


def dense(rng, n, h, dtype, full_range=False):
    if np.issubdtype(dtype, np.integer) and full_range:
        info = np.iinfo(dtype)
        return rng.integers(max(info.min, -2 ** 31), min(info.max, 2 ** 31 - 1), (n, h)).astype(dtype)
    return rng.integers(-8, 4, (n, h)).astype(dtype)               # spmm_test.py:70


def main():
    assert O.ref_available(), "oracle/_ref missing: run oracle/build_ref.sh (needs /root/reference)"
    out = {}
    rng = np.random.default_rng(20240217)
    for dt in DTYPES:
        name = np.dtype(dt).name
        for case, (n, m, h, dens, full) in {"small": (37, 29, 5, 0.2, False), "wide": (64, 96, 33, 0.1, False),
                                             "wrap": (48, 48, 16, 0.6, True)}.items():
            row, col, val, rowptr = random_sparse(rng, n, m, dens, dt, full_range=full, empty_rows=(0, n - 1),
                                                  long_row=3)
            x = dense(rng, m, h, dt, full_range=full)
            key = "%s_%s" % (name, case)
            out[key + "_row"], out[key + "_col"], out[key + "_val"] = row, col, val
            out[key + "_rowptr"], out[key + "_x"] = rowptr, x
            out[key + "_n"] = np.array([n, m, h])
            out[key + "_y_coo"] = O.ref_spmm_host_coo(row, col, val, x, n, "default")
            out[key + "_y_csr"] = O.ref_spmm_host_csr(rowptr, col, val, x, "grande")
            out[key + "_y_spmv"] = O.ref_spmm_host_coo(row, col, val, x, n, "spmv")
            out[key + "_y_csr_ones"] = O.ref_spmm_host_csr(rowptr, col, val, x, "default")
            # padded x stride (grande: ncols < ncols_pad)
            xp = np.concatenate([x, dense(rng, m, 3, dt)], axis=1)
            out[key + "_xpad"] = xp
            out[key + "_y_csr_pad"] = O.ref_spmm_host_csr(rowptr, col, val, xp, "grande", ncols=h)
        # group composition (ops.hpp:42-62): 3 sparse column parts x 2 dense parts, via the reference's
        # spmm_host_coo + add_2D exactly as spmm_host_coo_group chains them
        n, m, h = 41, 50, 12
        row, col, val, rowptr = random_sparse(rng, n, m, 0.15, dt)
        x = dense(rng, m, h, dt)
        widths = [17, 17, 16]
        hs = [7, 5]
        y = np.zeros((n, h), dtype=dt)
        cur_row = 0
        for w in widths:
            sel = (col >= cur_row) & (col < cur_row + w)
            cur_col = 0
            for hj in hs:
                xb = np.ascontiguousarray(x[cur_row:cur_row + w, cur_col:cur_col + hj])
                yt = O.ref_spmm_host_coo(row[sel], col[sel] - cur_row, val[sel], xb, n, "default")
                O.ref_add_2d(y, yt, 0, cur_col)
                cur_col += hj
            cur_row += w
        key = "%s_group" % name
        out[key + "_row"], out[key + "_col"], out[key + "_val"], out[key + "_x"] = row, col, val, x
        out[key + "_n"] = np.array([n, m, h])
        out[key + "_widths"], out[key + "_hs"], out[key + "_y"] = np.array(widths), np.array(hs), y
    path = os.path.join(HERE, "reference_host_oracles.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
