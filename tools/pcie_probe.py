#!/usr/bin/env python
"""PCIe rates the host entry points can count on: pinned host <-> device copies of a [233k x 128] float32 matrix as
one 1-D copy and as 2-D copies of 128 / 256 / 512-byte row pieces (the column tiles of pygim_spmm_run_many_host),
alone and with the opposite direction running at the same time."""
import ctypes as C
import sys

import torch

rt = C.CDLL("libcudart.so.12") if True else None
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
H2D, D2H = 1, 2
n, h = 232965, 128
host = torch.empty((n, h), dtype=torch.float32).pin_memory()
host2 = torch.empty((n, h), dtype=torch.float32).pin_memory()
dev = torch.empty((n, h), dtype=torch.float32, device="cuda")
dev2 = torch.empty((n, h), dtype=torch.float32, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def copy(kind, width_bytes, stream, a_host, a_dev):
    """the whole matrix, as full-row 1-D copy (width 0) or as column tiles of width_bytes"""
    row = h * 4
    if width_bytes == 0:
        args = (a_dev.data_ptr(), a_host.data_ptr()) if kind == H2D else (a_host.data_ptr(), a_dev.data_ptr())
        assert rt.cudaMemcpyAsync(args[0], args[1], n * row, kind, C.c_void_p(stream.cuda_stream)) == 0
        return
    for off in range(0, row, width_bytes):
        if kind == H2D:   # device tile is contiguous (as in the library), host is strided
            d, dp, s, sp = a_dev.data_ptr() + off * n, width_bytes, a_host.data_ptr() + off, row
        else:
            d, dp, s, sp = a_host.data_ptr() + off, row, a_dev.data_ptr() + off, row
        assert rt.cudaMemcpy2DAsync(d, dp, s, sp, width_bytes, n, kind, C.c_void_p(stream.cuda_stream)) == 0


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


gb = n * h * 4 / 1e9
for w in (0, 512, 256, 128, 64):
    a = timed(lambda: copy(H2D, w, s_in, host, dev))
    b = timed(lambda: copy(D2H, w, s_out, host2, dev2))
    both = timed(lambda: (copy(H2D, w, s_in, host, dev), copy(D2H, w, s_out, host2, dev2)))
    print("row piece %4s B:  H2D %5.1f GB/s   D2H %5.1f GB/s   both at once %5.1f GB/s each (%.2f ms)" %
          (w or "1-D", gb / a * 1e3, gb / b * 1e3, gb / both * 1e3, both), flush=True)
