"""Builds pygim_b200/libbackend_pim.so for sm_100a with nvcc (in-tree, no JIT cache).

    python -m pygim_b200.build [--force]

The library keeps the file name of the reference's per-configuration plugin
(backend_pim/<variant>/build/libbackend_pim.so, spmm_test.py:85) so `--lib_path` keeps meaning
"the aggregation backend to dlopen".  One translation unit per element type is compiled in
parallel (kernels_inst.cu with -DPYGIM_T/-DPYGIM_SFX) plus the C-ABI / plan unit.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libbackend_pim.so")

DTYPES = [("int8_t", "i8"), ("int16_t", "i16"), ("int32_t", "i32"), ("int64_t", "i64"), ("float", "f32"),
          ("double", "f64")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "177"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.access(cand, os.X_OK):
            return cand
    raise RuntimeError("nvcc not found: pygim_b200 needs the CUDA toolkit to build libbackend_pim.so")


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + \
        [os.path.join(HERE, "..", "include", "pygim_b200.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force: bool = False, verbose: bool = False, defines=None, out: str = None) -> str:
    """`defines` (e.g. ["PYGIM_CSR_UNROLL=4"]) + `out` build a tuning variant next to the default library."""
    global LIB, OBJ
    lib_default, obj_default = LIB, OBJ
    if defines or out:
        assert out, "a variant build needs an output name"
        LIB = os.path.join(HERE, out)
        OBJ = os.path.join(CSRC, "build_" + os.path.splitext(out)[0])
        force = True
    try:
        return _build(force, verbose, list(defines or []))
    finally:
        LIB, OBJ = lib_default, obj_default


def _build(force: bool, verbose: bool, defines) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + ["-D" + d for d in defines]
    jobs = []
    for ctype, sfx in DTYPES:
        obj = os.path.join(OBJ, "kernels_%s.o" % sfx)
        jobs.append((obj, [nvcc] + flags + ["-DPYGIM_T=" + ctype, "-DPYGIM_SFX=" + sfx, "-c",
                                            os.path.join(CSRC, "kernels_inst.cu"), "-o", obj]))
    for unit in ("backend_pim", "quantize"):
        obj = os.path.join(OBJ, unit + ".o")
        jobs.append((obj, [nvcc] + flags + ["-c", os.path.join(CSRC, unit + ".cu"), "-o", obj]))
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for out in ex.map(lambda j: _run(j[1]), jobs):
            if verbose and out:
                print(out)
    tmp = LIB + ".tmp"
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", tmp] +
         [j[0] for j in jobs])
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("-D", dest="defines", action="append", default=[])
    ap.add_argument("-o", dest="out", default=None, help="file name of a tuning variant, e.g. libbackend_pim_u4.so")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, defines=a.defines, out=a.out))
    sys.exit(0)
