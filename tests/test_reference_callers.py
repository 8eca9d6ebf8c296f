"""Pins the Python-side preprocessing and the quantisation helpers against the REFERENCE'S OWN modules, imported
from /root/reference in this container (skipped on the GPU box, where the reference tree does not exist).
`torch_sparse` is not installable, so a stand-in module exposing our SparseTensor is injected for the import; the
reference classes then run unchanged on top of it up to (not including) the `torch.ops.pim_ops` calls."""
import importlib.util
import os
import sys
import types

import pytest
import torch

from helpers import random_adj
from pygim_b200.backend_pim import grande as our_grande
from pygim_b200.backend_pim import spmm as our_spmm
from pygim_b200.backend_pim import spmv as our_spmv
from pygim_b200.models import quantize as our_quantize
from pygim_b200.sparse_tensor import SparseTensor

REF = os.environ.get("PYGIM_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "backend_pim")), reason="reference tree absent")


@pytest.fixture(scope="module")
def ref():
    shim = types.ModuleType("torch_sparse")
    shim.SparseTensor = SparseTensor
    shim.matmul = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("not needed"))
    class _Stub(types.ModuleType):          # import-only stand-ins: spmv.py imports PyG / ogb names it never uses
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return type(item, (), {})

    names = ["torch_sparse", "torch_geometric", "torch_geometric.transforms", "torch_geometric.utils",
             "torch_geometric.datasets", "ogb", "ogb.nodeproppred", "ogb.linkproppred", "ogb.graphproppred"]
    saved_all = {n: sys.modules.get(n) for n in names}
    for n in names[1:]:
        sys.modules[n] = _Stub(n)
    saved = saved_all["torch_sparse"]
    sys.modules["torch_sparse"] = shim
    mods = {}
    try:
        for name, rel in (("ref_spmm", "backend_pim/spmm.py"), ("ref_grande", "backend_pim/grande.py"),
                          ("ref_spmv", "backend_pim/spmv.py"), ("ref_quantize", "models/quantize.py")):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods[name] = mod
        yield types.SimpleNamespace(**mods)
    finally:
        for n, old in saved_all.items():
            if old is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = old


def test_type_tables_and_dense_split(ref):
    assert our_spmm.TORCH_TYPES == ref.ref_spmm.TORCH_TYPES == ref.ref_spmv.TORCH_TYPES
    assert our_grande.TYPES_MUL == ref.ref_grande.TYPES_MUL
    B = torch.arange(7 * 10, dtype=torch.float32).reshape(7, 10)
    for nparts in (1, 2, 3, 4, 5):
        a, b = our_spmm.dense_split(B, nparts), ref.ref_spmm.dense_split(B, nparts)
        assert len(a) == len(b) and all(torch.equal(x, y) and x.is_contiguous() for x, y in zip(a, b))
        a, b = our_spmv.dense_split(B, nparts), ref.ref_spmv.dense_split(B, nparts)
        assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))
    for dtype, ncols in ((torch.float32, [4, 3, 3]), (torch.int8, [3, 3, 2, 2]), (torch.int64, [5, 5]), (torch.int16, [10])):
        Bd = torch.arange(7 * 10).reshape(7, 10).to(dtype)
        a, b = our_grande.dense_split(Bd, ncols), ref.ref_grande.dense_split(Bd, ncols)
        assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("dtype", [torch.int8, torch.int32, torch.float32, torch.float64])
@pytest.mark.parametrize("nparts", [1, 2, 3, 5])
def test_col_split_and_part_arrays_match_the_reference(ref, dtype, nparts):
    for with_values in (False, True):
        adj = random_adj(23, 31, 0.25, seed=nparts, value_dtype=torch.float32 if with_values else None,
                         real_valued=True, empty_rows=(0, 22))
        ours = our_spmm.SparseTensorCOO(adj, dtype=dtype, format="CSR")
        theirs = ref.ref_spmm.SparseTensorCOO(adj, dtype=dtype, format="CSR")
        assert [p.sizes() for p in ours.col_split(nparts)] == [p.sizes() for p in theirs.col_split(nparts)]
        ours.build_csr()
        theirs.build_csr()
        ours.build_coo()
        theirs.build_coo()
        for a, b in zip(ours.csr, theirs.csr):
            assert torch.equal(a.crow_indices(), b.crow_indices()) and torch.equal(a.col_indices(), b.col_indices())
            assert torch.equal(a.values(), b.values()) and a.values().dtype == dtype
            assert tuple(a.size()) == tuple(b.size())
        for a, b in zip(ours.coo, theirs.coo):
            assert torch.equal(a.indices(), b.indices()) and torch.equal(a.values(), b.values())
            assert tuple(a.size()) == tuple(b.size())


@pytest.mark.parametrize("dtype", [torch.int8, torch.int16, torch.int32, torch.int64])
def test_spmv_padding_matches_the_reference(ref, dtype):
    adj = random_adj(13, 13, 0.3, seed=2)
    ours = our_spmv.SparseTensorCOO(adj, dtype=dtype, groups=4)
    theirs = ref.ref_spmv.SparseTensorCOO(adj, dtype=dtype, groups=4)
    ours.build_coo()
    theirs.build_coo()
    a, b = ours.coo[0], theirs.coo[0]
    assert tuple(a.size()) == tuple(b.size())
    assert torch.equal(a.indices(), b.indices()) and torch.equal(a.values(), b.values())


@pytest.mark.parametrize("dtype", [torch.int8, torch.int16, torch.int32, torch.float32, torch.float64, torch.int64])
def test_quantisation_matches_the_reference(ref, dtype):
    g = torch.Generator().manual_seed(3)
    for shape in ((50, 16), (7, 3)):
        v = torch.randn(shape, generator=g) * 3.7
        s1, q1 = our_quantize.symmetric_quantize(v, dtype)
        s2, q2 = ref.ref_quantize.symmetric_quantize(v, dtype)
        assert torch.equal(s1, s2) and q1.dtype == q2.dtype and torch.equal(q1, q2)
        out = torch.randint(-1000, 1000, shape).to(q1.dtype)
        assert torch.equal(our_quantize.symmetric_dequantize(out, 1.0, s1),
                           ref.ref_quantize.symmetric_dequantize(out, 1.0, s2))


def test_group_width_lists_match_the_reference(ref):
    """h_size / rank_h_size lists handed to the *_to_device_group ops (spmm.py:60-72, grande.py:63-72)."""
    from pygim_b200.backend_pim._common import split_widths
    for hidden in (1, 7, 32, 33, 100, 256):
        for parts in (1, 2, 3, 4, 7, 32):
            if parts > hidden:
                continue
            max_h = (hidden + parts - 1) // parts            # the reference's arithmetic, restated
            want = [max_h] * parts
            if parts * max_h != hidden:
                want[parts - 1] = hidden - (parts - 1) * max_h
            assert split_widths(hidden, parts) == want
