"""The product package must not import, link or execute anything under oracle/ (nor torch sparse
matmuls as a fallback): the CUDA library is the only implementation of the hot path."""
import ast
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pygim_b200")


def _py_files():
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(d, f)


def test_no_oracle_imports():
    for path in _py_files():
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                assert not re.match(r"^(oracle)(\.|$)", n), "%s imports %s" % (path, n)
        text = open(path).read()
        assert "liboracle" not in text and "oracle/_ref" not in text, path


def test_no_cpu_or_torch_sparse_fallback():
    banned = ("torch.sparse.mm", "torch.sparse_coo_tensor", "torch_sparse.matmul", "scipy.sparse", "index_add_(0, row")
    for path in _py_files():
        text = open(path).read()
        for b in banned:
            assert b not in text, "%s uses %s" % (path, b)


def test_native_sources_do_not_reference_the_oracle():
    for d, _, files in os.walk(os.path.join(PKG, "csrc")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(d, f)).read().lower(), f
