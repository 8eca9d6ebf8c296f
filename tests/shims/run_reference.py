"""Launcher for the reference's drivers, UNCHANGED: python tests/shims/run_reference.py <script.py> [args...]

It only prepares the import system the way the reference's environment has it: the stand-ins of tests/shims for
torch_sparse / torch_geometric / ogb, and `models` as a package rooted at <reference>/models (inference.py does
`from models.models import GCN` while models/models.py does `from pyg_gcn_conv import GCNConv`, so both the package
and its directory must be importable).  With PYGIM_USE_B200=1 the reference's `backend_pim` front-ends are replaced by
pygim_b200's (the drop-in): `--version spmm|grande|spmv` then runs on the GPU through libbackend_pim.so."""
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PYGIM_REFERENCE_ROOT", "/root/reference")


def prepare():
    for p in (os.path.join(REF, "models"), REF, ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    if os.environ.get("PYGIM_USE_B200") == "1":
        import pygim_b200.backend_pim as ours                 # registers torch.ops.pim_ops
        import pygim_b200.backend_pim.grande
        import pygim_b200.backend_pim.spmm
        import pygim_b200.backend_pim.spmv
        sys.modules["backend_pim"] = ours
        for name in ("spmm", "grande", "spmv"):
            sys.modules["backend_pim." + name] = getattr(ours, name)


if __name__ == "__main__":
    prepare()
    script = sys.argv[1]
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")
