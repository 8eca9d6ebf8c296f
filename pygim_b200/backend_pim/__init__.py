"""Drop-in replacement of PyGim's `backend_pim` Python package for B200.

    from pygim_b200.backend_pim.spmm import prepare_pim_spmm, pim_spmm
    from pygim_b200.backend_pim.grande import prepare_pim_spmm_grande
    from pygim_b200.backend_pim.spmv import prepare_pim_spmv

Importing the package registers the `pim_ops` operators under torch.ops so that the reference's
unchanged drivers (`torch.ops.pim_ops.dpu_init_ranks(...)`, spmm_test.py:112-118) find them.
"""
from . import pim_ops

pim_ops.register_torch_ops()
