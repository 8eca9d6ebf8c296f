import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu_backend():
    """Initialised aggregation backend on cuda:0; fails (not skips) when the CUDA library is absent."""
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    from pygim_b200.backend_pim import pim_ops
    pim_ops.dpu_init_ranks(1)
    yield pim_ops
    pim_ops.dpu_release()
