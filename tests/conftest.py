import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neuralquantum.jl_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def nq():
    """The product: ctypes bindings over libnqcuda (fails loudly if the library is missing)."""
    import nqcuda
    return nqcuda


@pytest.fixture(scope="session")
def ctx(nq):
    """One libnqcuda context on cuda:0 sharing torch's current stream."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.init()
    c = nq.Context(0, torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()
