import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def orc():
    """CPU oracle (oracle/liboracle.so) -- the checker."""
    import __graft_entry__ as ge
    return ge.load_oracle()


@pytest.fixture(scope="session")
def spfy():
    """The product package (loads libsparsifyme_b200.so; no fallback)."""
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def cuda(spfy):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a -m gpu test was collected without a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
