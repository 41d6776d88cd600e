import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    import openlte_b200
    return openlte_b200.load()


@pytest.fixture(scope="session")
def ctx(lib):
    import torch
    assert torch.cuda.is_available(), "gpu-marked test run without a CUDA device"
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    c = lib.new_ctx(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.lib()
    return p


@pytest.fixture(scope="session")
def refo():
    """The reference itself (oracle/_ref), or None where it has not been built."""
    from oracle import ref
    return ref if ref.available() else None
