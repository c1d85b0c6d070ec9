import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def ctx():
    """One integrator context on cuda:0 for the whole GPU session (fails loudly without the CUDA library)."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from integrator2_b200 import abi
    c = abi.Context(0)
    yield c
    c.close()
