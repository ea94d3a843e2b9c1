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
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def uf():
    import unsflow_b200
    return unsflow_b200


@pytest.fixture(scope="session")
def gpu_lib(uf):
    """the CUDA library, asserting a device is really there (no silent fallback)"""
    L = uf.lib()
    assert uf.device_count() >= 1, "no CUDA device: GPU tests must not run on a fallback"
    return L
