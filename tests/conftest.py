import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load_oracle()


@pytest.fixture(scope="session")
def ref():
    import oracle_lib
    r = oracle_lib.load_ref()
    if r is None:
        pytest.skip("oracle/_ref/libqref.so not available (reference sources absent and no prebuilt copy)")
    return r


@pytest.fixture(scope="session")
def qb():
    """The product (CUDA library via ctypes).  GPU tests only."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import qblas_b200
    qblas_b200.init()
    return qblas_b200
