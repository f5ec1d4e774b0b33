"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here on CPU (oracle vs golden vectors / netlib restatement, host logic,
C-ABI symbol export, gloo world_size-2); `-m gpu` are the parity tests proper and call the
CUDA path through the C ABI (include/lbk.h) on a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def blas():
    """The product: lambda_blas_b200 over the C ABI.  Fails loudly without the CUDA library."""
    import lambda_blas_b200 as lb
    return lb


@pytest.fixture(scope="session")
def ctx(blas):
    c = blas.Context(device=0)
    yield c
    c.close()
