import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    from tests.util import load_golden
    return load_golden()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (absent -> skip)."""
    from oracle import pyoracle
    try:
        pyoracle.build()
    except Exception:
        pass
    if not pyoracle.have_reference():
        pytest.skip("oracle/_ref/libvfgs_ref.so not built (reference sources not mounted)")
    return pyoracle.Reference()


@pytest.fixture(scope="session")
def hw_lib():
    """The product library; building/loading it needs nvcc but no GPU."""
    from versatilefilmgrain_b200 import VfgsHw
    return VfgsHw()
