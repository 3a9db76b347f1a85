import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib_built():
    """Build the shared library once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True
