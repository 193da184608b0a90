import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build the in-tree native pieces once per session (nvcc cross-compiles without a GPU)."""
    from emsar_b200 import build
    build.build_all()
    from oracle import oracle
    oracle.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    from emsar_b200.api import Context
    c = Context(0)
    yield c
    c.close()
