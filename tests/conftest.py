import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def po():
    """The CPU oracle (test infrastructure)."""
    import pyoracle
    pyoracle.build()
    pyoracle.set_threads(min(16, pyoracle.hw_threads()))
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    """A device context; GPU tests only.  Fails loudly if the CUDA extension is missing."""
    from icp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
