import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libnsx.so / liboracle.so when they are missing (they travel pre-built to the GPU box)."""
    import nsxlib
    if not (os.path.exists(nsxlib.LIBNSX) and os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so"))):
        import __graft_entry__
        __graft_entry__.build()
