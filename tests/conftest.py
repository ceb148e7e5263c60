import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib_built():
    """The C-ABI library must exist in-tree (built by __graft_entry__.build()); build it if a test box lacks it."""
    import __graft_entry__ as g
    g.build()
    from diffews_b200 import _lib
    return _lib
