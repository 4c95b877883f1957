import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def built_library():
    """The product library must exist in-tree; build it if this checkout has not been built yet."""
    from swf_renderer_b200 import capi

    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return capi.load()
