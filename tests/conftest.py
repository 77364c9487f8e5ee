import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def built_library():
    """Build (no-op when up to date) and load libneurons_mm.so.  nvcc cross-compiles without a GPU."""
    from neurons_b200 import build, lib
    if not os.path.isfile(lib.LIB_PATH):
        build.build()
    return lib.load()
