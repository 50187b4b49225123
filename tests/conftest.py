import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and return the path of libsomax_b200.so; nvcc needs no GPU."""
    from somax_b200 import _lib
    return _lib.build_library()
