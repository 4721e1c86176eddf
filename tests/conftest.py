import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree CUDA library; builds it when nvcc is available."""
    from gogp_b200 import _lib, build
    build.build()
    assert os.path.exists(_lib.LIB_PATH)
    return _lib.LIB_PATH
