import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library with a device selected; GPU tests fail loudly when it is missing."""
    from magudi_b200 import _lib
    return _lib.init(0)
