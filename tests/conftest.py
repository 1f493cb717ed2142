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
    """The CUDA library with a device selected.  On a box without any CUDA device the GPU tests are skipped
    (the library itself still fails loudly: tests/test_abi.py::test_no_cpu_fallback_without_device);
    with a device present every failure to initialise is an error."""
    from magudi_b200 import _lib
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = True
    if not have:
        pytest.skip("no CUDA device on this box")
    return _lib.init(0)
