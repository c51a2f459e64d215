import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')
    config.addinivalue_line('markers', 'slow: long-running')


@pytest.fixture(scope='session')
def gpu_device():
    """Fail loudly (not skip) if a gpu-marked test runs without the CUDA path."""
    from pysph_b200 import _lib
    _lib.load()
    return 0
