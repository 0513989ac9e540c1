import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope='session')
def lcb():
    """The built library; building is the job of __graft_entry__.build()."""
    from lightcurver_b200 import _lib
    return _lib


@pytest.fixture(scope='session')
def cuda_device(lcb):
    if lcb.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return 0
