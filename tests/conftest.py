import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_build():
    """Build the native pieces once (no-ops when the in-tree .so files are current)."""
    from solidstateamfoam_b200 import _build

    _build.build_meshgen()
    _build.build_oracle()
    _build.build_cuda()
    yield
