import os
import sys

# In-process rank groups (tests/test_gpu_local_ranks.py) must never block the host while one rank's stream is parked on a
# wait for another rank: load every kernel when its module is loaded instead of at first launch (a lazy load can
# synchronise the device).  Must be set before CUDA is initialised, i.e. before anything below touches libcuda.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count() -> int:
    """number of CUDA devices through the driver API (no torch import, no context created)"""
    import ctypes

    try:
        cu = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return 0
    n = ctypes.c_int(0)
    if cu.cuInit(0) != 0 or cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
        return 0
    return n.value


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device"""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_build():
    """Build the native pieces once (no-ops when the in-tree .so files are current)."""
    from solidstateamfoam_b200 import _build

    _build.build_meshgen()
    _build.build_oracle()
    _build.build_cuda()
    yield
