import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    try:
        from bore_b200 import _lib
        return int(_lib.load().bore_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without CUDA: gpu-marked tests are skipped, not failed (the
    product has no CPU fallback, so they could only raise BoreNativeError)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu") is not None]
    if not gpu_items or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (bore_b200 has no CPU fallback)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def native_lib():
    """libbore_b200.so, built on demand (nvcc cross-compiles without a GPU)."""
    from bore_b200 import _lib
    return _lib.load()
