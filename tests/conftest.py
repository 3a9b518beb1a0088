import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native():
    """The product library; building it is part of the CPU suite (nvcc cross-compiles)."""
    import megakv_b200
    from megakv_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", ROOT])
    return megakv_b200.lib()


@pytest.fixture(scope="session")
def gpu(native):
    import megakv_b200
    if native.gpuhash_device_count() < 1:
        pytest.fail("test marked gpu but no CUDA device is visible (no CPU fallback exists)")
    return megakv_b200


@pytest.fixture()
def rng():
    return np.random.default_rng(12345)
