import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", message="Sparse invariant checks")
warnings.filterwarnings("ignore", message="Named tensors")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """libdsw.so, built on demand (nvcc cross-compiles without a GPU)."""
    from deepsphere_weather_b200 import _lib, build

    build.build()
    return _lib.load()
