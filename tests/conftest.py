import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure csrc/libpvder_b200.so exists and is not older than its sources (nvcc cross-compiles without a GPU);
    a no-op when it is up to date.  Without nvcc the tests use whatever library is there and fail loudly if none is."""
    import shutil

    from gym_pvder_b200 import _cabi

    if shutil.which("nvcc") and not os.environ.get("PVDER_B200_LIB"):
        _cabi.build_library()
    yield


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gym_pvder_b200  # noqa: F401  (fails loudly if the extension is missing)
    from gym_pvder_b200 import _cabi

    _cabi.load()
    return torch.device("cuda:0")
