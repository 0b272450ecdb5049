import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of libcpc_b200.so, building it with nvcc when it is missing (cross-compiles without a GPU)."""
    from cpc_audio_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from cpc_audio_b200 import build
        build.build()
    return _lib.LIB_PATH
