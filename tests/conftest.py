import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def engine_lib():
    """Path of the built CUDA engine; built here if absent (nvcc cross-compiles without a GPU)."""
    from mpassit_b200 import build, lib

    if not os.path.exists(lib.LIB_PATH):
        build.build()
    return lib.LIB_PATH


@pytest.fixture(scope="session")
def have_gpu():
    import torch

    return torch.cuda.is_available()
