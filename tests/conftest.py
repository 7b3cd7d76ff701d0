import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib.load()


@pytest.fixture(scope="session")
def gpu_ops():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from maestro_b200 import lib

    return lib.init(0)
