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


@pytest.fixture(autouse=True)
def _library_defaults(request):
    """mgpu_set_option is process-global state: every GPU test starts and ends with the switches at their defaults"""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import torch

    if not torch.cuda.is_available():
        yield
        return
    from maestro_b200 import lib

    lib.set_option("defaults", 0)
    yield
    lib.set_option("defaults", 0)
