import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun)")


@pytest.fixture(scope="session")
def stp():
    """The checked ctypes call layer over libstp.so (built in-tree; fails loudly if absent)."""
    from segmentation_training_pipeline_b200 import lib
    return lib.Lib()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
