import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback for the product path")
    from aesrc2020_b200 import _shim
    _shim.lib()
    return torch.device("cuda:0")


def pytest_collection_modifyitems(config, items):
    """A protocol bug in a kernel can hang a GPU test: every GPU test gets a wall-clock limit (pytest-timeout, thread
    method: the process is killed, which also ends a hung CUDA call)."""
    for it in items:
        if "gpu" in it.keywords and not any(m.name == "timeout" for m in it.iter_markers()):
            it.add_marker(pytest.mark.timeout(240, method="thread"))
