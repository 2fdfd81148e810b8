import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never pass silently without a GPU: when selected with -m gpu on a box
    without CUDA they FAIL (the product has no CPU fallback)."""
    return


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library initialised on cuda:0 (GPU tests only)."""
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    from dense_linear_app_b200 import _lib
    _lib.call("chol_init", 0)
    return _lib
