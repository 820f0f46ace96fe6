import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(params=["chain", "one_cta"])
def reduction_path(request, monkeypatch):
    """Small matrices (n <= 256) have two reductions: the K1-K4 multi-kernel chain and the one-CTA kernel K5
    (small.cu, the default at that size).  ZQ_SMALL_N is read at every solve, so a test asks for either."""
    monkeypatch.setenv("ZQ_SMALL_N", "0" if request.param == "chain" else "256")
    return request.param
