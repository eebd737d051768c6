import os
import sys

import pytest
import torch

# The GPU suite runs the DEFAULT kernels (what bench.py times). ROIAlign forward additionally has a bit-exact parity
# mode (option COIN_ROI_EXACT=1: un-fused tap arithmetic in torchvision's order); tests that check it turn it on
# with the `roi_exact` fixture / `_lib.options(COIN_ROI_EXACT=1)`.

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture
def roi_exact():
    from coin_b200 import _lib
    with _lib.options(COIN_ROI_EXACT=1):
        yield


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)
