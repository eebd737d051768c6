import os
import sys

import pytest
import torch

# ROIAlign forward has a bit-exact parity mode (un-fused tap arithmetic) and a default FMA mode; the GPU
# suite runs in parity mode and checks the default mode separately within the 1e-5 tolerance.
os.environ.setdefault("COIN_ROI_EXACT", "1")

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


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)
