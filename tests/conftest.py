import os
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    warnings.filterwarnings("ignore", message=".*added jitter.*")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "volt_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def eval_golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "eval_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def c1_golden():
    """BASELINE config c1 (example.ipynb path, n = 256): tests/golden/make_golden_c1.py."""
    return torch.load(os.path.join(ROOT, "tests", "golden", "c1_golden.pt"), weights_only=False)
