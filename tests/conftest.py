import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(autouse=True)
def _fixed_seed():
    """Same inputs in every process: torch seeds its default generator from the system entropy when nothing else is
    said, and FreeFermion's Metropolis chains (seed = torch.initial_seed() at the first sample()) and every unseeded
    torch.randn follow it.  A test that wants another seed sets it itself."""
    import torch
    torch.manual_seed(int(os.environ.get("FF_TEST_SEED", "20240607")))
    yield
