import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))       # xr_stub

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fixtures():
    """The reference's own golden fixtures (re-packed by tests/golden/make_golden.py)."""
    return dict(np.load(os.path.join(GOLDEN, "fixtures.npz")))


@pytest.fixture(scope="session")
def live():
    """Outputs of the live reference on small seeded inputs."""
    return dict(np.load(os.path.join(GOLDEN, "live_cases.npz")))


@pytest.fixture(scope="session")
def live_next():
    """Live-reference outputs of predict / reconstructed_fields / hom-het patterns / bootstrapping."""
    return dict(np.load(os.path.join(GOLDEN, "live_next.npz")))


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
