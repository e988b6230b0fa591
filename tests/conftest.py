"""pytest configuration: the ``gpu`` marker and shared helpers."""

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_boundary():
    import numpy as np

    return np.load(ROOT / "tests" / "golden" / "boundary.npz")


@pytest.fixture(scope="session")
def golden_classes():
    import numpy as np

    return np.load(ROOT / "tests" / "golden" / "classes.npz")
