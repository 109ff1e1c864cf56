import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def relerr(a, b):
    """Norm-wise relative error max|a-b|/max|b| (SURVEY.md appendix B tolerance guidance)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


_worlds = {}


@pytest.fixture(scope="session")
def world():
    from spectra_without_windows_b200 import synthetic as syn

    def get(name):
        if name not in _worlds:
            _worlds[name] = syn.make_world(name)
        return _worlds[name]
    return get
