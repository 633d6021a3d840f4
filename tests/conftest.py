import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure; never imported by the package)."""
    from oracle import oracle

    oracle.build()
    return oracle


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


CASE_NAMES = ["case_overlap_wall", "case_overlap_free", "case_touch_wall", "case_touch_free", "case_near_wall"]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


# tolerances of BASELINE.json's north_star: <= 1e-12 relative in double, <= 1e-5 in float
TOL = {"double": 1e-12, "single": 1e-5}
