import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Fixtures produced by oracle/gen_golden.py from the unmodified reference (never read /root/reference here)."""
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def parity_set(g):
    """options whose REFERENCE result is finite with solver.error < 1e-2."""
    return np.isfinite(g["price"]) & (g["err"] < 1e-2)


# bars from BASELINE.json's north_star
PRICE_ABS_TOL = 1e-9
BOUNDARY_REL_TOL = 1e-10
