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


def load_golden(family):
    """tests/golden/<family>.npz -> {case: {key: array}} (written by oracle/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, family + ".npz"), allow_pickle=False)
    out = {}
    for key in z.files:
        case, name = key.split("/", 1)
        out.setdefault(case, {})[name] = z[key]
    return out


def normwise(a, b):
    """max|a-b| / max|b| -- the gradient metric of SURVEY.md section 0.5 / 8(c)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = np.abs(b).max()
    return float(np.abs(a - b).max() / denom) if denom > 0 else float(np.abs(a - b).max())


@pytest.fixture(scope="session")
def golden_uncl():
    return load_golden("uncl")


@pytest.fixture(scope="session")
def golden_fecl():
    return load_golden("fecl")
