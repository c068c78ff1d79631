import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


class Cases:
    """npz written by tools/make_golden.py: keys '<case>/<field>' plus 'names'."""

    def __init__(self, fname):
        self.z = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
        self.names = [str(s) for s in self.z["names"]]

    def get(self, name):
        pre = name + "/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}


@pytest.fixture(scope="session")
def fit_cases():
    return Cases("fit_cases.npz")


@pytest.fixture(scope="session")
def merit_cases():
    return Cases("merit_cases.npz")


@pytest.fixture(scope="session")
def candidate_cases():
    return Cases("candidate_cases.npz")


def relmax(a, b):
    """max |a - b| / max |b|  -- the parity metric of SURVEY section 8(c) / BASELINE north_star (1e-10)."""
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


class Problem:
    """Minimal stand-in for pySOT's OptimizationProblem attributes the candidate generators read."""

    def __init__(self, lb, ub, int_var=()):
        self.lb, self.ub = np.asarray(lb, float), np.asarray(ub, float)
        self.dim = len(self.lb)
        self.int_var = np.array(list(int_var), dtype=int)
        self.cont_var = np.array([i for i in range(self.dim) if i not in set(int_var)], dtype=int)
