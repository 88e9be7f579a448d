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
def gold_eval():
    return np.load(os.path.join(GOLDEN, "ks_eval.npz"))


@pytest.fixture(scope="session")
def gold_big():
    return np.load(os.path.join(GOLDEN, "ks_eval_256.npz"))


@pytest.fixture(scope="session")
def gold_hunt():
    return np.load(os.path.join(GOLDEN, "ks_hunt.npz"))


def golden_cases(gold, prefix):
    """Case keys 'prefix/<name>' present in a golden npz."""
    names = sorted({k.split("/")[1] for k in gold.files if k.startswith(prefix + "/")})
    return [f"{prefix}/{n}" for n in names]


def relerr(a, b, scale=None):
    """max|a-b| / scale; scale defaults to max|b| (pass the size of the cancelling terms for near-zero residuals)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if scale is None:
        scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b))) / scale if b.size else 0.0
