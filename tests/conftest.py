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


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def sample_inputs():
    return load_golden("sample_inputs")


def rel_err(a, b):
    """max |a-b| / max |b| — the tolerance norm used throughout (per-array, not per-element:
    coverage values span 40 orders of magnitude and only the large ones carry the objective)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0


def row_rel_err(a, b, floor=1e-4):
    """Per-row relative error for per-pose gradient arrays (W,3) / (W,4): every pose against ITS OWN magnitude,
    max_i |a_i - b_i|_inf / |b_i|_inf over the rows whose magnitude is at least `floor` of the largest row (rows below
    that are dominated by cancellation noise of the reference itself).  `rel_err` over the whole array would let a pose
    whose gradient is 1e-3 of the largest be 10 % wrong and still pass 1e-4."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a, b = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    if a.size == 0:
        return 0.0
    nb = np.abs(b).max(axis=1)
    keep = nb >= floor * max(nb.max(), 1e-300)
    if not keep.any():
        return 0.0
    return float((np.abs(a - b).max(axis=1)[keep] / nb[keep]).max())
