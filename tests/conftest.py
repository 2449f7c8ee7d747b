import glob
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


def golden(pattern):
    """Sorted list of (name, path) for fixtures tests/golden/<pattern>.npz."""
    return [(os.path.basename(p)[:-4], p) for p in sorted(glob.glob(os.path.join(GOLDEN, pattern + ".npz")))]


def load(path):
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def assert_close(a, b, rtol=1e-5, atol=1e-9, what=""):
    """Float parity bar of BASELINE.json: <= 1e-5 relative (+ a small absolute floor for values that are
    differences of nearly equal numbers, e.g. the variance channels). NaN must match NaN."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{what}: NaN pattern differs ({nan_a.sum()} vs {nan_b.sum()})"
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    ok |= nan_a
    ok |= (a == b)  # equal infinities
    if not ok.all():
        i = np.unravel_index(np.argmax(np.where(ok, 0, np.abs(a - b))), a.shape)
        raise AssertionError(f"{what}: {(~ok).sum()} mismatches, worst at {i}: got {a[i]!r} want {b[i]!r}")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
