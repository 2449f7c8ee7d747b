"""Stand-in for POT (`ot`), absent offline and unpinned in the reference (README.md:29, era 0.9.x).
TEST INFRASTRUCTURE ONLY: lets compute_otmi.py / gromov_wasserstein.py import and run unmodified
when golden vectors are generated.  Parity with real POT is UNPINNED (no POT tests/vectors in the
reference tree); the restated algorithms are documented in gromov.py.
"""
import numpy as np
from . import gromov  # noqa: F401


def unif(n):
    return np.ones((n,)) / n
