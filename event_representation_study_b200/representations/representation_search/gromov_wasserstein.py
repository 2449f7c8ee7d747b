"""Mirror of the in-scope parts of representations/representation_search/gromov_wasserstein.py:
`compute_repr` (reference :72-82), the 5-bin bilinear voxel grid used by its __main__ experiment, and `compute_kernel`.
The conditional-gradient Gromov-Wasserstein solve of that file (OTMI.solve, reference :62-69, POT's
ot.gromov.gromov_wasserstein with kl_loss) is not built yet ("what comes next")."""
import numpy as np

from ... import batched as eb
from ..._single import one_window
from .compute_otmi import compute_kernel  # noqa: F401  (same function in both reference files)

_T_SCALE = 2**30 - 2


def compute_repr(x, y, t, p, width, height, bins=5):
    """t in [0, 1] (the caller normalises); -> float64 (height, width, bins)."""
    t = np.asarray(t, np.float64)
    if len(t) < 2 or not (t.min() == 0.0 and t.max() == 1.0):
        raise ValueError("compute_repr on the GPU expects t normalised as (t - t[0]) / (t[-1] - t[0]) with at least two events")
    # the kernel re-normalises integer timestamps by (first, last); quantise [0, 1] on a 2^30 grid (error < 1e-9 in t)
    ti = np.rint(t * _T_SCALE).astype(np.int64)
    if ti[0] != 0 or ti[-1] != _T_SCALE:
        raise ValueError("events must be time sorted (t[0] == 0, t[-1] == 1)")
    ev = one_window(x, y, ti, np.asarray(p).astype(np.int64), height, width)
    return eb.voxel_grid(ev, height, width, bins, "gwd")[0].double().cpu().numpy()
