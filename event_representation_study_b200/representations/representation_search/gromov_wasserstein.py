"""Mirror of representations/representation_search/gromov_wasserstein.py: `compute_repr` (reference :72-82), the 5-bin
bilinear voxel grid used by its __main__ experiment, `compute_kernel`, and `OTMI` (reference :39-69), whose `solve()` is
POT's conditional-gradient Gromov-Wasserstein with the KL loss: here `evrep_gw_kl` (tcgen05 contraction on the GPU,
auction LMO on the GPU for n == m, exact transportation LMO on the host for n != m - the reference's own call shape,
N events against the non-empty pixels; see include/evrep.h)."""
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


class OTMI:
    """OTMI(Xs, Xt, h, reg=0.05).solve() -> (T, gw_dist) like the reference (gromov_wasserstein.py:39-69); `reg` is
    stored and unused there too.  T is the (n, m) float64 plan on the host."""

    def __init__(self, Xs, Xt, h, reg=0.05):
        self.Xs = np.asarray(Xs, np.float64)
        self.Xt = np.asarray(Xt, np.float64)
        self.h = h
        self.reg = reg
        self.P = None

    def solve(self):
        dist, _, plan = eb.gw_kl(self.Xs, self.Xt, self.h, return_plan=True)
        self.P = plan.double().cpu().numpy()
        return self.P, dist
