"""Mirror of representations/representation_search/compute_otmi.py (reference :6-211): the GWD used for ranking.

With the reference's `max_iter=0` and a loss that ignores its arguments, POT's sampled_gromov_wasserstein(...,
log=True)["gw_dist_estimated"] is the deterministic mean |pad(Ks) - pad(Kt)|; that scalar is what
`evrep_gwd_kernel_l1` computes on the GPU without materialising any n x n matrix.  The transport plan the reference
also returns is the product coupling p q^T (no Sinkhorn iteration runs)."""
import numpy as np
import torch

from ... import batched as eb


def compute_kernel(Cx, Cy, h):
    """Gaussian kernels of two distance matrices (reference :6-32).  Small helper kept for API compatibility: the GWD
    path below never forms these matrices; here they are evaluated with torch on the GPU."""
    dev = "cuda"
    Cx_t, Cy_t = torch.as_tensor(Cx, device=dev, dtype=torch.float64), torch.as_tensor(Cy, device=dev, dtype=torch.float64)
    std1 = torch.sqrt((Cx_t ** 2).mean() / 2)
    std2 = torch.sqrt((Cy_t ** 2).mean() / 2)
    Kx = torch.exp(-((Cx_t / (h * std1)) ** 2) / 2)
    Ky = torch.exp(-((Cy_t / (h * std2)) ** 2) / 2)
    return Kx.cpu().numpy(), Ky.cpu().numpy()


def pad_arrays_to_same_shape(arr1, arr2):
    """Zero-pad bottom/right to the common shape (reference :35-47)."""
    shape = tuple(max(a, b) for a, b in zip(arr1.shape, arr2.shape))
    out = []
    for a in (arr1, arr2):
        o = np.zeros(shape, a.dtype)
        o[tuple(slice(0, s) for s in a.shape)] = a
        out.append(o)
    return out[0], out[1]


class OTMI:
    def __init__(self, Xs, Xt, h, reg=0.05):
        self.Xs = np.asarray(Xs, np.float64)
        self.Xt = np.asarray(Xt, np.float64)
        self.h = h
        self.reg = reg

    def solve(self):
        n, m = len(self.Xs), len(self.Xt)
        cost = float(eb.gwd_kernel_l1([self.Xs], [self.Xt], self.h)[0].item())
        T = np.outer(np.ones(n) / n, np.ones(m) / m)
        return T, cost


def _otmi_pairs(events, rep, height, width, rep_size):
    """The data preparation of otmi() (reference :96-203) on the GPU (evrep_otmi_prepare): quadrant split, densest quadrant
    dropped, coordinates normalised, representation cropped with two positional channels, empty pixels removed."""
    ev = np.asarray(events) if not torch.is_tensor(events) else events
    if len(ev) == 0:
        raise ValueError("min() arg is an empty sequence")
    pairs, info = eb.otmi_prepare(ev, rep, height, width, rep_size)
    kept = [q for q in range(4) if q != info["dropped"]]
    if any(info["events_per_quadrant"][q] == 0 for q in kept):
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # t[0] of an empty first quadrant (reference :167)
    return pairs


def otmi(events, rep, height, width, rep_size):
    """Mean GWD-A cost over the three kept quadrants (reference :96-211)."""
    pairs = _otmi_pairs(events, rep, height, width, rep_size)
    costs = eb.gwd_kernel_l1([a for a, _ in pairs], [b for _, b in pairs], 0.7)
    return float(costs.mean().item())
