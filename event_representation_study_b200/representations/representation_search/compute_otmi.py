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
    """The data preparation of otmi() (reference :96-203) with torch on the GPU: quadrant split, densest quadrant
    dropped, coordinates normalised, representation cropped with two positional channels, empty pixels removed."""
    dev = "cuda"
    ev = torch.as_tensor(np.asarray(events) if not torch.is_tensor(events) else events).to(dev)
    X, Y = ev[:, 0], ev[:, 1]
    w2, h2 = width / 2 - 1, height / 2 - 1
    masks = [(X >= 0) & (X <= w2) & (Y >= 0) & (Y <= h2), (X > w2) & (X <= width - 1) & (Y >= 0) & (Y <= h2),
             (X >= 0) & (X <= w2) & (Y > h2) & (Y <= height - 1), (X > w2) & (X <= width - 1) & (Y > h2) & (Y <= height - 1)]
    quads = [ev[m].clone() for m in masks]
    sizes = [q.shape[0] for q in quads]
    ind = sizes.index(max(sizes))
    for q in quads[1:]:
        if q.shape[0] == 0:
            raise RuntimeError("min(): Expected reduction dim to be specified for input.numel() == 0")  # what torch raises in the reference
        q[:, 0] = q[:, 0] - q[:, 0].min()
        q[:, 1] = q[:, 1] - q[:, 1].min()
    r = rep_size
    xys = [([0, r // 2 - 1], [0, r / 2 - 1]), ([r / 2 - 1, r - 1], [0, r / 2 - 1]), ([0, r / 2 - 1], [r / 2 - 1, r - 1]),
           ([r / 2 - 1, r - 1], [r / 2 - 1, r - 1])]
    rep_t = torch.as_tensor(np.asarray(rep), device=dev).double()
    pairs = []
    for i, q in enumerate(quads):
        if i == ind:
            continue
        x = q[:, 0] / ((width - 1) // 2)
        y = q[:, 1] / ((height - 1) // 2)
        t = q[:, 2]
        t = (t - t[0]) / (t[-1] - t[0])
        p = q[:, 3]
        p = (p - p.min()) / (p.max() - p.min())
        mask = (q[:, 0] < (width - 1) // 2) & (q[:, 1] < (height - 1) // 2)
        Xs = torch.stack([x[mask], y[mask], t[mask], p[mask]], dim=-1).double()
        cx, cy = xys[i]
        rp = rep_t[int(cy[0]): int(cy[1]) + 1, int(cx[0]): int(cx[1]) + 1, :]
        a, b = rp.shape[0], rp.shape[1]
        xe = (torch.arange(a, device=dev, dtype=torch.float64) / (a - 1)).reshape(a, 1).expand(a, b)
        ye = (torch.arange(b, device=dev, dtype=torch.float64) / (b - 1)).reshape(1, b).expand(a, b)
        rp = torch.cat((rp, xe[..., None], ye[..., None]), dim=2).reshape(-1, rep_t.shape[2] + 2)
        rp = rp[rp[:, :-2].abs().sum(-1) > 0]
        pairs.append((Xs, rp))
    return pairs


def otmi(events, rep, height, width, rep_size):
    """Mean GWD-A cost over the three kept quadrants (reference :96-211)."""
    pairs = _otmi_pairs(events, rep, height, width, rep_size)
    costs = eb.gwd_kernel_l1([a for a, _ in pairs], [b for _, b in pairs], 0.7)
    return float(costs.mean().item())
