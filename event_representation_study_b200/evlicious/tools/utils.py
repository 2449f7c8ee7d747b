"""Mirror of ev-licious/src/evlicious/tools/utils.py::events_to_voxel_grid[_cuda] (reference :7-85).

Both run the same deterministic GPU kernel.  The reference's numpy version passes the integer bin to its bilinear weight
(`_bil_w(t_norm_int, tlim)`, :74), so it is a floor-bin polarity histogram; that quirk is kept.  The reference's torch
version scatters with `put_` WITHOUT accumulation (:38, last writer wins, nondeterministic) - a bug we do not reproduce:
`events_to_voxel_grid_cuda` returns the accumulate semantics of the numpy version, as a torch tensor on `device`.
Only integer pixel coordinates (divider == 1) are supported on the GPU."""
import numpy as np
import torch

from ... import batched as eb
from ..._single import one_window


def _grid(events, num_bins, normalize, t0_us, t1_us):
    if events.divider > 1:
        raise NotImplementedError("sub-pixel coordinates (divider > 1) are not supported by the GPU voxel grid")
    H, W = events.height, events.width
    ev = one_window(events.x, events.y, events.t, events.p, H, W)
    if len(events) < 2:
        return torch.zeros((num_bins, H, W), dtype=torch.float32, device=ev.device)
    t0 = int(t0_us) if t0_us is not None else int(events.t[0])
    t1 = int(t1_us) if t1_us is not None else int(events.t[-1])
    return eb.voxel_grid(ev, H, W, num_bins, "evlicious", normalize=normalize, t0_us=t0, t1_us=t1)[0]


def events_to_voxel_grid(events, num_bins, normalize=True, t0_us=None, t1_us=None):
    """-> numpy float32 (num_bins, height, width)."""
    return _grid(events, num_bins, normalize, t0_us, t1_us).cpu().numpy()


def events_to_voxel_grid_cuda(events, num_bins, normalize=True, t0_us=None, t1_us=None, device="cuda:0"):
    """-> torch float32 (num_bins, height, width) on `device`."""
    with torch.cuda.device(torch.device(device)):
        return _grid(events, num_bins, normalize, t0_us, t1_us).to(device)
