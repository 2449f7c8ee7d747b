"""Mirror of ev-licious/src/evlicious/tools/utils.py::events_to_voxel_grid[_cuda] (reference :7-85).

Both run the same deterministic GPU kernel.  The reference's numpy version passes the integer bin to its bilinear weight
(`_bil_w(t_norm_int, tlim)`, :74), so it is a floor-bin polarity histogram; that quirk is kept.  The reference's torch
version scatters with `put_` WITHOUT accumulation (:38, last writer wins, nondeterministic) - a bug we do not reproduce:
`events_to_voxel_grid_cuda` returns the accumulate semantics of the numpy version, as a torch tensor on `device`.
Sub-pixel events (`Events.divider > 1`: float32 coordinates, 4-tap bilinear scatter, utils.py:93-103) go through
evrep_voxel_subpixel_batched on the raw integer coordinates."""
import numpy as np
import torch

from ... import batched as eb
from ..._single import one_window


def _grid(events, num_bins, normalize, t0_us, t1_us):
    H, W = events.height, events.width
    div = int(events.divider)
    if div > 1:  # raw sub-pixel integers; one_window checks them against the unscaled extent
        ev = one_window(events._x, events._y, events.t, events.p, (H - 1) * div + 1, (W - 1) * div + 1)
    else:
        ev = one_window(events.x, events.y, events.t, events.p, H, W)
    if len(events) < 2:
        return torch.zeros((num_bins, H, W), dtype=torch.float32, device=ev.device)
    t0 = int(t0_us) if t0_us is not None else int(events.t[0])
    t1 = int(t1_us) if t1_us is not None else int(events.t[-1])
    return eb.voxel_grid(ev, H, W, num_bins, "evlicious", normalize=normalize, t0_us=t0, t1_us=t1, divider=max(div, 1))[0]


def events_to_voxel_grid(events, num_bins, normalize=True, t0_us=None, t1_us=None):
    """-> numpy float32 (num_bins, height, width)."""
    return _grid(events, num_bins, normalize, t0_us, t1_us).cpu().numpy()


def events_to_voxel_grid_cuda(events, num_bins, normalize=True, t0_us=None, t1_us=None, device="cuda:0"):
    """-> torch float32 (num_bins, height, width) on `device`."""
    with torch.cuda.device(torch.device(device)):
        return _grid(events, num_bins, normalize, t0_us, t1_us).to(device)


# ---- stateful per-pixel filters (reference :143-158, 169-200): same positional signatures, arrays updated in place ----
def _run_filter(kind, x, y, t, p, state_np, param):
    H, W = state_np.shape
    n = len(x)
    if n == 0:
        return np.zeros(0, bool)
    xs = np.asarray(x)
    ys = np.asarray(y)
    ev = one_window(xs, ys, np.asarray(t) if t is not None else np.arange(n, dtype=np.int64),
                    np.asarray(p) if p is not None else np.ones(n, np.int8), H, W)
    st = torch.from_numpy(np.ascontiguousarray(state_np)).to(ev.x.device)[None].contiguous()
    mask, st = eb.filter_events(ev, H, W, kind, param, st)
    state_np[...] = st[0].cpu().numpy()
    return mask.cpu().numpy().astype(bool)


def _refractory_period(mask, x, y, t, period, last_timestamp):
    """utils.py:193-200: mask[i] = False where t[i] - last_timestamp[y, x] < period, else last_timestamp[y, x] = t[i]"""
    keep = _run_filter("refractory", x, y, t, None, last_timestamp, period)
    mask[~keep] = False
    return mask


def _background_activity_filter(mask, timestamps, x, y, t, depth_us, radius=1):
    """utils.py:169-178: mask[i] = not (timestamps[y, x] > 0 and t - timestamps[y, x] > depth_us), then the block
    timestamps[y - radius : y + radius, x - radius : x + radius] = t"""
    H, W = timestamps.shape
    n = len(x)
    if n == 0:
        return mask
    ev = one_window(np.asarray(x), np.asarray(y), np.asarray(t), np.ones(n, np.int8), H, W)
    st = torch.from_numpy(np.ascontiguousarray(timestamps, dtype=np.float64)).to(ev.x.device)[None].contiguous()
    keep, st = eb.filter_events(ev, H, W, "background", float(depth_us), st, fx=int(radius))
    timestamps[...] = st[0].cpu().numpy()
    mask[...] = keep.cpu().numpy().astype(bool)
    return mask


def _contrast_threshold_control(activity, mask, x, y, p, factor):
    """utils.py:184-191: activity[y, x] += p; mask[i] = True and activity reset where |activity| >= factor"""
    keep = _run_filter("contrast", x, y, None, p, activity, factor)
    mask[keep] = True
    return mask


def _filter_events_resize(x, y, p, mask, change_map, fx, fy):
    """utils.py:143-158: per fx x fy cell, change += p / (fx fy); pass the event and subtract p when |change| >= 1"""
    H, W = change_map.shape
    n = len(x)
    if n:
        ev = one_window(np.asarray(x), np.asarray(y), np.arange(n, dtype=np.int64), np.asarray(p), H * fy, W * fx)
        st = torch.from_numpy(np.ascontiguousarray(change_map)).to(ev.x.device)[None].contiguous()
        keep, st = eb.filter_events(ev, H, W, "resize", 0.0, st, fx, fy)
        change_map[...] = st[0].cpu().numpy()
        mask[keep.cpu().numpy().astype(bool)] = True
    return mask, change_map
