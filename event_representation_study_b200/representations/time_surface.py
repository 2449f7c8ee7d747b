"""Mirror of representations/time_surface.py (reference :7-74)."""
from dataclasses import dataclass
from typing import Tuple, Union

import numpy as np

from .. import batched as eb
import torch

from .._single import one_window, to_host


def _surfaces(x, y, t, p, indices, n_pol, H, W, tau):
    p = np.asarray(p)
    if n_pol != 2 or (p.size and (p.min() < 0 or p.max() > 1)):
        raise ValueError("the GPU time surface supports sensor_size[2] == 2 with polarity indices in {0, 1}")
    t, tau = _integer_time(t, tau)
    ev = one_window(x, y, t, p, H, W, require_sorted=True)
    idx = np.asarray(indices, np.int64).reshape(1, -1)
    return to_host(eb.time_surface(ev, H, W, idx.shape[1], float(tau), indices=idx)[0], torch.float64)


def _integer_time(t, tau):
    """The kernels take integer timestamps; the reference decays on the float values unchanged (time_surface.py:66-72).
    Fractional timestamps (float seconds of the N-ImageNet path, sub-microsecond stamps) are rescaled, together with tau,
    onto a grid of 2^30 steps over the window, so that exp((t_i - t_j) / tau) keeps its value to ~1e-9 of the window / tau."""
    t = np.asarray(t)
    if t.dtype.kind != "f" or t.size == 0 or np.all(t == np.rint(t)):
        return t, tau
    span = float(t.max() - t.min())
    if not span > 0:
        return np.zeros(t.shape, np.int64), tau
    scale = float(2**30 - 2) / span
    return np.rint((t - t.min()) * scale).astype(np.int64), float(tau) * scale


@dataclass(frozen=True)
class ToTimesurface:
    """Global exponential time surfaces (HOTS); same constructor and call as the reference."""

    sensor_size: Tuple[int, int, int]
    surface_dimensions: Union[None, Tuple[int, int]] = None
    tau: float = 5e3
    decay: str = "lin"

    def __call__(self, events, indices):
        W, H, P = self.sensor_size
        if len(indices) == 0:
            return np.zeros((0, P, H, W))
        return _surfaces(events["x"], events["y"], events["t"], events["p"], indices, P, H, W, self.tau)


def to_timesurface_numpy(x, y, t, p, indices, timestamp_memory, all_surfaces, tau=5e3):
    """In-place variant of the reference (time_surface.py:52-74): fills `all_surfaces`; `timestamp_memory` receives the
    last timestamp per (polarity, y, x) over the events the reference's loop visits."""
    P, H, W = timestamp_memory.shape
    if not np.all(timestamp_memory == timestamp_memory.flat[0]) or not (timestamp_memory.flat[0] <= -(3 * tau)):
        raise ValueError("to_timesurface_numpy on the GPU supports the reference's own initial memory (a constant <= -3 tau, "
                         "time_surface.py:26-29): pixels without an event contribute exp(-3 - ...) ~ 0 either way")
    all_surfaces[...] = _surfaces(x, y, t, p, indices, P, H, W, tau) if len(indices) else 0.0
    # events visited: up to and including the last strictly increasing, in-range index (the loop breaks there)
    n, prev, n_valid = len(t), -1, 0
    for i in indices:
        if i <= prev or i >= n:
            break
        prev, n_valid = int(i), n_valid + 1
    last = prev + 1 if n_valid == len(indices) else n
    import torch
    dev = "cuda"
    lin = (torch.as_tensor(np.asarray(p[:last]).astype(np.int64), device=dev) * H + torch.as_tensor(np.asarray(y[:last]).astype(np.int64), device=dev)) * W \
        + torch.as_tensor(np.asarray(x[:last]).astype(np.int64), device=dev)
    order = torch.arange(last, device=dev)
    latest = torch.full((P * H * W,), -1, dtype=torch.int64, device=dev).scatter_reduce_(0, lin, order, "amax", include_self=True)
    mem = torch.as_tensor(timestamp_memory.reshape(-1), device=dev).clone()
    has = latest >= 0
    mem[has] = torch.as_tensor(np.asarray(t[:last]).astype(np.float64), device=dev)[latest[has]]
    timestamp_memory[...] = mem.cpu().numpy().reshape(P, H, W)
