"""Helpers shared by the per-window drop-in classes: one window of host events -> EventBatch on the GPU."""
import numpy as np
import torch

from . import batched as eb


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("event_representation_study_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_u16_coords(v, name, limit):
    """Integer-valued coordinate array -> uint16 values stored as int16 (the C ABI's layout).
    Out-of-range coordinates raise IndexError like the reference's numpy / torch_scatter indexing does."""
    a = np.asarray(v)
    if a.size and (a.min() < 0 or a.max() >= limit):
        raise IndexError(f"{name} coordinate outside [0, {limit}) (min {a.min()}, max {a.max()})")
    return a.astype(np.uint16).view(np.int16)


T_SPAN_LIMIT = 2**30  # the kernels key events by t - t_first in 31 bits (include/evrep.h, EVREP_WF_T_RANGE)


def one_window(x, y, t, p, H, W, dev=None, require_sorted=False):
    """Arrays of one window -> EventBatch with B = 1 (t as int64 when it does not fit int32).

    Everything the kernels would only FLAG per window (include/evrep.h EVREP_WF_*) is checked here on the host and raised,
    so that a per-window drop-in never returns a tensor with silently dropped events: coordinates outside the sensor
    (IndexError, like the reference's indexing), polarities outside {-1, 0, 1}, a timestamp span of 2^30 us or more
    (the reference handles any int64 span; rescale the timestamps or use the batched API and read window_flags), and, for
    the order-dependent representations, unsorted timestamps."""
    dev = dev or device()
    t = np.asarray(t)
    if t.dtype.kind == "f":
        t = t.astype(np.int64)  # the reference's own .astype(np.int64) (mixed_density_event_stack.py:29)
    t = t.astype(np.int64, copy=False)
    if t.size and int(t.max()) - int(t.min()) >= T_SPAN_LIMIT:
        raise ValueError(f"timestamps span {int(t.max()) - int(t.min())} us; the GPU kernels need t.max() - t.min() < 2^30 us (about 17.9 min): "
                         "rescale or split the window")
    if require_sorted and t.size > 1 and np.any(t[1:] < t[:-1]):
        raise ValueError("this representation depends on the event order: the window must be time sorted")
    t_dtype = np.int32 if (t.size == 0 or (t.min() >= -2**31 and t.max() < 2**31)) else np.int64
    pp = np.asarray(p)
    if pp.size and (pp.min() < -1 or pp.max() > 1):
        raise ValueError("polarities must be in {-1, 0, 1}")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return eb.EventBatch(up(as_u16_coords(x, "x", W)), up(as_u16_coords(y, "y", H)), up(t.astype(t_dtype)), up(pp.astype(np.int8)),
                         np.array([0, len(t)], np.int64))
