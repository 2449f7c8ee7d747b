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


def to_host(dev_tensor, dtype=None, scale=None):
    """Device tensor -> numpy array of `dtype` (default: the tensor's own), optionally multiplied by `scale` in that dtype first.
    The cast and the scaling run on the GPU (float32 -> float64 is exact and a float64 product is the same IEEE operation as
    numpy's, so the values equal `t.cpu().numpy().astype(dtype) * scale`), and the copy lands in pinned host memory from
    torch's caching host allocator, which the returned array keeps alive: one DMA at link speed instead of a staged pageable
    copy followed by two passes over the array on one host core."""
    import torch
    d = dev_tensor if dtype is None or dev_tensor.dtype == dtype else dev_tensor.to(dtype)
    if scale is not None:
        d = d * scale
    if d.numel() * d.element_size() < (1 << 20):
        return d.cpu().numpy()
    h = torch.empty(d.shape, dtype=d.dtype, pin_memory=True)
    h.copy_(d)
    return h.numpy()


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


def one_window_structured(es, H, W, dev=None):
    """Fast path of `one_window` for the detection loaders' record layout - a C-contiguous structured array whose only fields
    are x, y, t, p, all `<i4` (gen1_2yolo.py:567-571 fix_events_training): ONE host-to-device copy of the records as they
    lie in memory, the field split, the range checks and the narrowing casts on the GPU, one 8-scalar read-back for the
    checks.  Same exceptions as `one_window`.  Returns None when the array is not of that layout (the caller then takes the
    general host path)."""
    import torch
    if not (isinstance(es, np.ndarray) and es.dtype.names and es.ndim == 1 and es.flags.c_contiguous and len(es)):
        return None
    names = es.dtype.names
    if sorted(names) != ["p", "t", "x", "y"] or any(es.dtype[k] != np.dtype("<i4") for k in names) or es.dtype.itemsize != 16:
        return None
    dev = dev or device()
    raw = torch.from_numpy(es.view(np.int32).reshape(len(es), 4)).to(dev)
    col = {k: raw[:, i] for i, k in enumerate(names)}
    x, y, t, p = col["x"], col["y"], col["t"], col["p"]
    lo = torch.stack([x.min(), y.min(), t.min(), p.min()])
    hi = torch.stack([x.max(), y.max(), t.max(), p.max()])
    (x0, y0, t0, p0), (x1, y1, t1, p1) = (v.tolist() for v in torch.stack([lo, hi]).cpu())
    if x0 < 0 or x1 >= W:
        raise IndexError(f"x coordinate outside [0, {W}) (min {x0}, max {x1})")
    if y0 < 0 or y1 >= H:
        raise IndexError(f"y coordinate outside [0, {H}) (min {y0}, max {y1})")
    if t1 - t0 >= T_SPAN_LIMIT:
        raise ValueError(f"timestamps span {t1 - t0} us; the GPU kernels need t.max() - t.min() < 2^30 us (about 17.9 min): "
                         "rescale or split the window")
    if p0 < -1 or p1 > 1:
        raise ValueError("polarities must be in {-1, 0, 1}")
    return eb.EventBatch(x.to(torch.int16), y.to(torch.int16), t.contiguous(), p.to(torch.int8), np.array([0, len(es)], np.int64))
