"""Upstream N-ImageNet representations on the GPU (n_imagenet/real_cnn_model/data/imagenet.py:169-870): the count /
latest-time / earliest-time / presence loaders the classifier's `loader_type` can select, each an instance of the
mixed-density per-pixel reductions and therefore ONE launch of evrep_mixed_density_batched.  Same names, arguments
(`event_tensor`: (N, 4) torch tensor [x, y, t, p], p in {-1, +1}; `augment`; `height` / `width` keywords) and return layout
(float32 torch CPU tensors, channels first) as the reference.

The six `reshape_then_{voxel_grid, optimized, event_stack, to_image, tore, time_surface}` wrappers of imagenet.py:1002-1134
are CALLERS of the hot path (caller ii) and stay in the user's own imagenet.py: they switch to this package
by changing the four imports at the top of that file ("N-ImageNet"); tests/nimagenet_callers.py holds that
call sequence for the parity tests.
"""
import numpy as np
import torch

IMAGE_H = 224
IMAGE_W = 224


# ---------------------------------------------------------------------------------------------------------------------
# Upstream N-ImageNet count / latest-timestamp representations (imagenet.py:169-343) that are instances of the same
# per-pixel reductions: each is a MixedDensityEventStack tuple over the whole window (window 0), so one CUDA launch
# through evrep_mixed_density_batched produces all of its planes.
#   bincount of positive / negative events          = (count_pos | count_neg, sum)
#   scatter_max of (t - t_first) / (t_last - t_first) = (timestamp_pos | timestamp_neg, max)   (untouched pixels 0 in both)
# Timestamps arrive as float seconds; only their position inside the window matters, so they go onto a 2^30 integer
# grid first (error < 1e-9 of the window).
#   scatter_min of the same normalised times          = (timestamp_pos | timestamp_neg, min)   (EVREP_AGG_MIN)
#   "some event touched the pixel" (reshape_then_flat*) = (count | count_pos | count_neg, max)
# The sorted / DiST variants (reshape_then_acc_sort, _acc_intensity, _acc_adj_sort, imagenet.py:513-999) are not built.
# ---------------------------------------------------------------------------------------------------------------------
_SPLITS_BY_POLARITY = {"count_pos", "count_neg", "timestamp_pos", "timestamp_neg", "polarity"}


def _window_reduce(event_tensor, H, W, functions, aggregations, need_time=True):
    from . import batched as eb
    from ._single import one_window
    ev_np = event_tensor.numpy() if torch.is_tensor(event_tensor) else np.asarray(event_tensor)
    if len(ev_np) == 0:  # torch.bincount(minlength=H * W) / scatter on no events: zeros
        return torch.zeros((H, W, len(functions)), dtype=torch.float32)
    if need_time:
        t = ev_np[:, 2].astype(np.float64)
        span = float(t[-1] - t[0])
        if not (span > 0 and np.all(np.diff(t) >= 0)):
            raise ValueError("the N-ImageNet time planes on the GPU need time-sorted events with t[-1] > t[0]")
        ti = np.rint((t - t[0]) / span * float(2**30 - 2)).astype(np.int64)
    else:  # count / presence planes: the timestamps are never read (imagenet.py:296-343, 397-438)
        ti = np.arange(len(ev_np), dtype=np.int64)
    p = ev_np[:, 3]
    if np.any(p == 0) and (_SPLITS_BY_POLARITY & set(functions)):
        raise ValueError("polarities must be -1 / +1 (imagenet.py splits on p > 0 / p < 0)")
    ev = one_window(ev_np[:, 0].astype(np.int64), ev_np[:, 1].astype(np.int64), ti, np.sign(p).astype(np.int8), H, W)
    return eb.mixed_density(ev, H, W, [0] * len(functions), functions, aggregations)[0]  # (H, W, C) float32 on the GPU


def _empty_guard(event_tensor):
    """imagenet.py:258-262: an empty sample becomes ten fake positive events at pixel (0, 0)"""
    if len(event_tensor) == 0:
        event_tensor = torch.zeros([10, 4]).float()
        event_tensor[:, 2] = torch.arange(10) / 10.0
        event_tensor[:, -1] = 1
    return event_tensor


def reshape_then_acc_count(event_tensor, augment=None, **kwargs):
    """imagenet.py:250-293 -> (4, H, W): positive count, latest positive time, negative count, latest negative time"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    event_tensor = _empty_guard(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["count_pos", "timestamp_pos", "count_neg", "timestamp_neg"], ["sum", "max", "sum", "max"])
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_acc(event_tensor, augment=None, **kwargs):
    """imagenet.py:169-210: like reshape_then_acc_count with each count plane divided by its maximum"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["count_pos", "timestamp_pos", "count_neg", "timestamp_neg"], ["sum", "max", "sum", "max"])
    rep = rep.permute(2, 0, 1).contiguous()
    rep[0] = rep[0] / rep[0].max()
    rep[2] = rep[2] / rep[2].max()
    return rep.float().cpu()


def reshape_then_acc_count_pol(event_tensor, augment=None, **kwargs):
    """imagenet.py:296-321 -> (2, H, W): positive count, negative count"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["count_pos", "count_neg"], ["sum", "sum"], need_time=False)
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_acc_count_only(event_tensor, augment=None, **kwargs):
    """imagenet.py:324-343 -> (1, H, W): events per pixel"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["count"], ["sum"], need_time=False)
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_acc_time(event_tensor, augment=None, **kwargs):
    """imagenet.py:213-247 -> (4, H, W): earliest / latest positive time, earliest / latest negative time"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["timestamp_pos", "timestamp_pos", "timestamp_neg", "timestamp_neg"], ["min", "max", "min", "max"])
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_acc_all(event_tensor, augment=None, **kwargs):
    """imagenet.py:346-394 -> (6, H, W): counts, latest times, earliest times (positive, negative each).  An empty sample
    gives zeros of the DEFAULT size, whatever height / width say (imagenet.py:353-354)."""
    if augment is not None:
        event_tensor = augment(event_tensor)
    if event_tensor.shape[0] == 0:
        return torch.zeros([6, IMAGE_H, IMAGE_W])
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["count_pos", "count_neg", "timestamp_pos", "timestamp_neg", "timestamp_pos", "timestamp_neg"],
                         ["sum", "sum", "max", "max", "min", "min"])
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_acc_time_pol(event_tensor, augment=None, **kwargs):
    """imagenet.py:475-510 -> (2, H, W): latest positive time, latest negative time"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    event_tensor = _empty_guard(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["timestamp_pos", "timestamp_neg"], ["max", "max"])
    return rep.permute(2, 0, 1).float().cpu()


EXP_TAU = 0.3  # imagenet.py:20


def reshape_then_acc_exp(event_tensor, augment=None, **kwargs):
    """imagenet.py:441-472 -> (2, H, W): exp(-(1 - latest time) / EXP_TAU) per polarity; untouched pixels count as time 0,
    like the reference (the exponential is taken of the whole scatter_max plane).  The exponential is an elementwise
    epilogue on the GPU tensor."""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["timestamp_pos", "timestamp_neg"], ["max", "max"])
    rep = torch.exp(-(1 - rep.double()) / EXP_TAU)
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_flat(event_tensor, augment=None, **kwargs):
    """imagenet.py:397-413 -> (1, H, W): 1 where any event fell (the reference augments AFTER reading the sizes)"""
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    if augment is not None:
        event_tensor = augment(event_tensor)
    if len(event_tensor) == 0:
        return torch.zeros([1, H, W])
    rep = _window_reduce(event_tensor, H, W, ["count"], ["max"], need_time=False)
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_flat_pol(event_tensor, augment=None, **kwargs):
    """imagenet.py:416-438 -> (2, H, W): 1 where a positive / a negative event fell"""
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    if augment is not None:
        event_tensor = augment(event_tensor)
    if len(event_tensor) == 0:
        return torch.zeros([2, H, W])
    rep = _window_reduce(event_tensor, H, W, ["count_pos", "count_neg"], ["max", "max"], need_time=False)
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_acc_intensity(event_tensor, augment=None, **kwargs):
    """imagenet.py:841-870 -> (1, H, W): positive minus negative count per pixel, min-max normalised over the plane (0 / 0 =
    NaN for a constant plane, as in the reference).  The difference is the ("polarity", "sum") plane; the normalisation is an
    elementwise epilogue on the GPU tensor."""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    if len(event_tensor) == 0:
        rep = torch.zeros((H, W, 1))
    else:
        rep = _window_reduce(event_tensor, H, W, ["polarity"], ["sum"], need_time=False)
    lo, hi = rep.min(), rep.max()
    return ((rep - lo) / (hi - lo)).permute(2, 0, 1).float().cpu()


def loader_for(loader_type):
    """ImageNetDataset.__init__'s choice of loader (imagenet.py:1232-1260) for the upstream representations this module
    provides: `loader_type` string -> function.  The sorted / DiST loaders ("sorted_time_surface", "dist", ...) are not built and
    raise NotImplementedError; any other string gives None (the six `reshape_then_*` wrappers of imagenet.py:1261-1272 live in
    the caller's file, and the reference leaves `self.loader` unset for unknown strings)."""
    table = [
        ((None, "event_image", "reshape_then_acc"), reshape_then_acc),
        (("reshape_then_acc_time",), reshape_then_acc_time),
        (("reshape_then_acc_count",), reshape_then_acc_count),
        (("reshape_then_acc_all",), reshape_then_acc_all),
        (("reshape_then_flat_pol",), reshape_then_flat_pol),
        (("binary_event_image", "reshape_then_flat"), reshape_then_flat),
        (("timestamp_image", "reshape_then_acc_time_pol"), reshape_then_acc_time_pol),
        (("event_histogram", "reshape_then_acc_count_pol"), reshape_then_acc_count_pol),
        (("reshape_then_acc_exp",), reshape_then_acc_exp),
        (("reshape_then_acc_intensity",), reshape_then_acc_intensity),
    ]
    for names, fn in table:
        if loader_type in names:
            return fn
    if loader_type in ("sorted_time_surface", "reshape_then_acc_sort", "dist", "DiST", "reshape_then_acc_adj_sort"):
        raise NotImplementedError(f"loader_type {loader_type!r}: the sorted / DiST representations (imagenet.py:513-999) are not built")
    return None
