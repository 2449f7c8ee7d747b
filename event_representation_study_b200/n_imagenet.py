"""Drop-in mirrors of the N-ImageNet loader wrappers around the hot path (caller ii):
n_imagenet/real_cnn_model/data/imagenet.py:1002-1134 - fix_events_training and reshape_then_{voxel_grid, optimized,
event_stack, to_image, tore, time_surface}.  Same names, arguments (`event_tensor`: (N, 4) torch tensor [x, y, t, p],
p in {-1, +1}; `augment`; `height` / `width` keywords) and return layout (float32 torch CPU tensors, channels first
except to_image, which the reference leaves channels last); each one calls the same representation entry point as the
reference, here the CUDA-backed mirror.

Two reference lines cannot run on current numpy / at all and are mirrored by intent, not by exception:
  * reshape_then_to_image ends with `rep.float()` on a numpy array (AttributeError in the reference, imagenet.py:1077);
    here the array is converted to a float32 tensor like every other wrapper does.
  * reshape_then_time_surface uses `np.int`, removed in numpy 1.24 (imagenet.py:1125-1126); plain `int` is used.
"""
import numpy as np
import numpy.lib.recfunctions as rfn
import torch

from . import tonic_compat as tonic_transforms
from .representations.event_stack import EventStack
from .representations.optimized_representation import get_optimized_representation
from .representations.time_surface import ToTimesurface
from .representations.tore import events2ToreFeature

IMAGE_H = 224
IMAGE_W = 224


def fix_events_training(events):
    """imagenet.py:1002-1006: (N, 4) float64 array -> structured array with f8 fields x, y, t, p"""
    events = rfn.unstructured_to_structured(events)
    events.dtype = [("x", "<f8"), ("y", "<f8"), ("t", "<f8"), ("p", "<f8")]
    return events


def reshape_then_voxel_grid(event_tensor, augment=None, **kwargs):
    """imagenet.py:1009-1022"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    transformation = tonic_transforms.ToVoxelGrid((W, H, 2), n_time_bins=12)
    reshaped_return_data = fix_events_training(event_tensor.numpy())
    rep = transformation(reshaped_return_data)
    rep = torch.tensor(rep.transpose(0, 2, 3, 1)[..., 0])
    return rep.float()


def reshape_then_optimized(event_tensor, augment=None, **kwargs):
    """imagenet.py:1025-1039"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    reshaped_return_data = fix_events_training(event_tensor.numpy())
    rep = get_optimized_representation(reshaped_return_data, reshaped_return_data.shape[0], H, W)
    rep = torch.tensor(rep.transpose(2, 0, 1))
    return rep.float()


def reshape_then_event_stack(event_tensor, augment=None, **kwargs):
    """imagenet.py:1042-1060"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    reshaped_return_data = fix_events_training(event_tensor.numpy())
    reshaped_return_data["p"] = (reshaped_return_data["p"] + 1) // 2
    stack_size = 12
    transformation = EventStack(stack_size, reshaped_return_data.shape[0], H, W)
    pre_stack = transformation.pre_stack(reshaped_return_data, reshaped_return_data[-1]["t"])
    post_stack = transformation.post_stack(pre_stack)
    rep = torch.tensor(post_stack.transpose(3, 0, 1, 2)[..., 0])
    return rep.float()


def reshape_then_to_image(event_tensor, augment=None, **kwargs):
    """imagenet.py:1063-1077 (see the module docstring for the last line)"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    reshaped_return_data = fix_events_training(event_tensor.numpy())
    transformation = tonic_transforms.ToImage((W, H, 2))
    reshaped_return_data["p"] = (reshaped_return_data["p"] + 1) // 2
    rep = transformation(reshaped_return_data)
    rep = rep.transpose(1, 2, 0)
    return torch.tensor(np.ascontiguousarray(rep)).float()


def reshape_then_tore(event_tensor, augment=None, **kwargs):
    """imagenet.py:1080-1107"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    k = 6
    reshaped_return_data = fix_events_training(event_tensor.numpy())
    x, y, ts, pol = (reshaped_return_data["x"], reshaped_return_data["y"], reshaped_return_data["t"], reshaped_return_data["p"])
    x = x - min(x) + 1
    y = y - min(y) + 1
    sampleTimes = ts[-1]
    frameSize = (H, W)
    rep = events2ToreFeature(x, y, ts, pol, sampleTimes, k, frameSize)
    rep = torch.tensor(rep.transpose(2, 0, 1))
    return rep.float()


def reshape_then_time_surface(event_tensor, augment=None, **kwargs):
    """imagenet.py:1110-1134"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    reshaped_return_data = fix_events_training(event_tensor.numpy())
    reshaped_return_data["p"] = ((reshaped_return_data["p"] + 1) / 2).astype(np.int8)
    transform = ToTimesurface(sensor_size=(W, H, 2), surface_dimensions=None, tau=50000, decay="exp")
    t = reshaped_return_data["t"]
    t_norm = (t - t[0]) / (t[-1] - t[0]) * 6
    idx = np.searchsorted(t_norm, np.arange(6) + 1)
    reshaped_return_data["x"] = reshaped_return_data["x"].astype(int)
    reshaped_return_data["y"] = reshaped_return_data["y"].astype(int)
    rep = transform(reshaped_return_data, idx)
    rep = rep.reshape((-1, rep.shape[-2], rep.shape[-1]))
    rep = torch.tensor(rep.transpose(1, 2, 0)) if not torch.is_tensor(rep) else rep.permute(1, 2, 0)
    return rep.float()


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
def _window_reduce(event_tensor, H, W, functions, aggregations, need_time=True):
    from . import batched as eb
    from ._single import one_window
    ev_np = event_tensor.numpy() if torch.is_tensor(event_tensor) else np.asarray(event_tensor)
    if need_time:
        t = ev_np[:, 2].astype(np.float64)
        span = float(t[-1] - t[0])
        if not (span > 0 and np.all(np.diff(t) >= 0)):
            raise ValueError("the N-ImageNet representations on the GPU need time-sorted events with t[-1] > t[0]")
        ti = np.rint((t - t[0]) / span * float(2**30 - 2)).astype(np.int64)
    else:  # presence planes: the timestamps are never read (imagenet.py:397-438)
        ti = np.arange(len(ev_np), dtype=np.int64)
    p = ev_np[:, 3]
    if np.any(p == 0):
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
    rep = _window_reduce(event_tensor, H, W, ["count_pos", "count_neg"], ["sum", "sum"])
    return rep.permute(2, 0, 1).float().cpu()


def reshape_then_acc_count_only(event_tensor, augment=None, **kwargs):
    """imagenet.py:324-343 -> (1, H, W): events per pixel"""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    rep = _window_reduce(event_tensor, H, W, ["count"], ["sum"])
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
    """ImageNetDataset.__init__'s choice of loader (imagenet.py:1232-1272): `loader_type` string -> function.  The two sorted /
    DiST loaders ("sorted_time_surface", "dist", ...) are not built and raise NotImplementedError; an unknown string gives
    None, where the reference leaves `self.loader` unset."""
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
        (("reshape_then_voxel_grid",), reshape_then_voxel_grid),
        (("reshape_then_optimized",), reshape_then_optimized),
        (("reshape_then_event_stack",), reshape_then_event_stack),
        (("reshape_then_to_image",), reshape_then_to_image),
        (("reshape_then_tore",), reshape_then_tore),
        (("reshape_then_time_surface",), reshape_then_time_surface),
    ]
    for names, fn in table:
        if loader_type in names:
            return fn
    if loader_type in ("sorted_time_surface", "reshape_then_acc_sort", "dist", "DiST", "reshape_then_acc_adj_sort"):
        raise NotImplementedError(f"loader_type {loader_type!r}: the sorted / DiST representations (imagenet.py:513-999) are not built")
    return None
