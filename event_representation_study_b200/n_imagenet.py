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
# The two rank-based loaders (reshape_then_acc_sort: "sorted time surface", reshape_then_acc_adj_sort: DiST,
# imagenet.py:513-999) take their per-pixel counts and latest / earliest stamps from the same launch with the DENSE RANK of
# the timestamp as integer time (exact: a rank below 2^21 survives the kernel's float32 normalisation and is rounded back),
# and do their image-domain steps (pooling, clipping, ranking of H x W values) with torch on the GPU like the reference does
# on the CPU.
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


TIME_SCALE = 1000000    # imagenet.py:21
CLIP_COUNT_RATE = 0.99  # imagenet.py:24
DISC_ALPHA = 3.0        # imagenet.py:25
_MAX_RANKS = 1 << 21


def _rank_planes(ev_np, key, H, W, functions, aggregations):
    """One mixed-density launch over the sample with the dense rank of `key` (non-decreasing, one entry per event) as the
    timestamp.  -> (planes (H, W, C) on the GPU: counts as float32, timestamp planes as int64 RANKS (0 for untouched pixels,
    like scatter_max / scatter_min leave them), uniq: the sorted distinct keys, so that uniq[rank] is the stamp itself)."""
    from . import batched as eb
    from ._single import one_window
    if np.any(np.diff(key) < 0):
        raise ValueError("the rank-based N-ImageNet loaders on the GPU need time-sorted events")
    p = ev_np[:, 3]
    if np.any(p == 0):
        raise ValueError("polarities must be -1 / +1 (imagenet.py splits on p > 0 / p < 0)")
    uniq, rank = np.unique(key, return_inverse=True)
    R = len(uniq)
    if R > _MAX_RANKS:
        raise NotImplementedError(f"{R} distinct timestamps in one sample; the rank planes are exact up to {_MAX_RANKS}")
    is_time = [f.startswith("timestamp") for f in functions]
    # all stamps equal: every rank is 0 and the kernel's (t - t_min) / (t_max - t_min) would be 0 / 0
    t_int = rank.astype(np.int64) if R > 1 else np.arange(len(rank), dtype=np.int64)
    ev = one_window(ev_np[:, 0].astype(np.int64), ev_np[:, 1].astype(np.int64), t_int, np.sign(p).astype(np.int8), H, W)
    rep = eb.mixed_density(ev, H, W, [0] * len(functions), functions, aggregations)[0]
    planes = []
    for c, tplane in enumerate(is_time):
        if not tplane:
            planes.append(rep[:, :, c])
        elif R > 1:
            planes.append(torch.round(rep[:, :, c].double() * float(R - 1)).long())
        else:
            planes.append(torch.zeros((H, W), dtype=torch.long, device=rep.device))
    return planes, uniq


def _dev_scalar(v, like, dtype=None):
    """A 0-dim tensor on `like`'s device.  torch's CUDA kernels turn `tensor / python_scalar` into a multiplication by the
    reciprocal (one rounding more than the CPU's true division the reference runs); dividing by a device tensor does not."""
    return torch.tensor(v, dtype=dtype or like.dtype, device=like.device)


def _dense_rank_normalised(values, touched):
    """imagenet.py:567-583: the kept events (one per touched pixel, in stream order) are ranked by their distinct stamps,
    1-based, then min-max normalised -> rank / (K - 1) as float32, 0 when there is one distinct value; untouched pixels 0"""
    out = torch.zeros(values.shape, dtype=torch.float32, device=values.device)
    if bool(touched.any()):
        u, inv = torch.unique(values[touched], return_inverse=True)
        if u.numel() > 1:
            out[touched] = inv.float() / _dev_scalar(float(u.numel() - 1), out)
    return out


def _hot_check(plane):
    """imagenet.py:588-590, 752-754: `hot = sort[sort > 0]; hot.max()` - the normalised copy is never written back, but the
    reduction raises on a plane without a positive entry, and so does this"""
    if not bool((plane > 0).any()):
        raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0. Specify the reduction dim with the 'dim' argument.")


def _quantize(sort, q):
    """imagenet.py:606-621 / 783-809"""
    if q is None:
        return sort
    if type(q) == int:  # noqa: E721 (the reference's own test)
        return torch.round(sort * q) / _dev_scalar(q, sort)
    if type(q) == list:  # noqa: E721
        return torch.stack([torch.round(sort * k) / _dev_scalar(k, sort) for k in q], dim=2)
    return sort


def reshape_then_acc_sort(event_tensor, augment=None, **kwargs):
    """imagenet.py:513-838, the "sorted time surface": per pixel (and polarity) the latest event's stamp - the dense rank of
    its microsecond (global_time) or the microsecond itself - raw, or with `strict` re-ranked among the per-pixel winners and
    normalised; optionally a presence image in front and quantised copies.  Keywords as the dataset passes them
    (imagenet.py:1287-1299).  `denoise_image` / `denoise_sort` call a function the reference never defines: NameError, here too.
    The per-polarity ranks the reference computes for global_time=False are discarded by it (:527-537): the stamps are then the
    raw microseconds."""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    ev = event_tensor.numpy() if torch.is_tensor(event_tensor) else np.asarray(event_tensor)
    ev = ev.astype(np.float64)
    time_idx = (ev[:, 2] * TIME_SCALE).astype(np.int64)  # .long(): truncation
    if kwargs["neglect_polarity"]:
        (count, last), uniq = _rank_planes(ev, time_idx, H, W, ["count", "timestamp"], ["sum", "max"])
        groups = [(count, last)]
    else:
        pos_n, neg_n = int((ev[:, 3] > 0).sum()), int((ev[:, 3] < 0).sum())
        (cp, cn, lp, ln), uniq = _rank_planes(ev, time_idx, H, W, ["count_pos", "count_neg", "timestamp_pos", "timestamp_neg"],
                                              ["sum", "sum", "max", "max"])
        groups = []
        for n_ev, c, l in ((pos_n, cp, lp), (neg_n, cn, ln)):
            if n_ev == 0:  # imagenet.py:641-646: a polarity without events becomes ONE fake event at pixel (0, 0), stamp 0
                c, l = torch.zeros_like(c), torch.zeros_like(l)
                c[0, 0] = 1.0
            groups.append((c, l))
    dev = groups[0][0].device
    uniq_t = torch.as_tensor(uniq, dtype=torch.float64, device=dev)
    planes = []
    for count, last in groups:
        touched = count > 0
        if kwargs["use_image"]:
            if kwargs["denoise_image"]:
                raise NameError("name 'density_filter_event_image' is not defined")
            planes.append(touched.float())
        # the value scatter_max sees: the rank itself (global_time), else the microsecond stamp; 0 where no event fell
        value = last.double() if kwargs["global_time"] else torch.where(touched, uniq_t[last], torch.zeros((), dtype=torch.float64, device=dev))
        if kwargs["strict"]:
            sort = _dense_rank_normalised(value, touched)
        else:
            _hot_check(value)
            sort = value
        if kwargs["denoise_sort"]:
            raise NameError("name 'density_filter_event_image' is not defined")
        planes.append(_quantize(sort, kwargs["quantize_sort"]))
    if planes[-1].dim() == 2:
        result = torch.stack(planes, dim=2)
    else:
        result = torch.cat([q.unsqueeze(-1) if q.dim() == 2 else q for q in planes], dim=2)
    return result.permute(2, 0, 1).float().cpu()


def reshape_then_acc_adj_sort(event_tensor, augment=None, **kwargs):
    """imagenet.py:873-999, DiST: latest normalised stamp per pixel and polarity, discounted by the temporal spread of its 5 x 5
    neighbourhood over the (clipped) neighbourhood count, then replaced by its dense rank among the H x W values.  Counts and
    latest / earliest stamps come from one kernel launch; the 5 x 5 pooling and the ranking of the image are torch ops on the
    GPU, the same ones the reference runs on the CPU (float32, unfused), so the planes agree to the bit."""
    if augment is not None:
        event_tensor = augment(event_tensor)
    H = kwargs.get("height", IMAGE_H)
    W = kwargs.get("width", IMAGE_W)
    ev = event_tensor.numpy() if torch.is_tensor(event_tensor) else np.asarray(event_tensor)
    ev = ev.astype(np.float64)
    t = ev[:, 2]
    (cp, cn, lp, ln, ep, en), uniq = _rank_planes(ev, t, H, W, ["count_pos", "count_neg", "timestamp_pos", "timestamp_neg", "timestamp_pos", "timestamp_neg"],
                                                  ["sum", "sum", "max", "max", "min", "min"])
    dev = cp.device
    uniq_t = torch.as_tensor(uniq, dtype=torch.float64, device=dev)
    start_time, time_length = float(t[0]), float(t[-1]) - float(t[0])
    patch = 5
    outs = []
    for count, last, first in ((cp, lp, ep), (cn, ln, en)):
        count = count.clone()
        touched = count > 0
        zero = torch.zeros((), dtype=torch.float64, device=dev)
        norm = lambda r: torch.where(touched, (uniq_t[r] - start_time) / _dev_scalar(time_length, uniq_t), zero).float()  # noqa: E731  scatter leaves 0
        out, min_out = norm(last), norm(first)
        # clip the counts at the number of distinct count values that cover less than 99 % of the pixels (:894-902)
        unique_count = torch.unique(count, return_counts=True)[1]
        sum_subset = torch.cumsum(unique_count, dim=0)
        th_clip = int((sum_subset < H * W * CLIP_COUNT_RATE).sum())
        count[count > th_clip] = th_clip
        min_out[count == 0] = 1.0
        neighbor_count = patch**2 * torch.nn.functional.avg_pool2d(count.unsqueeze(0), patch, stride=1, padding=patch // 2)
        disc = (torch.nn.functional.max_pool2d(out.unsqueeze(0), patch, stride=1, padding=patch // 2)
                + torch.nn.functional.max_pool2d(-min_out.unsqueeze(0), patch, stride=1, padding=patch // 2)) / neighbor_count
        live = count > 0
        out[live] = out[live] - DISC_ALPHA * disc.squeeze()[live]
        out[out < 0] = 0
        out[neighbor_count.squeeze() == 1.0] = 0
        flat = out.reshape(H * W)
        unq, inv = torch.unique(flat, return_inverse=True)  # sorted distinct values: `inv` is the rank torch.sort + unique_consecutive give
        outs.append((inv.float() / _dev_scalar(float(unq.shape[0]), out)).reshape(H, W))
    return torch.stack(outs, dim=2).permute(2, 0, 1).float().cpu()


def loader_for(loader_type):
    """ImageNetDataset.__init__'s choice of loader (imagenet.py:1232-1260) for the upstream representations this module
    provides: `loader_type` string -> function.  Any other string gives None (the six `reshape_then_*` wrappers of
    imagenet.py:1261-1272 live in the caller's file, and the reference leaves `self.loader` unset for unknown strings)."""
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
        (("sorted_time_surface", "reshape_then_acc_sort"), reshape_then_acc_sort),
        (("dist", "DiST", "reshape_then_acc_adj_sort"), reshape_then_acc_adj_sort),
    ]
    for names, fn in table:
        if loader_type in names:
            return fn
    return None
