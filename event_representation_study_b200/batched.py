"""Batched engine API: many event windows -> dense tensors in one call per representation.

This is the call the B200 engine is built around: the caller collates raw events of B
windows into CSR-packed SoA arrays (`EventBatch`), already on the GPU or uploaded once, and each function
below enqueues the kernels of libevrep.so on the current CUDA stream and returns a CUDA tensor in the
layout the reference produces for one window, with a leading batch axis.  The per-window classes in
`event_representation_study_b200.representations` are thin wrappers over these functions.

Nothing here computes on the CPU; without a CUDA device or without the built library every call raises.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import AGGS, FUNCS, STACKING, check, lib

ERGO12_V2 = (
    [0, 3, 2, 6, 5, 6, 2, 5, 1, 0, 4, 1],
    ["polarity", "timestamp_neg", "count_neg", "polarity", "count_pos", "count", "timestamp_pos", "count_neg",
     "timestamp_neg", "timestamp_pos", "timestamp", "count"],
    ["variance", "variance", "mean", "sum", "mean", "sum", "mean", "mean", "max", "max", "max", "mean"],
)  # representations/optimized_representation.py:86-115
ERGO12_V1 = (
    [0, 2, 2, 3, 5, 0, 0, 4, 2, 6, 1, 1],
    ["timestamp", "timestamp_pos", "timestamp_neg", "count_neg", "count_pos", "polarity", "timestamp", "count",
     "timestamp_pos", "count", "timestamp_pos", "timestamp_neg"],
    ["max", "sum", "mean", "sum", "mean", "variance", "variance", "sum", "mean", "sum", "sum", "sum"],
)  # representations/optimized_representation.py:16-66 (commented-out first version)


@dataclass
class EventBatch:
    """CSR-packed SoA events of B windows.  x, y: uint16 values (torch.uint16 or int16 storage); t: int32 or
    int64 microseconds; p: int8 in {-1, 0, +1}; all CUDA, contiguous, same length.  offsets: host int64
    numpy array of B+1 entries; window b is [offsets[b], offsets[b+1]).  Events of a window are in stream
    order (time sorted), as every loader of the reference delivers them."""
    x: torch.Tensor
    y: torch.Tensor
    t: torch.Tensor
    p: torch.Tensor
    offsets: np.ndarray

    def __post_init__(self):
        self.offsets = np.ascontiguousarray(np.asarray(self.offsets, dtype=np.int64))
        n = int(self.offsets[-1]) if len(self.offsets) else 0
        for name, ok in (("x", (torch.uint16, torch.int16)), ("y", (torch.uint16, torch.int16)),
                         ("t", (torch.int32, torch.int64)), ("p", (torch.int8,))):
            v = getattr(self, name)
            if not v.is_cuda:
                raise ValueError(f"EventBatch.{name} must be a CUDA tensor (there is no CPU path)")
            if v.dtype not in ok:
                raise TypeError(f"EventBatch.{name} has dtype {v.dtype}, expected one of {ok}")
            if not v.is_contiguous() or v.dim() != 1:
                raise ValueError(f"EventBatch.{name} must be 1-D contiguous")
            if v.numel() < n:
                raise ValueError(f"EventBatch.{name} has {v.numel()} elements, offsets need {n}")
        if len(self.offsets) < 1 or np.any(np.diff(self.offsets) < 0) or (len(self.offsets) and self.offsets[0] < 0):
            raise ValueError("offsets must be non-decreasing and non-negative")

    @property
    def B(self):
        return len(self.offsets) - 1

    @property
    def device(self):
        return self.x.device

    @property
    def total(self):
        return int(self.offsets[-1])


def pack_events(windows, device="cuda", t_dtype=np.int32, pin=False):
    """List of windows -> EventBatch on `device` (one host->device copy per field).
    A window is a dict / structured array with fields x, y, t, p (any integer dtypes)."""
    offs = np.zeros(len(windows) + 1, np.int64)
    offs[1:] = np.cumsum([len(w["x"]) for w in windows])

    def cat(k, dt):
        if not windows:
            return np.zeros(0, dt)
        return np.concatenate([np.asarray(w[k]).astype(dt, copy=False) for w in windows])

    def up(a):
        h = torch.from_numpy(a)
        if pin:
            h = h.pin_memory()
        return h.to(device, non_blocking=pin)

    return EventBatch(up(cat("x", np.uint16).view(np.int16)), up(cat("y", np.uint16).view(np.int16)), up(cat("t", t_dtype)),
                      up(cat("p", np.int8)), offs)


def split_collated(events, num_windows=None):
    """The reference's raw-event mini-batch - rows [x, y, t, p, b] with b the sample index, as `collate_fn` builds it for the
    learned representation (ev-YOLOv6/yolov6/data/gen1_2yolo.py:433-445) and `QuantizationLayer.forward` consumes it
    (models/learned_repr.py:143-156) - split into the SoA fields and CSR offsets of an EventBatch, on the tensor's own device:
    -> (x, y, t, p, offsets).  Rows must be grouped by ascending b (the collate concatenates the samples in order);
    t is truncated to integer microseconds like the reference's own `.astype(np.int64)` and narrowed to int32 when it fits;
    `num_windows` keeps trailing empty samples (default: b.max() + 1)."""
    ev = torch.as_tensor(events)
    if ev.dim() != 2 or ev.shape[1] != 5:
        raise ValueError(f"expected rows [x, y, t, p, b]: shape (N, 5), got {tuple(ev.shape)}")
    n = ev.shape[0]
    b = ev[:, 4].to(torch.int64)
    if n and bool((b[1:] < b[:-1]).any()):
        raise ValueError("rows must be grouped by sample index b in ascending order")
    B = int(num_windows) if num_windows is not None else (int(b[-1]) + 1 if n else 0)
    if n and (int(b[0]) < 0 or int(b[-1]) >= B):
        raise ValueError(f"sample indices must lie in [0, {B})")
    offsets = np.zeros(B + 1, np.int64)
    if n:
        offsets[1:] = torch.bincount(b, minlength=B).cumsum(0).cpu().numpy()
    xy = ev[:, :2].to(torch.int64)
    if n and (int(xy.min()) < 0 or int(xy.max()) > 65535):
        raise IndexError("x / y outside the uint16 range of the event layout")
    u16 = lambda v: (v.to(torch.int32) & 0xFFFF).to(torch.int16).contiguous()  # noqa: E731  (uint16 values in int16 storage)
    t = ev[:, 2].to(torch.int64)
    if n == 0 or (int(t.min()) >= -2**31 and int(t.max()) < 2**31):
        t = t.to(torch.int32)
    p = ev[:, 3].to(torch.int64)
    if n and (int(p.min()) < -1 or int(p.max()) > 1):
        raise ValueError("polarities must be in {-1, 0, 1}")
    return u16(xy[:, 0]), u16(xy[:, 1]), t.contiguous(), p.to(torch.int8).contiguous(), offsets


def from_collated(events, num_windows=None, device="cuda"):
    """`split_collated` + upload: the (N, 5) [x, y, t, p, b] tensor of the reference's collate -> EventBatch on `device`
    (the batched call site the learned representation already has)."""
    x, y, t, p, offsets = split_collated(events, num_windows)
    return EventBatch(x.to(device), y.to(device), t.to(device), p.to(device), offsets)


_workspaces = {}


def _workspace(device, stream_ptr, nbytes):
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _prep(ev, op, H, W, C):
    if not torch.cuda.is_available():
        raise RuntimeError("event_representation_study_b200 needs a CUDA device; there is no CPU fallback")
    _require_current(ev.device)
    stream = torch.cuda.current_stream(ev.device).cuda_stream
    nbytes = lib.evrep_workspace_bytes(op, ev.B, ev.total, H, W, C)
    if nbytes == 0:
        raise ValueError(f"invalid geometry B={ev.B} total={ev.total} H={H} W={W}")
    ws = _workspace(ev.device, stream, nbytes)
    head = (ev.x.data_ptr(), ev.y.data_ptr(), ev.t.data_ptr(), ev.t.element_size(), ev.p.data_ptr(), ev.offsets.ctypes.data,
            ev.B, H, W)
    return head, ws, stream


def _require_current(device):
    """libevrep launches on the calling thread's current device: refuse tensors that live elsewhere instead of failing
    inside CUDA with an invalid resource handle"""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        raise ValueError(f"the events live on cuda:{idx} but the current device is cuda:{torch.cuda.current_device()}: wrap the call in "
                         f"`with torch.cuda.device({idx}):`")


def _out(ev, shape, out):
    if out is None:
        return torch.empty(shape, dtype=torch.float32, device=ev.device)
    if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != tuple(shape) or out.device != ev.device:
        raise ValueError(f"out must be a contiguous float32 tensor of shape {tuple(shape)} on {ev.device}")
    return out


def window_flags(ev):
    """EVREP_WF_* status word of every window for the last op run on the current stream (synchronises)."""
    stream = torch.cuda.current_stream(ev.device).cuda_stream
    key = (ev.device.index if ev.device.index is not None else torch.cuda.current_device(), stream)
    ws = _workspaces[key]
    flags = np.zeros(ev.B, np.uint32)
    check(lib.evrep_window_flags(ws.data_ptr(), ev.B, flags.ctypes.data, stream))
    return flags


WF_NAMES = {_lib.WF_OUT_OF_RANGE: "an event outside the sensor was dropped", _lib.WF_UNSORTED: "timestamps decrease inside the window",
            _lib.WF_T_RANGE: "|t - t_first| >= 2^30 us: the event was dropped", _lib.WF_BAD_POLARITY: "a polarity outside {-1, 0, 1} was clamped"}


def raise_on_flags(ev, ignore=0):
    """Reads the per-window status words of the last op on the current stream (synchronises) and raises ValueError when any
    window carries an EVREP_WF_* bit not in `ignore` - the batched counterpart of the exceptions the reference's per-window
    code raises (the kernels themselves only flag: they cannot raise)."""
    flags = window_flags(ev)
    bad = flags & ~np.uint32(ignore)
    if bad.any():
        w = int(np.nonzero(bad)[0][0])
        what = "; ".join(v for k, v in WF_NAMES.items() if int(bad[w]) & k) or hex(int(bad[w]))
        raise ValueError(f"window {w} (of {int((bad != 0).sum())} flagged): {what}")
    return flags


def _codes(windows, functions, aggregations):
    C = len(windows)
    if not (len(functions) == len(aggregations) == C):
        raise ValueError("windows, functions and aggregations must have the same length")
    win = np.array([int(w) if -128 <= int(w) <= 127 else 127 for w in windows], np.int8)
    func = np.array([FUNCS.get(f, -1) if isinstance(f, str) else int(f) for f in functions], np.int8)
    agg = np.array([AGGS.get(a, -1) if isinstance(a, str) else int(a) for a in aggregations], np.int8)
    return win, func, agg, C


# tuples whose background compilation mixed_density(specialize="auto") has already asked for (one request per tuple and process)
_AUTO_SPECIALIZE_SEEN = set()
AUTO_SPECIALIZE_MIN_EVENTS = 4_000_000  # batches below this are launch bound: the interpreted kernel costs them nothing


def mixed_density(ev, H, W, windows, functions, aggregations, stacking="SBN", out=None, specialize="auto"):
    """MixedDensityEventStack.stack for every window -> (B, H, W, C) float32.
    (representations/representation_search/mixed_density_event_stack.py:25-46)
    specialize: "auto" - a batch of 4 M events or more asks for kernels compiled for this tuple on a background thread (the call
    itself never waits; later calls with the tuple run them, see specialize_mixed_density); True - compile now, then run;
    False - leave the choice of kernels to what has been specialised so far."""
    win, func, agg, C = _codes(windows, functions, aggregations)
    if stacking not in STACKING:
        # create_windows builds only window 0 for an unknown stacking type; every other index raises -> zero channel
        win = np.where((win == 0) | (win == -1), 0, 127).astype(np.int8)
        st = STACKING["SBN"]
    else:
        st = STACKING[stacking]
    if specialize is True or (specialize == "auto" and ev.total >= AUTO_SPECIALIZE_MIN_EVENTS):
        key = (st, win.tobytes(), func.tobytes(), agg.tobytes())
        if specialize is True or key not in _AUTO_SPECIALIZE_SEEN:
            if len(_AUTO_SPECIALIZE_SEEN) > 65536:
                _AUTO_SPECIALIZE_SEEN.clear()
            _AUTO_SPECIALIZE_SEEN.add(key)
            n_max = int(np.diff(ev.offsets).max()) if ev.B else 1
            fn = lib.evrep_mixed_density_specialize if specialize is True else lib.evrep_mixed_density_specialize_async
            with torch.cuda.device(ev.device):
                rc = fn(win.ctypes.data, func.ctypes.data, agg.ctypes.data, C, st, max(n_max, 1 << 20))
            if rc not in (_lib.OK, _lib.EUNSUPPORTED, _lib.EINVAL):  # outside the envelope / bad spec: the interpreted kernel serves (or reports) it
                check(rc)
    head, ws, stream = _prep(ev, _lib.OP_MIXED_DENSITY, H, W, C)
    out = _out(ev, (ev.B, H, W, C), out)
    check(lib.evrep_mixed_density_batched(*head, win.ctypes.data, func.ctypes.data, agg.ctypes.data, C, st, out.data_ptr(),
                                          ws.data_ptr(), ws.numel(), stream))
    return out


def specialize_mixed_density(windows, functions, aggregations, stacking="SBN", max_events_per_window=1 << 20, device=None, wait=True):
    """Compile kernels for ONE (windows, functions, aggregations) tuple at run time (NVRTC, a few seconds, once per process) and
    load them on `device` (default: the current one); `mixed_density` calls with that tuple then run 4 - 6 x faster than the
    interpreted kernel every other non-ERGO tuple takes (C-ABI: evrep_mixed_density_specialize).  For running a searched
    representation over a dataset.  Returns True, or False when the tuple is outside the
    specialised envelope (accumulators beyond a tile's shared memory; list entries the reference would swallow
    into zero channels are fine) - such tuples keep running on the interpreted kernel.  wait=False compiles on a background
    thread and returns at once; `mixed_density` switches kernels when the program is ready (mixed_density_is_specialized)."""
    win, func, agg, C = _codes(windows, functions, aggregations)
    if stacking not in STACKING:
        return False
    st = STACKING[stacking]
    if not wait:
        rc = lib.evrep_mixed_density_specialize_async(win.ctypes.data, func.ctypes.data, agg.ctypes.data, C, st, int(max_events_per_window))
        if rc == _lib.EUNSUPPORTED:
            return False
        check(rc)
        return True
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        rc = lib.evrep_mixed_density_specialize(win.ctypes.data, func.ctypes.data, agg.ctypes.data, C, st, int(max_events_per_window))
    if rc == _lib.EUNSUPPORTED:
        return False
    check(rc)
    return True


def mixed_density_is_specialized(windows, functions, aggregations, stacking="SBN", max_events_per_window=1):
    win, func, agg, C = _codes(windows, functions, aggregations)
    if stacking not in STACKING:
        return False
    return bool(lib.evrep_mixed_density_is_specialized(win.ctypes.data, func.ctypes.data, agg.ctypes.data, C, STACKING[stacking], int(max_events_per_window)))


def ergo12(ev, H, W, version=2, out=None):
    """ERGO-12 (get_optimized_representation, representations/optimized_representation.py:86-134)
    for every window -> (B, H, W, 12) float32."""
    head, ws, stream = _prep(ev, _lib.OP_MIXED_DENSITY, H, W, 12)
    out = _out(ev, (ev.B, H, W, 12), out)
    check(lib.evrep_ergo12_batched(*head, int(version), out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out


def event_stack(ev, H, W, stack_size=12, out=None):
    """EventStack pre_stack + post_stack (representations/event_stack.py:15-63) -> (B, H, W, stack_size) in {-1,0,1}."""
    head, ws, stream = _prep(ev, _lib.OP_EVENT_STACK, H, W, stack_size)
    out = _out(ev, (ev.B, H, W, stack_size), out)
    check(lib.evrep_event_stack_batched(*head, int(stack_size), out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out


def time_surface(ev, H, W, n_surfaces=6, tau=50000.0, indices=None, out=None):
    """ToTimesurface (representations/time_surface.py:25-74) -> (B, S, 2, H, W) float32.
    indices: None for the gen1_transforms.py:78-80 rule, else an int64 array (B, S) of snapshot event indices."""
    S = int(n_surfaces)
    idx_ptr = None
    if indices is not None:
        indices = np.ascontiguousarray(np.asarray(indices, np.int64).reshape(ev.B, S))
        idx_ptr = indices.ctypes.data
    head, ws, stream = _prep(ev, _lib.OP_TIME_SURFACE, H, W, 2 * S)
    out = _out(ev, (ev.B, S, 2, H, W), out)
    check(lib.evrep_time_surface_batched(*head, idx_ptr, S, float(tau), out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out


def tore(ev, H, W, k=6, out=None):
    """events2ToreFeature (representations/tore.py:6-83) with sampleTimes = last timestamp of each window and
    0-based pixel coordinates -> (B, H, W, 2k) float32."""
    head, ws, stream = _prep(ev, _lib.OP_TORE, H, W, 2 * k)
    out = _out(ev, (ev.B, H, W, 2 * k), out)
    check(lib.evrep_tore_batched(*head, int(k), out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out


def order_ops_fused(ev, H, W, tau=50000.0, out=None):
    """BASELINE configs[2] in one call: (EventStack(12) (B, H, W, 12), TimeSurface(6 snapshots) (B, 6, 2, H, W), TORE(k=6)
    (B, H, W, 12)) from a single bucketing pass; each equals the separate call bit for bit.  The fused record carries a 20-bit
    stream index and a 10-bit pixel: a batch with a window of 2^20 events or more, or a sensor too large for 1024-pixel tiles,
    is served by the three separate calls instead (same outputs, two more bucketing passes)."""
    es, ts, tr = out if out is not None else (None, None, None)
    n_max = int(np.diff(ev.offsets).max()) if ev.B else 0
    if n_max >= (1 << 20) or H * W > 1024 * _lib.MAX_TILES:
        return event_stack(ev, H, W, 12, out=es), time_surface(ev, H, W, 6, tau, out=ts), tore(ev, H, W, 6, out=tr)
    head, ws, stream = _prep(ev, _lib.OP_TORE, H, W, 12)
    es = _out(ev, (ev.B, H, W, 12), es)
    ts = _out(ev, (ev.B, 6, 2, H, W), ts)
    tr = _out(ev, (ev.B, H, W, 12), tr)
    check(lib.evrep_order_ops_fused_batched(*head, float(tau), es.data_ptr(), ts.data_ptr(), tr.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return es, ts, tr


def voxel_grid(ev, H, W, n_bins, flavour="tonic", normalize=True, t0_us=None, t1_us=None, out=None, divider=1):
    """flavour 'tonic' -> (B, n_bins, H, W); 'evlicious' -> (B, n_bins, H, W); 'gwd' -> (B, H, W, n_bins).
    divider > 1 (ev-licious only): ev.x / ev.y hold sub-pixel integers, the event sits at (x / divider, y / divider) and is
    spread over its four neighbours (utils.py:93-103); H x W is the grid."""
    fl = {"tonic": _lib.VOXEL_TONIC, "evlicious": _lib.VOXEL_EVLICIOUS, "gwd": _lib.VOXEL_GWD}[flavour]
    if divider != 1 and fl != _lib.VOXEL_EVLICIOUS:
        raise ValueError("divider applies to the ev-licious flavour only")
    t01 = None
    if (t0_us is not None or t1_us is not None) and fl == _lib.VOXEL_EVLICIOUS:
        if t0_us is None or t1_us is None:
            raise ValueError("give both t0_us and t1_us or neither (the batched call cannot read single timestamps back)")
        t01 = np.array([int(t0_us), int(t1_us)], np.int64)
    head, ws, stream = _prep(ev, _lib.OP_VOXEL, H, W, n_bins)
    shape = (ev.B, H, W, n_bins) if fl == _lib.VOXEL_GWD else (ev.B, n_bins, H, W)
    out = _out(ev, shape, out)
    if divider != 1:
        check(lib.evrep_voxel_subpixel_batched(*head, int(divider), int(n_bins), int(bool(normalize)), None if t01 is None else t01.ctypes.data,
                                               out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return out
    check(lib.evrep_voxel_batched(*head, fl, int(n_bins), int(bool(normalize)), None if t01 is None else t01.ctypes.data,
                                  out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out


def histogram(ev, H, W, out=None):
    """tonic ToImage counts (gen1_transforms.py:44-49) -> (B, 2, H, W) float32, plane 1 = p > 0."""
    head, ws, stream = _prep(ev, _lib.OP_HISTOGRAM, H, W, 2)
    out = _out(ev, (ev.B, 2, H, W), out)
    check(lib.evrep_histogram_batched(*head, out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out


def gwd_pack(point_sets, device="cuda"):
    """List of (n_i, d) arrays / tensors -> (float64 CUDA tensor (sum n_i, d), host int64 offsets (len + 1)): the layout
    evrep_gwd_kernel_l1 consumes.  Pack once when the same point sets are evaluated repeatedly (or build the layout directly:
    otmi_prepare already returns views of it)."""
    dev = torch.device(device)
    ts = [torch.as_tensor(a).to(device=dev, dtype=torch.float64).reshape(len(a), -1) for a in point_sets]
    d = ts[0].shape[1] if ts else 1
    if any(t.shape[1] != d for t in ts):
        raise ValueError("all point sets must share the feature width")
    offs = np.zeros(len(ts) + 1, np.int64)
    offs[1:] = np.cumsum([t.shape[0] for t in ts])
    return (torch.cat(ts, 0).contiguous() if ts else torch.zeros((0, d), dtype=torch.float64, device=dev)), offs


def gwd_kernel_l1(Xs_list, Xt_list, h=0.7, device="cuda", s_offsets=None, t_offsets=None):
    """GWD-A cost of every (Xs, Xt) pair (compute_otmi.py:50-93 in closed form) -> float64 CUDA tensor (n_pairs,).
    Xs_list / Xt_list: lists of (n_i, ds) / (m_i, dt) arrays or tensors (any float dtype), packed here on every call - or,
    with s_offsets / t_offsets (host int64, n_pairs + 1), the already packed float64 CUDA tensors of gwd_pack."""
    dev = torch.device(device)
    if s_offsets is not None:
        Xs, Xt = Xs_list, Xt_list
        so, to = np.ascontiguousarray(s_offsets, np.int64), np.ascontiguousarray(t_offsets, np.int64)
        if not (torch.is_tensor(Xs) and torch.is_tensor(Xt) and Xs.is_cuda and Xt.is_cuda and Xs.dtype == torch.float64 and Xt.dtype == torch.float64
                and Xs.is_contiguous() and Xt.is_contiguous() and Xs.dim() == 2 and Xt.dim() == 2):
            raise ValueError("packed inputs must be contiguous 2-D float64 CUDA tensors (gwd_pack)")
        if len(so) != len(to) or Xs.shape[0] < so[-1] or Xt.shape[0] < to[-1]:
            raise ValueError("offsets do not match the packed tensors")
        n_pairs, ds, dt = len(so) - 1, int(Xs.shape[1]), int(Xt.shape[1])
        dev = Xs.device
    else:
        if len(Xs_list) != len(Xt_list):
            raise ValueError("need as many Xs as Xt")
        n_pairs = len(Xs_list)
        if n_pairs:
            (Xs, so), (Xt, to) = gwd_pack(Xs_list, dev), gwd_pack(Xt_list, dev)
            ds, dt = int(Xs.shape[1]), int(Xt.shape[1])
    if n_pairs == 0:
        return torch.zeros(0, dtype=torch.float64, device=dev)
    nbytes = lib.evrep_gwd_workspace_bytes(so.ctypes.data, to.ctypes.data, n_pairs)
    if nbytes == 0:
        raise ValueError("invalid pair sizes")
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = _workspace(dev, stream, nbytes)
    out = torch.empty(n_pairs, dtype=torch.float64, device=dev)
    check(lib.evrep_gwd_kernel_l1(Xs.data_ptr(), so.ctypes.data, ds, Xt.data_ptr(), to.ctypes.data, dt, n_pairs, float(h),
                                  out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return out


def gemm_nt_3xtf32(A, B, alpha=1.0, row_vec=None, col_vec=None, out=None, packed=True):
    """alpha * A @ B.T + row_vec[:, None] + col_vec[None, :] on the tcgen05 tensor cores with fp32-class accuracy
    (3 x TF32 split).  A (M, K), B (N, K): float32 CUDA tensors.  This is the contraction of GWD-B's tensor product."""
    if not (A.is_cuda and B.is_cuda):
        raise ValueError("gemm_nt_3xtf32 needs CUDA tensors (there is no CPU path)")
    A = A.contiguous().float()
    B = B.contiguous().float()
    M, K = A.shape
    N, K2 = B.shape
    if K != K2:
        raise ValueError("A and B must share the contraction length")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    rv = row_vec.contiguous().float() if row_vec is not None else None
    cv = col_vec.contiguous().float() if col_vec is not None else None
    stream = torch.cuda.current_stream(A.device).cuda_stream
    ws = _workspace(A.device, stream, lib.evrep_gemm_workspace_bytes(M, N, K)) if packed else None  # packed: TMA-fed kernel
    check(lib.evrep_gemm_nt_3xtf32(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, float(alpha),
                                   rv.data_ptr() if rv is not None else None, cv.data_ptr() if cv is not None else None,
                                   ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0, stream))
    return out


LMO = {"auction": 0, "host": 1}


def gw_kl(Xs, Xt, h=0.7, max_iter=10000, tol_rel=1e-9, tol_abs=1e-9, return_plan=False, device="cuda", lmo="auction", stats=None):
    """GWD-B (gromov_wasserstein.py:39-69): Gaussian kernels of Xs (n, ds) and Xt (m, dt), then conditional-gradient
    Gromov-Wasserstein with the KL loss.  -> (gw_dist, iterations[, plan (n, m) float32 CUDA tensor]).
    lmo: "auction" (assignment solved on the GPU, n == m) or "host" (exact solver on the CPU; always used for n != m, where
    the vertex is a transportation plan); `stats`, if a dict, receives the auction's round / bid counts and the number of
    steps solved on the host."""
    import ctypes
    dev = torch.device(device)
    Xs = torch.as_tensor(Xs).to(device=dev, dtype=torch.float64).contiguous()
    Xt = torch.as_tensor(Xt).to(device=dev, dtype=torch.float64).contiguous()
    Xs = Xs.reshape(len(Xs), -1)
    Xt = Xt.reshape(len(Xt), -1)
    n, ds = Xs.shape
    m, dt = Xt.shape
    nbytes = lib.evrep_gw_kl_workspace_bytes(n, m)
    if nbytes == 0:
        raise ValueError("invalid sizes")
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = _workspace(dev, stream, nbytes)
    plan = torch.empty((n, m), dtype=torch.float32, device=dev) if return_plan else None
    dist, iters = ctypes.c_double(0.0), ctypes.c_int(0)
    lst = (ctypes.c_int * 3)()
    check(lib.evrep_gw_kl(Xs.data_ptr(), n, ds, Xt.data_ptr(), m, dt, float(h), int(max_iter), float(tol_rel), float(tol_abs), LMO[lmo],
                          ctypes.byref(dist), plan.data_ptr() if plan is not None else None, ctypes.byref(iters), lst, ws.data_ptr(),
                          ws.numel(), stream))
    if stats is not None:
        stats.update(auction_rounds=lst[0], auction_bids=lst[1], host_fallbacks=lst[2])
    return (dist.value, iters.value, plan) if return_plan else (dist.value, iters.value)


IMG_MODES = {"letterbox": 0, "squash": 1}
INTERP = {"auto": 0, "linear": 1, "area": 2, "linear_torch": 3}


def detector_input(rep, img_size=640, mode="letterbox", interp="auto", scale_in=255.0, scale_out=1.0 / 255.0, pad_value=114.0,
                   reverse_channels=True, out=None):
    """The reference's per-sample image pipeline for a whole batch, fused on the GPU: (B, H, W, C) representation ->
    x255 -> cv2.resize per channel -> letterbox(114) (or squash) -> CHW with reversed channel order -> /255, i.e. what
    Gen1H5.__getitem__ (gen1_2yolo.py:230-265, 321-341, 397) plus Trainer.prepro_data (engine.py:629-635) hand to the
    detector.  Returns a float32 CUDA tensor (B, C, img_size, img_size)."""
    if not rep.is_cuda:
        raise ValueError("detector_input needs a CUDA tensor (there is no CPU path)")
    rep = rep.contiguous().float()
    B, H, W, C = rep.shape
    if out is None:
        out = torch.empty((B, C, img_size, img_size), dtype=torch.float32, device=rep.device)
    stream = torch.cuda.current_stream(rep.device).cuda_stream
    check(lib.evrep_image_pipeline_batched(rep.data_ptr(), B, H, W, C, int(img_size), IMG_MODES[mode], INTERP[interp], float(scale_in),
                                           float(scale_out), float(pad_value), 1 if reverse_channels else 0, out.data_ptr(), stream))
    return out


def augment_affine(img, M, flip_ud=None, flip_lr=None, out_size=None, border=(114.0, 114.0, 114.0, 0.0), reverse_channels=True, scale_out=1.0 / 255.0,
                   out=None):
    """The training-time augmentation of the detector datasets on a batch (gen1_2yolo.py:365-391): random_affine's
    cv2.warpAffine (data_augment.py:110-123, arithmetic of cv::warpAffine INTER_LINEAR reproduced tap for tap) and the flips of
    general_augment (:210-228), then CHW channel reversal and / 255.  img: (B, C, h, w) float32 CUDA planes - detector_input(...,
    scale_out=1.0, reverse_channels=False) gives exactly the letterboxed image the reference warps; M: (B, 2, 3) or (B, 3, 3)
    FORWARD matrices as get_transform_matrix returns them (the caller draws the random numbers, as it does for the labels);
    flip_ud / flip_lr: per-window booleans.  `border` is what cv2 makes of borderValue=(114, 114, 114): channel k is padded
    with border[k & 3], i.e. every fourth channel with 0.  -> (B, C, out_size, out_size)."""
    if not img.is_cuda:
        raise ValueError("augment_affine needs a CUDA tensor (there is no CPU path)")
    img = img.contiguous().float()
    B, C, h, w = img.shape
    Mh = np.ascontiguousarray(np.asarray(M, np.float64).reshape(B, -1, 3)[:, :2, :].reshape(B, 6))
    flips = np.zeros(B, np.int32)
    if flip_ud is not None:
        flips |= np.asarray(flip_ud, bool).astype(np.int32)
    if flip_lr is not None:
        flips |= np.asarray(flip_lr, bool).astype(np.int32) << 1
    oh, ow = (h, w) if out_size is None else ((out_size, out_size) if np.isscalar(out_size) else tuple(out_size))
    if out is None:
        out = torch.empty((B, C, oh, ow), dtype=torch.float32, device=img.device)
    b4 = np.asarray(border, np.float32)
    if b4.shape != (4,):
        raise ValueError("border must have four entries (cv2's scalar)")
    _require_current(img.device)
    stream = torch.cuda.current_stream(img.device).cuda_stream
    check(lib.evrep_warp_affine_batched(img.data_ptr(), B, C, h, w, Mh.ctypes.data, flips.ctypes.data, int(oh), int(ow), b4.ctypes.data,
                                        1 if reverse_channels else 0, float(scale_out), out.data_ptr(), stream))
    return out


def assignment_auction(cost, eps_rel=1e-9):
    """Min-cost assignment of a square float32 CUDA matrix on the GPU (the LMO of gw_kl) -> (sigma int32 CUDA tensor,
    stats dict).  Optimal up to n * eps_rel * (cost range)."""
    if not cost.is_cuda:
        raise ValueError("assignment_auction needs a CUDA tensor (there is no CPU path)")
    cost = cost.contiguous().float()
    n = cost.shape[0]
    if cost.shape != (n, n):
        raise ValueError("cost must be square")
    sigma = torch.empty(n, dtype=torch.int32, device=cost.device)
    st = torch.zeros(3, dtype=torch.int32, device=cost.device)
    check(lib.evrep_assignment_auction(cost.data_ptr(), n, float(eps_rel), sigma.data_ptr(), st.data_ptr(),
                                       torch.cuda.current_stream(cost.device).cuda_stream))
    r, b, status = (int(v) for v in st.cpu())
    return sigma, {"rounds": r, "bids": b, "status": status}


class GraphedCall:
    """A fixed sequence of batched calls captured into a CUDA graph: `fn` is run (twice eagerly, then once under capture) on
    a private stream and `replay()` re-issues all of its kernels with ONE launch.  Valid as long as the device buffers `fn`
    touches keep their addresses and the window offsets stay the same (fixed-size windows, e.g. Gen1's num_events = 50 000,
    ev-YOLOv6/tools/train.py:43); event values may change between replays.  Batches of up to 319 windows only: their window
    tables travel as kernel arguments, larger ones need a host-to-device copy that a capture cannot hold.  This is what
    makes small batches (BASELINE configs[1]: 6.4 M events per step) stop being launch bound."""

    def __init__(self, fn, warmup=2):
        self._stream = torch.cuda.Stream()
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            for _ in range(max(1, warmup)):  # workspaces and shared-memory attributes are set up outside the capture
                fn()
        torch.cuda.current_stream().wait_stream(self._stream)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self._stream):
            self.result = fn()

    def replay(self):
        self.graph.replay()
        return self.result


def transport_plan_host(cost):
    """The rectangular LMO of gw_kl on its own (HOST, no GPU involved): an optimal vertex of
    min <cost, G> s.t. G 1 = 1/n, G^T 1 = 1/m, G >= 0 for a float32 (n, m) numpy cost -> dense float64 (n, m) plan."""
    import ctypes
    import numpy as np
    cost = np.ascontiguousarray(cost, dtype=np.float32)
    n, m = cost.shape
    cap = 2 * (n + m) + 16
    rp = np.zeros(n + 1, np.int32)
    col = np.zeros(cap, np.int32)
    w = np.zeros(cap, np.float64)
    nnz = ctypes.c_int(0)
    check(lib.evrep_transport_plan_host(cost.ctypes.data, n, m, cap, rp.ctypes.data, col.ctypes.data, w.ctypes.data, ctypes.byref(nnz)))
    G = np.zeros((n, m))
    for i in range(n):
        G[i, col[rp[i]:rp[i + 1]]] = w[rp[i]:rp[i + 1]]
    return G


FILTERS = {"refractory": 0, "contrast": 1, "resize": 2, "background": 3}
_FILTER_STATE_DTYPE = {"refractory": torch.float64, "contrast": torch.int32, "resize": torch.float32, "background": torch.float64}


def otmi_prepare(events, rep, height, width, rep_size):
    """The data preparation of otmi() (compute_otmi.py:96-203) in eight small kernels (evrep_otmi_prepare): `events` (N, 4)
    [x, y, t, p], an integer array / tensor (the reference's own call: exact differences, float32 quotients like torch's int / int)
    or float32 / float64 (the array's precision), `rep`
    (rep_size, rep_size, C).  -> (pairs, info): pairs = three (Xs (n_i, 4), Xt (m_i, C + 2)) float64 CUDA tensor views, ready
    for gwd_kernel_l1; info = {"dropped": quadrant, "events_per_quadrant": [...]}.  Raises ValueError when one of the quadrants
    2..4 is empty (the reference's min() of an empty column)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    ev = events if torch.is_tensor(events) else torch.as_tensor(np.ascontiguousarray(events))
    if ev.dtype not in (torch.int32, torch.float32, torch.float64):
        if ev.dtype.is_floating_point:
            ev = ev.float()  # half precisions: torch would divide in them; not a case the reference produces
        else:
            if ev.numel() and (int(ev.min()) < -2**31 or int(ev.max()) >= 2**31):
                raise ValueError("integer events outside the int32 range")
            ev = ev.to(torch.int32)  # torch divides any integer tensor in float32
    ev_type = {torch.int32: 0, torch.float32: 1, torch.float64: 2}[ev.dtype]
    ev = ev.to(dev).contiguous()
    if ev.dim() != 2 or ev.shape[1] != 4:
        raise ValueError("events must be (N, 4): x, y, t, p")
    rp = (rep if torch.is_tensor(rep) else torch.as_tensor(np.ascontiguousarray(rep))).to(dev).double().contiguous()
    if rp.dim() != 3 or rp.shape[0] < rep_size or rp.shape[1] < rep_size:
        raise ValueError("rep must be (rep_size, rep_size, C)")
    if rp.shape[0] != rep_size or rp.shape[1] != rep_size:
        raise ValueError(f"rep is {tuple(rp.shape[:2])}, rep_size says {rep_size}")
    N, C = int(ev.shape[0]), int(rp.shape[2])
    cap_s, cap_t = max(N, 1), (rep_size // 2 + 2) ** 2
    Xs = torch.empty((3, cap_s, 4), dtype=torch.float64, device=dev)
    Xt = torch.empty((3, cap_t, C + 2), dtype=torch.float64, device=dev)
    info = np.zeros(12, np.int64)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = _workspace(dev, stream, lib.evrep_otmi_workspace_bytes(N, rep_size))
    check(lib.evrep_otmi_prepare(ev.data_ptr(), ev_type, N, rp.data_ptr(), rep_size, C, int(height), int(width), Xs.data_ptr(), cap_s,
                                 Xt.data_ptr(), cap_t, info.ctypes.data, ws.data_ptr(), ws.numel(), stream))
    if info[7]:
        raise ValueError("min() arg is an empty sequence")  # compute_otmi.py:140-147 on an empty quadrant
    pairs = [(Xs[s, :int(info[s])], Xt[s, :int(info[3 + s])]) for s in range(3)]
    return pairs, {"dropped": int(info[6]), "events_per_quadrant": [int(v) for v in info[8:12]]}


def filter_state(kind, B, H, W, device="cuda"):
    """Fresh state for `filter_events`: -inf last timestamps (refractory, background), zero activity (contrast) / change map (resize)."""
    fill = float("-inf") if kind in ("refractory", "background") else 0
    return torch.full((B, H, W), fill, dtype=_FILTER_STATE_DTYPE[kind], device=device)


def filter_events(ev, H, W, kind, param=0.0, state=None, fx=1, fy=1):
    """ev-licious' per-pixel stateful filters over a batch of streams (utils.py:143-158, 184-200) -> (mask, state):
    mask is a uint8 CUDA tensor with one entry per event (1 = the event passes), state the (B, H, W) per-pixel state,
    updated in place, to pass to the next call of the same streams.  kind: "refractory" (param = period), "contrast"
    (param = factor), "resize" (H, W = the change-map size, fx, fy = cell size) or "background" (utils.py:169-178; param =
    depth_us, fx = radius; the polarities are not read)."""
    B = len(ev.offsets) - 1
    dev = ev.x.device
    if state is None:
        state = filter_state(kind, B, H, W, dev)
    if state.dtype != _FILTER_STATE_DTYPE[kind] or tuple(state.shape) != (B, H, W) or not state.is_cuda or not state.is_contiguous():
        raise ValueError(f"state must be a contiguous CUDA {_FILTER_STATE_DTYPE[kind]} tensor of shape {(B, H, W)}")
    total = int(ev.offsets[-1])
    mask = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    offs = np.ascontiguousarray(ev.offsets, np.int64)
    if kind == "background":
        radius = int(fx)
        need = lib.evrep_filter_background_workspace_bytes(B, total, H, W, radius, ev.t.element_size())
        if need == 0:
            raise ValueError(f"background-activity filter: unsupported radius {radius} (1..4) or sizes")
        ws = _workspace(dev, stream, need)
        check(lib.evrep_filter_background_batched(ev.x.data_ptr(), ev.y.data_ptr(), ev.t.data_ptr(), ev.t.element_size(), offs.ctypes.data, B, H, W,
                                                  float(param), radius, state.data_ptr(), mask.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return mask[:total], state
    ws = _workspace(dev, stream, lib.evrep_workspace_bytes(7, B, total, H, W, 1))
    check(lib.evrep_filter_batched(ev.x.data_ptr(), ev.y.data_ptr(), ev.t.data_ptr(), ev.t.element_size(), ev.p.data_ptr(), offs.ctypes.data,
                                   B, H, W, FILTERS[kind], float(param), int(fx), int(fy), state.data_ptr(), mask.data_ptr(), ws.data_ptr(),
                                   ws.numel(), stream))
    return mask[:total], state
