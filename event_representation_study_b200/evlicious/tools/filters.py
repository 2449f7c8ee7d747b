"""Mirror of ev-licious/src/evlicious/tools/filters.py (reference :7-125): the filter objects the ev-licious tools build
from command-line flags.  Same class names, constructor arguments, `insert(events) -> Events` behaviour and state
attributes (numpy arrays, updated by every call); the per-event work runs on the GPU through the function mirrors in
`tools/utils.py` (evrep_filter_batched / evrep_filter_background_batched) and, for HotPixel, a count plane from the
mixed-density kernel.  `Random` has no per-pixel work and stays a host-side index draw, as in the reference."""
import enum

import numpy as np

from .utils import _background_activity_filter, _contrast_threshold_control, _refractory_period


class Filtering_Type(enum.IntEnum):
    BackgroundActivity = enum.auto()
    Random = enum.auto()
    ContrastThresholdIncrease = enum.auto()
    RefractoryPeriod = enum.auto()
    HotPixel = enum.auto()

    @classmethod
    def summary(cls):
        return "".join(f" {name}={int(member)} " for name, member in cls.__members__.items())


def _pixel_counts(events):
    """events per pixel, float64 (H, W): one ("count", "sum") plane from the GPU"""
    from ... import batched as eb
    from ..._single import one_window
    H, W = events.height, events.width
    n = len(events)
    if n == 0:
        return np.zeros((H, W))
    ev = one_window(events.x, events.y, np.arange(n, dtype=np.int64), np.ones(n, np.int8), H, W)
    return eb.mixed_density(ev, H, W, [0], ["count"], ["sum"])[0, :, :, 0].double().cpu().numpy()


class HotPixel:
    """filters.py:23-53: the first batch calibrates a pixel mask (pixels whose count is below `threshold` of the busiest
    pixel pass, provided the hot ones are at least twice as busy as every other pixel); later batches are gated by it."""

    def __init__(self):
        self.hot_pixel_mask = None

    def calibrate(self, events, debug=False, threshold=0.6):
        if debug:
            raise NotImplementedError("the matplotlib debug view of the reference is not mirrored")
        count = _pixel_counts(events)
        mask = count / np.max(count) < threshold
        busiest_kept_out = np.min(count[~mask])
        busiest_let_through = np.max(count[mask])
        if float(busiest_kept_out) / busiest_let_through > 2:
            return mask
        return np.ones(shape=(events.height, events.width)) > 0

    def insert(self, events):
        if self.hot_pixel_mask is None:
            self.hot_pixel_mask = self.calibrate(events)
        return events[self.hot_pixel_mask[events.y, events.x]]


class BackgroundActivity:
    """filters.py:56-69"""

    def __init__(self, depth_us, radius):
        self.radius = radius
        self.depth_us = depth_us
        self.timestamps = None

    def insert(self, events):
        if self.timestamps is None:
            self.timestamps = np.full(shape=(events.height, events.width), fill_value=-np.inf)
        mask = np.ones_like(events.x) > 0
        return events[_background_activity_filter(mask, self.timestamps, events.x, events.y, events.t, self.depth_us, self.radius)]


class Random:
    """filters.py:72-79: keeps len(events) // factor events drawn without replacement (in the drawn order)"""

    def __init__(self, random_downsampling_factor):
        self.random_downsampling_factor = random_downsampling_factor

    def insert(self, events):
        return events[np.random.choice(len(events), len(events) // self.random_downsampling_factor, replace=False)]


class ContrastThresholdIncrease:
    """filters.py:82-96"""

    def __init__(self, contrast_threshold_multiplier):
        self.contrast_threshold_multiplier = contrast_threshold_multiplier
        self.counter_map = None

    def insert(self, events):
        if self.counter_map is None:
            self.counter_map = np.zeros(shape=(events.height, events.width), dtype="int32")
        mask = np.ones_like(events.x) < 0
        return events[_contrast_threshold_control(self.counter_map, mask, events.x, events.y, events.p, self.contrast_threshold_multiplier)]


class RefractoryPeriod:
    """filters.py:99-111"""

    def __init__(self, depth_us):
        self.depth_us = depth_us
        self.timestamps = None

    def insert(self, events):
        if self.timestamps is None:
            self.timestamps = np.full(shape=(events.height, events.width), fill_value=-np.inf)
        mask = np.ones_like(events.x) > 0
        return events[_refractory_period(mask, events.x, events.y, events.t, self.depth_us, self.timestamps)]


def from_flags(flags):
    """filters.py:114-129"""
    kind = flags.filter_type
    if kind == int(Filtering_Type.BackgroundActivity):
        assert flags.depth_us > 0
        assert flags.radius > 0
        return BackgroundActivity(depth_us=flags.depth_us, radius=flags.radius)
    if kind == int(Filtering_Type.Random):
        assert flags.random_downsampling_factor > 0
        return Random(random_downsampling_factor=flags.random_downsampling_factor)
    if kind == int(Filtering_Type.ContrastThresholdIncrease):
        assert flags.contrast_threshold_multiplier > 0
        return ContrastThresholdIncrease(contrast_threshold_multiplier=flags.contrast_threshold_multiplier)
    if kind == int(Filtering_Type.RefractoryPeriod):
        assert flags.depth_us > 0
        return RefractoryPeriod(depth_us=flags.depth_us)
    if kind == int(Filtering_Type.HotPixel):
        return HotPixel()
    raise ValueError("Filter unknown")
