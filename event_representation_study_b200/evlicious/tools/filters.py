"""Mirror of ev-licious/src/evlicious/tools/filters.py (reference :7-129): the filter objects the ev-licious tools build
from command-line flags.  Same class names, constructor arguments, `insert(events) -> Events` behaviour and state
attributes (numpy arrays, created on the first call and updated by every call); the per-event work runs on the GPU through
the function mirrors in `tools/utils.py` (evrep_filter_batched / evrep_filter_background_batched) and, for HotPixel, a
count plane from the mixed-density kernel.  `Random` has no per-pixel work and stays a host-side index draw."""
import enum

import numpy as np

from . import utils as _u


class Filtering_Type(enum.IntEnum):
    BackgroundActivity = 1
    Random = 2
    ContrastThresholdIncrease = 3
    RefractoryPeriod = 4
    HotPixel = 5

    @classmethod
    def summary(cls):
        return "".join(f" {name}={int(member)} " for name, member in cls.__members__.items())


class _PixelStateFilter:
    """Shared shape of the three stateful filters: a per-pixel numpy array under the attribute name the reference uses,
    allocated from the first batch's sensor size, and one GPU pass per `insert` that yields the keep-mask."""
    _state_name = None   # attribute holding the (height, width) state
    _state_fill = 0
    _state_dtype = np.float64
    _mask_start = True   # the reference starts from an all-True (refractory, background) or all-False (contrast) mask

    def _state(self, events):
        if getattr(self, self._state_name) is None:
            setattr(self, self._state_name, np.full((events.height, events.width), self._state_fill, dtype=self._state_dtype))
        return getattr(self, self._state_name)

    def _decide(self, mask, state, events):
        raise NotImplementedError

    def insert(self, events):
        state = self._state(events)
        mask = np.full(len(events.x), self._mask_start, dtype=bool)
        return events[self._decide(mask, state, events)]


class BackgroundActivity(_PixelStateFilter):
    """filters.py:56-69; state `timestamps` (float64, -inf)"""
    _state_name, _state_fill = "timestamps", -np.inf

    def __init__(self, depth_us, radius):
        self.radius = radius
        self.depth_us = depth_us
        self.timestamps = None

    def _decide(self, mask, state, events):
        return _u._background_activity_filter(mask, state, events.x, events.y, events.t, self.depth_us, self.radius)


class RefractoryPeriod(_PixelStateFilter):
    """filters.py:99-111; state `timestamps` (float64, -inf)"""
    _state_name, _state_fill = "timestamps", -np.inf

    def __init__(self, depth_us):
        self.depth_us = depth_us
        self.timestamps = None

    def _decide(self, mask, state, events):
        return _u._refractory_period(mask, events.x, events.y, events.t, self.depth_us, state)


class ContrastThresholdIncrease(_PixelStateFilter):
    """filters.py:82-96; state `counter_map` (int32, 0)"""
    _state_name, _state_dtype, _mask_start = "counter_map", np.int32, False

    def __init__(self, contrast_threshold_multiplier):
        self.contrast_threshold_multiplier = contrast_threshold_multiplier
        self.counter_map = None

    def _decide(self, mask, state, events):
        return _u._contrast_threshold_control(state, mask, events.x, events.y, events.p, self.contrast_threshold_multiplier)


class Random:
    """filters.py:72-79: keeps len(events) // factor events drawn without replacement, in the drawn order"""

    def __init__(self, random_downsampling_factor):
        self.random_downsampling_factor = random_downsampling_factor

    def insert(self, events):
        n = len(events)
        return events[np.random.choice(n, n // self.random_downsampling_factor, replace=False)]


def _pixel_counts(events):
    """events per pixel, float64 (H, W): one ("count", "sum") plane from the GPU"""
    from ... import batched as eb
    from ..._single import one_window
    H, W, n = events.height, events.width, len(events)
    if n == 0:
        return np.zeros((H, W))
    ev = one_window(events.x, events.y, np.arange(n, dtype=np.int64), np.ones(n, np.int8), H, W)
    return eb.mixed_density(ev, H, W, [0], ["count"], ["sum"])[0, :, :, 0].double().cpu().numpy()


class HotPixel:
    """filters.py:23-53: the first batch calibrates a pixel mask - pixels whose count stays below `threshold` of the busiest
    pixel pass, provided the excluded ones are more than twice as busy as every pixel that passes; otherwise nothing is
    excluded.  Later batches are gated by that mask."""

    def __init__(self):
        self.hot_pixel_mask = None

    def calibrate(self, events, debug=False, threshold=0.6):
        if debug:
            raise NotImplementedError("the matplotlib debug view of the reference is not mirrored")
        count = _pixel_counts(events)
        passes = count / np.max(count) < threshold
        quietest_excluded, busiest_passing = np.min(count[~passes]), np.max(count[passes])
        if float(quietest_excluded) / busiest_passing > 2:
            return passes
        return np.ones((events.height, events.width)) > 0

    def insert(self, events):
        if self.hot_pixel_mask is None:
            self.hot_pixel_mask = self.calibrate(events)
        return events[self.hot_pixel_mask[events.y, events.x]]


def from_flags(flags):
    """filters.py:114-129: the filter named by flags.filter_type, its parameters asserted positive like the reference does"""
    def positive(*names):
        for name in names:
            assert getattr(flags, name) > 0
        return {name: getattr(flags, name) for name in names}

    kind = flags.filter_type
    if kind == int(Filtering_Type.BackgroundActivity):
        return BackgroundActivity(**positive("depth_us", "radius"))
    if kind == int(Filtering_Type.Random):
        return Random(**positive("random_downsampling_factor"))
    if kind == int(Filtering_Type.ContrastThresholdIncrease):
        return ContrastThresholdIncrease(**positive("contrast_threshold_multiplier"))
    if kind == int(Filtering_Type.RefractoryPeriod):
        return RefractoryPeriod(**positive("depth_us"))
    if kind == int(Filtering_Type.HotPixel):
        return HotPixel()
    raise ValueError("Filter unknown")
