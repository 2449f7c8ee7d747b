"""Mirror of ev-licious/src/evlicious/io/utils/events.py::Events (reference :7-67): the SoA input type of the operator API.
Same constructor, field names (`_x`, `_y` raw sub-pixel integers; `x`, `y` divided by `divider`), dtype / range assertions,
the in-place rewrite of polarity 0 to -1, `len`, masking / slicing (copies) and the dict / array exports.  Rendering, HDF5
output and the array / dict constructors of the reference are file-format and visualiser code (out of scope)."""
import numpy as np

TYPES = dict(_x=np.uint16, _y=np.uint16, t=np.int64, p=np.int8, x=np.uint16, y=np.uint16)
_STORED = ("_x", "_y", "t", "p")


class Events:
    def __init__(self, x, y, t, p, width, height, divider=1):
        self._x, self._y, self.t, self.p = x, y, t, p
        self.width, self.height, self.divider = width, height, divider
        self._check()

    def _check(self):
        for name in _STORED:
            have, want = getattr(self, name).dtype, TYPES[name]
            assert have == want, f"Field {name} does not have type {want}, but {have}."
        shapes = {getattr(self, name).shape for name in ("x", "y", "p", "t")}
        assert len(shapes) == 1
        assert self.x.ndim == 1
        if self._x.size == 0:
            return
        assert np.max(self.p) <= 1
        self.p[self.p == 0] = -1  # the operator API works with -1 / +1
        for coord, limit in ((self.x, self.width), (self.y, self.height)):
            assert np.max(coord) <= limit - 1, np.max(coord)
            assert np.min(coord) >= 0

    def _scaled(self, raw):
        return raw.astype("float32") / self.divider if self.divider > 1 else raw

    @property
    def x(self):
        return self._scaled(self._x)

    @property
    def y(self):
        return self._scaled(self._y)

    def __len__(self):
        return len(self.x)

    def __getitem__(self, item):
        """a COPY of the selected events (boolean mask, index array or slice), same sensor geometry"""
        picked = {name.lstrip("_"): getattr(self, name)[item].copy() for name in _STORED}
        return Events(width=self.width, height=self.height, divider=self.divider, **picked)

    def to_dict(self, format="xytp"):
        return {k: getattr(self, k) for k in format}

    def to_array(self, format="xytp"):
        return np.stack([getattr(self, k) for k in format], axis=-1)
