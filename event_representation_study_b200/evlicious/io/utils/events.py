"""Mirror of ev-licious/src/evlicious/io/utils/events.py::Events (reference :7-58): the SoA input type of the operator API."""
import numpy as np

TYPES = dict(_x=np.uint16, _y=np.uint16, t=np.int64, p=np.int8, x=np.uint16, y=np.uint16)


class Events:
    def __init__(self, x, y, t, p, width, height, divider=1):
        self._x = x
        self._y = y
        self.t = t
        self.p = p
        self.width = width
        self.height = height
        self.divider = divider
        for k, ty in TYPES.items():
            if k not in ["x", "y"]:
                assert getattr(self, k).dtype == ty, f"Field {k} does not have type {ty}, but {getattr(self, k).dtype}."
        assert self.x.shape == self.y.shape == self.p.shape == self.t.shape
        assert self.x.ndim == 1
        if self._x.size > 0:
            assert np.max(self.p) <= 1
            self.p[self.p == 0] = -1
            assert np.max(self.x) <= self.width - 1, np.max(self.x)
            assert np.max(self.y) <= self.height - 1, np.max(self.y)
            assert np.min(self.x) >= 0
            assert np.min(self.y) >= 0

    @property
    def x(self):
        if self.divider > 1:
            return self._x.astype("float32") / self.divider
        return self._x

    @property
    def y(self):
        if self.divider > 1:
            return self._y.astype("float32") / self.divider
        return self._y

    def __len__(self):
        return len(self.x)

    def __getitem__(self, item):
        """events.py:60-67: a copy of the selected events (mask, index array or slice)"""
        return Events(x=self._x[item].copy(), y=self._y[item].copy(), t=self.t[item].copy(), p=self.p[item].copy(), width=self.width,
                      height=self.height, divider=self.divider)

    def to_dict(self, format="xytp"):
        return {k: getattr(self, k) for k in format}

    def to_array(self, format="xytp"):
        return np.stack([getattr(self, k) for k in format], axis=-1)
