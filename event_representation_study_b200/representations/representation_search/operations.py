"""Mirror of representations/representation_search/operations.py (reference :5-89): one scatter-reduce channel.

`events` is the reference's (n, 4) float64 array [x, y, t_s, p] with t_s already normalised to [0, 1].  The GPU
kernel works on integer timestamps, so t_s is mapped onto a 2^30-step integer grid spanning [min, max] and the
result mapped back (exact for counts / polarities; within 2e-9 absolute for timestamp channels)."""
import numpy as np

from ... import batched as eb
from ..._single import one_window

_SCALE = float(2**30 - 2)


class Operations(object):
    def __init__(self, func, aggregation, height, width):
        self.func = func
        self.aggregation = aggregation
        self.height = height
        self.width = width

    def __call__(self, events):
        return self.exec(events)

    def run(self, src, index):
        raise NotImplementedError("Operations.run is torch_scatter plumbing in the reference; the GPU path fuses it into exec()")

    def exec(self, events):
        if self.func not in eb.FUNCS:
            raise UnboundLocalError("local variable 'event_surface' referenced before assignment")  # what the reference raises
        if self.aggregation not in eb.AGGS:
            raise ValueError(f"unknown reduce '{self.aggregation}'")
        events = np.asarray(events)
        H, W = self.height, self.width
        if events.shape[0] == 0:
            return np.zeros((H, W))
        ts = events[:, 2].astype(np.float64)
        lo, hi = float(np.nanmin(ts)), float(np.nanmax(ts))
        span = hi - lo
        if not np.isfinite(span):
            raise ValueError("timestamps must be finite")
        ti = np.zeros(len(ts), np.int64) if span == 0 else np.rint((ts - lo) / span * _SCALE).astype(np.int64)
        ev = one_window(events[:, 0].astype(np.int64), events[:, 1].astype(np.int64), ti, events[:, 3].astype(np.int64), H, W)
        out = eb.mixed_density(ev, H, W, [0], [self.func], [self.aggregation], "SBN")[0, :, :, 0].double().cpu().numpy()
        if self.func.startswith("timestamp"):
            # kernel value v is on the normalised grid (t - lo) / span; map back to the caller's t_s scale
            touched = out != 0 if self.aggregation != "variance" else None
            if span == 0:
                # all timestamps equal: the kernel's 0/0 = NaN marks touched pixels
                nanmask = np.isnan(out)
                if self.aggregation == "variance":
                    out = np.where(nanmask, 0.0, out)
                elif self.aggregation == "sum":
                    cnt = eb.mixed_density(ev, H, W, [0], [self.func.replace("timestamp", "count")], ["sum"], "SBN")[0, :, :, 0].double().cpu().numpy()
                    out = cnt * lo
                else:
                    out = np.where(nanmask, lo, 0.0)
            elif self.aggregation == "variance":
                out = out * span * span
            elif self.aggregation == "sum":
                cnt = eb.mixed_density(ev, H, W, [0], [self.func.replace("timestamp", "count")], ["sum"], "SBN")[0, :, :, 0].double().cpu().numpy()
                out = out * span + cnt * lo
            else:  # mean / max: affine, untouched pixels stay 0
                cnt = eb.mixed_density(ev, H, W, [0], [self.func.replace("timestamp", "count")], ["max"], "SBN")[0, :, :, 0].double().cpu().numpy()
                out = np.where(cnt > 0, out * span + lo, 0.0)
        return out
