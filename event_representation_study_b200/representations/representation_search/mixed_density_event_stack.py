"""Mirror of representations/representation_search/mixed_density_event_stack.py (reference :8-151)."""
import numpy as np

from ... import batched as eb
import torch

from ..._single import one_window, one_window_structured, to_host
from .operations import Operations


# Not in the reference: after this many `stack` calls with one (windows, functions, aggregations) tuple - counted per
# process, the reference builds a new instance per sample (optimized_representation.py:131-134) - the tuple is compiled into
# specialised kernels on a background thread (batched.specialize_mixed_density(wait=False): no call ever stalls; ~1 s
# later calls run the ERGO-12 pipeline with this tuple's kernels instead of the interpreted kernel).  A training run on a
# searched tuple crosses the threshold within its first batch; None disables it.
SPECIALIZE_AFTER_CALLS = 64
_TUPLE_CALLS = {}


class MixedDensityEventStack:
    def __init__(self, stack_size, num_of_events, height, width, indexes_functions_aggregations, stacking_type):
        self.stack_size = stack_size
        self.num_of_events = num_of_events
        self.height = height
        self.width = width
        self.indexes_functions_aggregations = indexes_functions_aggregations
        self.stacking_type = stacking_type

    def _maybe_specialize(self, w, f, a, n):
        if SPECIALIZE_AFTER_CALLS is None or self.stacking_type not in ("SBN", "SBT"):
            return
        key = (self.stacking_type, tuple(w), tuple(f), tuple(a))
        calls = _TUPLE_CALLS.get(key, 0) + 1
        if calls <= SPECIALIZE_AFTER_CALLS:
            if len(_TUPLE_CALLS) > 4096:
                _TUPLE_CALLS.clear()
            _TUPLE_CALLS[key] = calls
        if calls == SPECIALIZE_AFTER_CALLS:
            eb.specialize_mixed_density(w, f, a, self.stacking_type, max_events_per_window=max(int(n), 1 << 20), wait=False)

    def _spec(self):
        w, f, a = self.indexes_functions_aggregations
        n = self.stack_size
        # make_stack indexes the three lists with range(stack_size): shorter lists raise -> zero channel (reference :116-127)
        pad = lambda lst, fill: [lst[i] if i < len(lst) else fill for i in range(n)]
        return pad(list(w), 127), pad(list(f), "<missing>"), pad(list(a), "<missing>")

    def _run(self, x, y, p, t, _scale=None, _records=None):
        if len(t) == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")  # t.min() (reference :33)
        w, f, a = self._spec()
        try:
            ev = one_window_structured(_records, self.height, self.width) if _records is not None else None
            if ev is None:
                ev = one_window(x, y, t, p, self.height, self.width)
        except IndexError:
            # an out-of-range pixel makes every torch_scatter call raise inside make_stack: all channels zero (reference :120-127)
            return np.zeros((self.height, self.width, self.stack_size), np.float64)
        self._maybe_specialize(w, f, a, len(t))
        return to_host(eb.mixed_density(ev, self.height, self.width, w, f, a, self.stacking_type)[0], torch.float64, _scale)

    def stack(self, event_sequence, _scale=None):
        """_scale (not in the reference): multiply the float64 result on the GPU before it is copied back - the `rep *= 255`
        of get_item_transform (gen1_transforms.py:36) without a pass over 88 MB on one host core."""
        x, y, p, t = event_sequence["x"], event_sequence["y"], event_sequence["p"], event_sequence["t"]
        assert len(x) == len(y) == len(p) == len(t)
        if isinstance(event_sequence, np.ndarray) and all(v.dtype == np.dtype("<i4") for v in (x, y, p, t)):
            # the `<i4` record layout of the detection loaders: the reference's .astype(np.int32) / (np.int64) are value
            # preserving, so the records go to the GPU as they are (one copy) and are split there
            return self._run(x, y, p, t, _scale, _records=event_sequence)
        return self._run(x.astype(np.int32), y.astype(np.int32), p.astype(np.int32), t.astype(np.int64), _scale)

    def create_windows(self, x, y, p, t):
        """The 7 (SBN) / 8 (SBT) event subsets, as host array views like the reference (reference :48-109).
        Pure slicing - kept for API compatibility; `stack` never materialises them."""
        windows = [(x, y, p, t)]
        n3 = x.shape[0] // 3
        if self.stacking_type == "SBN":
            for i in range(3):
                s = slice(i * n3, (i + 1) * n3)
                windows.append((x[s], y[s], p[s], t[s]))
            c = len(t)
            for _ in range(3):
                c = c // 2
                x, y, p, t = x[c:], y[c:], p[c:], t[c:]
                windows.append((x, y, p, t))
        elif self.stacking_type == "SBT":
            f = 1 / 3
            for i in range(3):
                m = np.logical_and(t <= (i + 1) * f, t >= i * f)
                windows.append((x[m], y[m], p[m], t[m]))
            factor = 1
            for _ in range(4):
                factor = factor / 2
                m = t <= factor
                x, y, p, t = x[m], y[m], p[m], t[m]
                windows.append((x, y, p, t))
        return windows

    def make_stack(self, x, y, p, t):
        """-> list of {name: (H, W) float64} like the reference (reference :111-130)."""
        rep = self._run(x, y, p, t)
        w, f, a = self._spec()
        out = []
        for i in range(self.stack_size):
            valid = isinstance(f[i], str) and isinstance(a[i], str) and f[i] in eb.FUNCS and a[i] in eb.AGGS
            name = "_".join([f[i].capitalize(), a[i].capitalize()]) if valid else ""
            out.append({name: rep[:, :, i]})
        return out

    def stack_data(self, x, y, p, t_s, func, aggregation):
        assert len(x) == len(y) == len(p) == len(t_s)
        events = np.concatenate([x[..., np.newaxis], y[..., np.newaxis], t_s[..., np.newaxis], p[..., np.newaxis]], axis=1)
        surface = Operations(func, aggregation, self.height, self.width)(events)
        return {"_".join([func.capitalize(), aggregation.capitalize()]): surface}
