"""Mirror of representations/tore.py::events2ToreFeature (reference :6-83)."""
import numpy as np
import torch

from .. import batched as eb
from .._single import device, to_host


def events2ToreFeature(x, y, ts, pol, sampleTimes, k, frameSize):
    """-> float32 (frameSize[0], frameSize[1], 2k).  Pixels are [y-1, x-1] with numpy's negative wrap-around, events
    with ts >= sampleTimes are ignored, channels [0,k) hold the k most recent ages of pol > 0 events (ascending),
    [k,2k) the same for pol <= 0, log-compressed in float32 (tore.py:69-79)."""
    Hf, Wf = int(frameSize[0]), int(frameSize[1])
    dev = device()
    xi = np.asarray(x).astype(np.int64) - 1
    yi = np.asarray(y).astype(np.int64) - 1
    if xi.size and (xi.min() < -Wf or xi.max() >= Wf or yi.min() < -Hf or yi.max() >= Hf):
        raise IndexError("index out of bounds for the TORE frame")  # what numpy raises in the reference
    xi = np.where(xi < 0, xi + Wf, xi).astype(np.uint16).view(np.int16)
    yi = np.where(yi < 0, yi + Hf, yi).astype(np.uint16).view(np.int16)
    # integer microseconds on the GPU; fractional stamps are rounded to the nearest (the reference ages on the floats:
    # |d age| <= 0.5 us, i.e. <= 3e-3 relative in log(age + 1) - log(151) at the 150 us floor and far less above it)
    t = np.rint(np.asarray(ts, dtype=np.float64)).astype(np.int64) if np.asarray(ts).dtype.kind == "f" else np.asarray(ts).astype(np.int64)
    T = int(np.rint(float(np.asarray(sampleTimes).reshape(-1)[0])))
    if t.size and max(int(t.max()), T) - min(int(t.min()), T) >= 2**30:
        raise ValueError("timestamps span 2^30 us or more; the GPU kernels need a window shorter than about 17.9 min")
    p = np.where(np.asarray(pol) > 0, 1, -1).astype(np.int8)
    # a sentinel event at the sample time makes it the window's last timestamp; the kernel drops it (t < T is strict)
    xi, yi = np.append(xi, np.int16(0)), np.append(yi, np.int16(0))
    t, p = np.append(t, T), np.append(p, np.int8(1))
    t_dtype = np.int32 if (t.min() >= -2**31 and t.max() < 2**31) else np.int64
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ev = eb.EventBatch(up(xi), up(yi), up(t.astype(t_dtype)), up(p), np.array([0, len(t)], np.int64))
    return to_host(eb.tore(ev, Hf, Wf, int(k))[0])
