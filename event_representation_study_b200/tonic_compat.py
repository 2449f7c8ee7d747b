"""GPU stand-ins for the two tonic transforms the reference dispatches on
(representations/gen1_transforms.py:21-25, :44-49; n_imagenet/real_cnn_model/data/imagenet.py:1017-1020, :1072-1074).
`str(ToVoxelGrid)` / `str(ToImage)` contain the class name, which is all `get_item_transform` looks at.
Semantics restated from tonic 1.x (to_voxel_grid_numpy / to_frame_numpy); tonic itself is not installable offline,
so parity with the real package is unpinned."""
import numpy as np

from . import batched as eb
from ._single import one_window


class ToVoxelGrid:
    def __init__(self, sensor_size, n_time_bins):
        self.sensor_size = tuple(sensor_size)
        self.n_time_bins = int(n_time_bins)

    def __call__(self, events):
        W, H = self.sensor_size[0], self.sensor_size[1]
        if len(events) < 2:
            raise IndexError("ToVoxelGrid needs at least two events")  # t[-1] - t[0] on a shorter stream fails in tonic
        p = events["p"]
        p[p == 0] = -1  # tonic rewrites the caller's polarity field in place
        t = np.asarray(events["t"])
        if t.dtype.kind == "f" and not np.array_equal(t, np.floor(t)):
            # fractional timestamps (N-ImageNet hands seconds as float64, imagenet.py:1002-1020): tonic only uses
            # (t - t[0]) / (t[-1] - t[0]), so put that on a 2^30 integer grid (error < 1e-9 of the window)
            span = float(t[-1] - t[0])
            if not (span > 0 and np.all(t >= t[0]) and np.all(t <= t[-1])):
                raise ValueError("ToVoxelGrid on the GPU needs t[0] <= t <= t[-1] with t[-1] > t[0] for fractional timestamps")
            t = np.rint((t - t[0]) / span * float(2**30 - 2)).astype(np.int64)
        ev = one_window(events["x"], events["y"], t, p, H, W)
        out = eb.voxel_grid(ev, H, W, self.n_time_bins, "tonic")[0]
        return out.double().cpu().numpy()[:, None]  # (n_bins, 1, H, W) float64


class ToImage:
    def __init__(self, sensor_size):
        self.sensor_size = tuple(sensor_size)

    def __call__(self, events):
        W, H = self.sensor_size[0], self.sensor_size[1]
        p = np.asarray(events["p"])
        if p.size and (p.min() < 0 or p.max() > 1):
            raise IndexError("polarity index out of range for a 2-channel frame")
        ev = one_window(events["x"], events["y"], np.arange(len(p), dtype=np.int64), p, H, W)
        return eb.histogram(ev, H, W)[0].cpu().numpy().astype(np.int16)  # (2, H, W) int16 counts
