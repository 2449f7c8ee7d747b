"""Restatement of the two tonic 1.x transforms the reference calls
(gen1_transforms.py:21-25,44-49; gen4_transforms.py:15-19,38-43; imagenet.py:1017-1020,1072-1074).

TEST INFRASTRUCTURE ONLY. tonic is not installable here, so this is restated from the published
tonic 1.4 sources (`tonic/functional/to_voxel_grid.py::to_voxel_grid_numpy`,
`tonic/functional/to_frame.py::to_frame_numpy`, `tonic/transforms.py::ToVoxelGrid/ToImage`);
parity with real tonic is UNPINNED (no tonic tests or vectors exist in the reference tree).
"""
from dataclasses import dataclass
from typing import Tuple

import numpy as np


def to_voxel_grid_numpy(events, sensor_size, n_time_bins=10):
    assert sensor_size[2] == 2
    voxel_grid = np.zeros((n_time_bins, sensor_size[1], sensor_size[0]), float).ravel()
    ts = n_time_bins * (events["t"].astype(float) - events["t"][0]) / (events["t"][-1] - events["t"][0])
    xs = events["x"].astype(int)
    ys = events["y"].astype(int)
    pols = events["p"]
    pols[pols == 0] = -1  # in place on the caller's array, as tonic does
    tis = ts.astype(int)
    dts = ts - tis
    vals_left = pols * (1.0 - dts)
    vals_right = pols * dts
    valid = tis < n_time_bins
    np.add.at(voxel_grid, xs[valid] + ys[valid] * sensor_size[0] + tis[valid] * sensor_size[0] * sensor_size[1], vals_left[valid])
    valid = (tis + 1) < n_time_bins
    np.add.at(voxel_grid, xs[valid] + ys[valid] * sensor_size[0] + (tis[valid] + 1) * sensor_size[0] * sensor_size[1], vals_right[valid])
    return np.reshape(voxel_grid, (n_time_bins, 1, sensor_size[1], sensor_size[0]))


@dataclass(frozen=True)
class ToVoxelGrid:
    sensor_size: Tuple[int, int, int]
    n_time_bins: int

    def __call__(self, events):
        return to_voxel_grid_numpy(events.copy(), self.sensor_size, self.n_time_bins)


@dataclass(frozen=True)
class ToImage:
    sensor_size: Tuple[int, int, int]

    def __call__(self, events):
        # to_frame_numpy(event_count=len(events)) -> one int16 frame (1, P, H, W); ToImage squeezes it
        frames = np.zeros((1, self.sensor_size[2], self.sensor_size[1], self.sensor_size[0]), dtype=np.int16)
        if len(events):
            np.add.at(frames, (0, events["p"].astype(int), events["y"].astype(int), events["x"].astype(int)), 1)
        return frames.squeeze(0)
