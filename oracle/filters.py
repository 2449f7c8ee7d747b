"""CPU ORACLE for ev-licious' stateful per-pixel filters.  TEST INFRASTRUCTURE ONLY.

Plain-Python restatements of the numba loops of ev-licious/src/evlicious/tools/utils.py (small inputs only):
_filter_events_resize :143-158, _background_activity_filter :169-178, _contrast_threshold_control :184-191,
_refractory_period :193-200, and of HotPixel.calibrate (tools/filters.py:27-47).  Pinned by
tests/golden/filter_*.npz, produced by oracle/gen_golden_filters.py with the reference's own numba functions."""
import numpy as np


def refractory_period(mask, x, y, t, period, last_timestamp):
    for i in range(len(x)):
        if t[i] - last_timestamp[y[i], x[i]] < period:
            mask[i] = False
            continue
        last_timestamp[y[i], x[i]] = t[i]
    return mask


def contrast_threshold_control(activity, mask, x, y, p, factor):
    for i in range(len(x)):
        activity[y[i], x[i]] += p[i]
        if np.abs(activity[y[i], x[i]]) >= factor:
            mask[i] = True
            activity[y[i], x[i]] = 0
    return mask


def filter_events_resize(x, y, p, mask, change_map, fx, fy):
    for i in range(len(x)):
        x_l = x[i] // fx
        y_l = y[i] // fy
        change_map[y_l, x_l] = np.float32(np.float64(change_map[y_l, x_l]) + p[i] * 1.0 / (fx * fy))
        if np.abs(change_map[y_l, x_l]) >= 1:
            mask[i] = True
            change_map[y_l, x_l] -= p[i]
    return mask, change_map


def background_activity_filter(mask, timestamps, x, y, t, depth_us, radius=1):
    for i in range(len(x)):
        x_, y_, t_ = int(x[i]), int(y[i]), t[i]
        t_last = timestamps[y_, x_]
        mask[i] = not (t_last > 0 and t_ - t_last > depth_us)
        timestamps[max(y_ - radius, 0):y_ + radius, max(x_ - radius, 0):x_ + radius] = t_
    return mask


def hot_pixel_mask(x, y, H, W, threshold=0.6):
    """HotPixel.calibrate (tools/filters.py:27-47, debug=False)"""
    count = np.zeros((H, W))
    np.add.at(count, (y, x), 1.0)
    mask = count / np.max(count) < threshold
    if float(np.min(count[~mask])) / np.max(count[mask]) > 2:
        return mask
    return np.ones((H, W)) > 0
