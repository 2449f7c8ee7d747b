"""The N-ImageNet loader wrappers around the hot path (n_imagenet/real_cnn_model/data/imagenet.py:1002-1134), as a maintainer
would have them after switching the four imports at the top of imagenet.py to this package.  TEST
INFRASTRUCTURE: the wrappers are callers of the drop-in boundary (caller ii), not part of the product; the
bodies below are the reference's own call sequences so that the fixtures made from the reference wrappers
(oracle/gen_golden_nimagenet.py) pin the mirrors they call.

Two reference lines cannot run on current numpy / at all and are restated by intent: `rep.float()` on a numpy array
(imagenet.py:1077) and `np.int` (imagenet.py:1125-1126)."""
import numpy as np
import numpy.lib.recfunctions as rfn
import torch

# the four-import switch --------------------------------------------------------------------------------------------
from event_representation_study_b200 import tonic_compat as tonic_transforms
from event_representation_study_b200.representations.event_stack import EventStack
from event_representation_study_b200.representations.optimized_representation import get_optimized_representation
from event_representation_study_b200.representations.time_surface import ToTimesurface
from event_representation_study_b200.representations.tore import events2ToreFeature
# ---------------------------------------------------------------------------------------------------------------------

IMAGE_H = IMAGE_W = 224


def fix_events_training(events):
    events = rfn.unstructured_to_structured(events)
    events.dtype = [("x", "<f8"), ("y", "<f8"), ("t", "<f8"), ("p", "<f8")]
    return events


def _prep(event_tensor, augment, kwargs):
    if augment is not None:
        event_tensor = augment(event_tensor)
    return fix_events_training(event_tensor.numpy()), kwargs.get("height", IMAGE_H), kwargs.get("width", IMAGE_W)


def reshape_then_voxel_grid(event_tensor, augment=None, **kwargs):
    data, H, W = _prep(event_tensor, augment, kwargs)
    rep = tonic_transforms.ToVoxelGrid((W, H, 2), n_time_bins=12)(data)
    return torch.tensor(rep.transpose(0, 2, 3, 1)[..., 0]).float()


def reshape_then_optimized(event_tensor, augment=None, **kwargs):
    data, H, W = _prep(event_tensor, augment, kwargs)
    rep = get_optimized_representation(data, data.shape[0], H, W)
    return torch.tensor(rep.transpose(2, 0, 1)).float()


def reshape_then_event_stack(event_tensor, augment=None, **kwargs):
    data, H, W = _prep(event_tensor, augment, kwargs)
    data["p"] = (data["p"] + 1) // 2
    tr = EventStack(12, data.shape[0], H, W)
    post = tr.post_stack(tr.pre_stack(data, data[-1]["t"]))
    return torch.tensor(post.transpose(3, 0, 1, 2)[..., 0]).float()


def reshape_then_to_image(event_tensor, augment=None, **kwargs):
    data, H, W = _prep(event_tensor, augment, kwargs)
    tr = tonic_transforms.ToImage((W, H, 2))
    data["p"] = (data["p"] + 1) // 2
    return torch.tensor(np.ascontiguousarray(tr(data).transpose(1, 2, 0))).float()


def reshape_then_tore(event_tensor, augment=None, **kwargs):
    data, H, W = _prep(event_tensor, augment, kwargs)
    x, y, ts, pol = data["x"], data["y"], data["t"], data["p"]
    rep = events2ToreFeature(x - min(x) + 1, y - min(y) + 1, ts, pol, ts[-1], 6, (H, W))
    return torch.tensor(rep.transpose(2, 0, 1)).float()


def reshape_then_time_surface(event_tensor, augment=None, **kwargs):
    data, H, W = _prep(event_tensor, augment, kwargs)
    data["p"] = ((data["p"] + 1) / 2).astype(np.int8)
    tr = ToTimesurface(sensor_size=(W, H, 2), surface_dimensions=None, tau=50000, decay="exp")
    t = data["t"]
    idx = np.searchsorted((t - t[0]) / (t[-1] - t[0]) * 6, np.arange(6) + 1)
    data["x"] = data["x"].astype(int)
    data["y"] = data["y"].astype(int)
    rep = tr(data, idx)
    rep = rep.reshape((-1, rep.shape[-2], rep.shape[-1]))
    return torch.tensor(rep.transpose(1, 2, 0)).float()
