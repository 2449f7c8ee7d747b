"""CPU ORACLE (torch CPU) for the learned EST quantisation layer, forward pass.

TEST INFRASTRUCTURE ONLY (see oracle/representations.py for the rules).

Restates ev-YOLOv6/yolov6/models/learned_repr.py: ValueLayer.forward (:32-43) and QuantizationLayer.forward (:143-179,
with letterbox_image_batch :94-141) on CPU tensors - the reference class itself moves its MLP to "cuda" in __init__ and
ends forward with `.cuda()`.  Pinned by tests/golden/est_*.npz, produced by oracle/gen_golden_est.py with the reference's
own classes (ValueLayer trained by its init_kernel, QuantizationLayer.forward with Tensor.cuda patched to the identity).
"""
import numpy as np
import torch
import torch.nn.functional as F


def value_layer(weights, biases, x, negative_slope=0.1):
    """ValueLayer.forward: x (N,) float32 -> (N,) float32"""
    h = x[None, ..., None]
    for w, b in zip(weights[:-1], biases[:-1]):
        h = F.leaky_relu(F.linear(h, w, b), negative_slope)
    return F.linear(h, weights[-1], biases[-1]).squeeze()


def letterbox_image_batch(image_batch, size, color=114):
    """learned_repr.py:94-141"""
    bsz, c, oh, ow = image_batch.shape
    scale = min(size / ow, size / oh)
    nw, nh = int(ow * scale), int(oh * scale)
    resized = F.interpolate(image_batch, size=(nh, nw), mode="bilinear", align_corners=False)
    out = torch.full((bsz, c, size, size), fill_value=color, dtype=image_batch.dtype)
    top, left = (size - nh) // 2, (size - nw) // 2
    out[:, :, top:top + nh, left:left + nw] = resized
    return out


def est_forward(events, weights, biases, dim, image_size):
    """events: (N, 5) float32 tensor [x, y, t, p, b] (p in {0, 1}); returns (vox (B, 2C, H, W), letterboxed (B, 2C, S, S))"""
    events = events.clone()
    weights = [torch.as_tensor(w, dtype=torch.float32) for w in weights]
    biases = [torch.as_tensor(b, dtype=torch.float32) for b in biases]
    B = int((1 + events[-1, -1]).item())
    C, H, W = dim
    vox = events[0].new_full([int(2 * np.prod(dim) * B)], fill_value=0)
    x, y, t, p, b = events.t()
    for bi in range(B):
        t[events[:, -1] == bi] /= t[events[:, -1] == bi].max()
    idx_before_bins = x + W * y + 0 + W * H * C * p + W * H * C * 2 * b
    for i_bin in range(C):
        values = t * value_layer(weights, biases, t - i_bin / (C - 1))
        idx = torch.clamp((idx_before_bins + W * H * i_bin).long(), 0, int(np.prod(vox.shape)) - 1)
        vox.put_(idx, values, accumulate=True)
    vox = vox.view(-1, 2, C, H, W)
    vox = torch.cat([vox[:, 0, ...], vox[:, 1, ...]], 1)
    return vox, letterbox_image_batch(vox, image_size)
