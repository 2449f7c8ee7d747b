"""EST (learned Event Spike Tensor quantisation layer): forward and backward with respect to the ValueLayer weights.

Reference: ev-YOLOv6/yolov6/models/learned_repr.py - ValueLayer (:9-77, an MLP 1 -> 100 -> 100 -> 1 with LeakyReLU(0.1)
applied to ONE scalar per event and bin) and QuantizationLayer.forward (:143-179).

A LeakyReLU network with a scalar input is a piecewise-linear function of that scalar.  `compile_value_layer` turns the
weights into that function exactly (float64 breakpoints, slope and intercept per segment): hidden layer by hidden layer,
every segment on which the current activations are affine is cut where a pre-activation changes sign.  The CUDA kernel
(csrc/est.cu) then evaluates the layer with a binary search and one FMA per (event, bin) instead of 2 x 10^4 MACs.

Training: inside segment j the layer is f(u) = a_j u + c_j with a_j, c_j smooth functions of the weights, so
dL/dtheta = sum_j (G1_j da_j/dtheta + G0_j dc_j/dtheta), where G0_j / G1_j are the sums of g and g u over the (event, bin)
samples of the segment (g = upstream gradient times t).  `evrep_est_backward_batched` reduces the upstream gradient to those
2 (K + 1) numbers in one pass over the events; `EstQuantize.backward` then evaluates the torch MLP at two points per segment
with coefficients that reproduce (G0, G1) and lets autograd produce the parameter gradients - exactly the gradient the
reference gets from back-propagating through C MLP evaluations per event (learned_repr.py:164-172).
"""
import numpy as np
import torch

from . import batched as eb
from ._lib import check, lib


def compile_value_layer(weights, biases, negative_slope=0.1, lo=-2.0, hi=2.0):
    """weights[l]: (out_l, in_l) arrays of the Linear layers (in_0 == 1, out_last == 1), biases[l]: (out_l,).
    -> (breaks (K,), slope (K + 1,), icpt (K + 1,)) float64: f(u) = slope[j] * u + icpt[j] with j = #{breaks <= u}.
    Exact on [lo, hi] (the layer sees u = t - i / (C - 1) in [-1, 1]); outside, the outermost segments are extended."""
    W = [np.asarray(w, np.float64) for w in weights]
    Bv = [np.asarray(b, np.float64).reshape(-1) for b in biases]
    if W[0].shape[1] != 1 or W[-1].shape[0] != 1:
        raise ValueError("the value layer maps one scalar to one scalar")
    # a segment is (a, b, A, c): on a <= u < b the current activations are A * u + c
    segs = [(float(lo), float(hi), np.ones(1), np.zeros(1))]
    for l in range(len(W) - 1):  # hidden layers
        nxt = []
        for a, b, A, c in segs:
            alpha, beta = W[l] @ A, W[l] @ c + Bv[l]  # pre-activations alpha * u + beta
            with np.errstate(divide="ignore", invalid="ignore"):
                roots = -beta / alpha
            cuts = np.unique(roots[np.isfinite(roots) & (roots > a) & (roots < b)])
            edges = np.concatenate([[a], cuts, [b]])
            for e0, e1 in zip(edges[:-1], edges[1:]):
                mid = 0.5 * (e0 + e1)
                s = np.where(alpha * mid + beta > 0, 1.0, negative_slope)
                nxt.append((e0, e1, s * alpha, s * beta))
        segs = nxt
    slope = np.array([float((W[-1] @ A)[0]) for _, _, A, _ in segs])
    icpt = np.array([float((W[-1] @ c)[0] + Bv[-1][0]) for _, _, _, c in segs])
    breaks = np.array([s[0] for s in segs[1:]], np.float64)
    return breaks, slope, icpt


def value_layer_tables(value_layer, negative_slope=0.1, device="cuda"):
    """Compile a torch ValueLayer-like module (attribute `.mlp`: ModuleList of nn.Linear) and upload the tables."""
    ws = [m.weight.detach().cpu().double().numpy() for m in value_layer.mlp]
    bs = [m.bias.detach().cpu().double().numpy() for m in value_layer.mlp]
    br, sl, ic = compile_value_layer(ws, bs, negative_slope)
    dev = torch.device(device)
    return tuple(torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for v in (br, sl, ic))


def quantize(ev, H, W, C, tables, t_float=None):
    """QuantizationLayer.forward up to (not including) the letterbox: EventBatch -> (B, H, W, 2C) float32 CUDA tensor,
    channel = p * C + bin.  `tables` from value_layer_tables.  t_float: the timestamps as float32 (defaults to ev.t.float(),
    i.e. what `torch.tensor(events).float()` gives the reference)."""
    breaks, slope, icpt = tables
    dev = ev.x.device
    t = ev.t.float() if t_float is None else t_float.float().contiguous()
    B = len(ev.offsets) - 1
    out = torch.empty((B, H, W, 2 * C), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = eb._workspace(dev, stream, lib.evrep_est_workspace_bytes(B))
    offs = np.ascontiguousarray(ev.offsets, np.int64)
    check(lib.evrep_est_quantize_batched(ev.x.data_ptr(), ev.y.data_ptr(), t.data_ptr(), ev.p.data_ptr(), offs.ctypes.data, B, H, W, C,
                                         breaks.data_ptr(), slope.data_ptr(), icpt.data_ptr(), int(breaks.numel()), out.data_ptr(),
                                         ws.data_ptr(), ws.numel(), stream))
    return out


def _mlp_double(params, u, negative_slope):
    """ValueLayer.forward (learned_repr.py:32-43) in float64 on a vector of scalars; params = [w0, b0, w1, b1, ...]"""
    h = u.reshape(-1, 1)
    n_layers = len(params) // 2
    for l in range(n_layers):
        h = torch.nn.functional.linear(h, params[2 * l].double(), params[2 * l + 1].double())
        if l + 1 < n_layers:
            h = torch.nn.functional.leaky_relu(h, negative_slope)
    return h.reshape(-1)


class EstQuantize(torch.autograd.Function):
    """out = quantize(events; value-layer weights), differentiable with respect to the weights.
    apply(ev, H, W, C, negative_slope, t_float, *params) with params = [w0, b0, w1, b1, ...] (the nn.Linear weights and biases
    of ValueLayer.mlp, any device) -> (B, H, W, 2C) float32 CUDA tensor."""

    @staticmethod
    def forward(ctx, ev, H, W, C, negative_slope, t_float, *params):
        ws = [p.detach().cpu().double().numpy() for p in params[0::2]]
        bs = [p.detach().cpu().double().numpy() for p in params[1::2]]
        br, sl, ic = compile_value_layer(ws, bs, negative_slope)
        dev = ev.x.device
        tables = tuple(torch.as_tensor(v, dtype=torch.float64, device=dev).contiguous() for v in (br, sl, ic))
        t = ev.t.float() if t_float is None else t_float.float().contiguous()
        out = quantize(ev, H, W, C, tables, t_float=t)
        ctx.ev, ctx.geom, ctx.tables, ctx.t, ctx.slope = ev, (H, W, C), tables, t, negative_slope
        ctx.save_for_backward(*params)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ev, (H, W, C), (breaks, _, _), t = ctx.ev, ctx.geom, ctx.tables, ctx.t
        params = ctx.saved_tensors
        dev = ev.x.device
        K = int(breaks.numel())
        seg = torch.empty(2 * (K + 1), dtype=torch.float64, device=dev)
        g = grad_out.contiguous().float()
        B = len(ev.offsets) - 1
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws = eb._workspace(dev, stream, lib.evrep_est_workspace_bytes(B))
        offs = np.ascontiguousarray(ev.offsets, np.int64)
        check(lib.evrep_est_backward_batched(ev.x.data_ptr(), ev.y.data_ptr(), t.data_ptr(), ev.p.data_ptr(), offs.ctypes.data, B, H, W, C,
                                             breaks.data_ptr(), K, g.data_ptr(), seg.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        G0, G1 = seg[:K + 1], seg[K + 1:]
        # two points strictly inside every segment (the outermost segments are half lines)
        if K:
            lo = torch.cat([breaks[:1] - 2.0, breaks])
            hi = torch.cat([breaks, breaks[-1:] + 2.0])
        else:
            lo, hi = torch.tensor([-2.0], dtype=torch.float64, device=dev), torch.tensor([2.0], dtype=torch.float64, device=dev)
        u1, u2 = lo + (hi - lo) / 3.0, lo + 2.0 * (hi - lo) / 3.0
        beta = (G1 - G0 * u1) / (u2 - u1)
        alpha = G0 - beta
        with torch.enable_grad():
            ps = [p.detach().to(dev).requires_grad_(True) for p in params]
            f1, f2 = _mlp_double(ps, u1, ctx.slope), _mlp_double(ps, u2, ctx.slope)
            surrogate = (alpha * f1 + beta * f2).sum()  # = sum_j (G1_j a_j + G0_j c_j) as a function of the weights
            grads = torch.autograd.grad(surrogate, ps, allow_unused=True)
        out = [None if gr is None else gr.to(device=p.device, dtype=p.dtype) for gr, p in zip(grads, params)]
        return (None, None, None, None, None, None, *out)


def quantize_trainable(ev, H, W, C, value_layer, negative_slope=0.1, t_float=None):
    """Differentiable `quantize`: gradients flow to `value_layer.mlp`'s weights and biases (a torch ValueLayer-like module)."""
    params = []
    for m in value_layer.mlp:
        params += [m.weight, m.bias]
    return EstQuantize.apply(ev, H, W, C, float(negative_slope), t_float, *params)


def forward(ev, H, W, C, tables, image_size=640, t_float=None):
    """The whole QuantizationLayer.forward (learned_repr.py:143-179): quantise, then letterbox_image_batch (bilinear
    interpolate with torch's float32 coordinates, pad 114) -> (B, 2C, image_size, image_size) float32."""
    vox = quantize(ev, H, W, C, tables, t_float)
    return eb.detector_input(vox, image_size, mode="letterbox", interp="linear_torch", scale_in=1.0, scale_out=1.0, pad_value=114.0,
                             reverse_channels=False)
