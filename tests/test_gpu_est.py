"""EST (learned quantisation layer, forward) on the GPU against fixtures produced by the reference's own ValueLayer /
QuantizationLayer (tests/golden/est_*.npz) and against the torch-CPU oracle on larger seeded inputs.  Bar: 1e-5 relative
with an absolute floor of 2e-5 - the reference evaluates its MLP in float32 (7e-7 absolute per evaluation against the
exact piecewise-linear function, measured) and both sides add a few to a few dozen such values per voxel with float
atomics in arbitrary order."""
import numpy as np
import pytest

from conftest import assert_close, golden, load

pytestmark = pytest.mark.gpu
ATOL = 2e-5


@pytest.fixture(scope="module")
def M(cuda_device):
    import event_representation_study_b200.batched as eb
    import event_representation_study_b200.est as est
    return eb, est


def _batch(eb, events, B):
    wins = []
    for b in range(B):
        e = events[events[:, 4] == b]
        wins.append({"x": e[:, 0].astype(np.uint16), "y": e[:, 1].astype(np.uint16), "t": e[:, 2].astype(np.int64), "p": e[:, 3].astype(np.int8)})
    return eb.pack_events(wins, "cuda")


def _tables(est, ws, bs):
    import torch
    br, sl, ic = est.compile_value_layer(ws, bs, 0.1)
    return tuple(torch.as_tensor(v, dtype=torch.float64, device="cuda") for v in (br, sl, ic))


@pytest.mark.parametrize("name,path", golden("est_*"), ids=[n for n, _ in golden("est_*")])
def test_est_matches_reference_fixtures(M, name, path):
    eb, est = M
    g = load(path)
    C, H, W = (int(v) for v in g["dim"])
    B = g["out"].shape[0]
    ev = _batch(eb, g["events"], B)
    tables = _tables(est, [g[f"w{i}"] for i in range(3)], [g[f"b{i}"] for i in range(3)])
    vox = est.quantize(ev, H, W, C, tables).permute(0, 3, 1, 2).cpu().numpy()
    assert_close(vox, g["vox"], rtol=1e-5, atol=ATOL, what=name + " vox")
    out = est.forward(ev, H, W, C, tables, image_size=int(g["image_size"])).cpu().numpy()
    assert_close(out, g["out"], rtol=1e-5, atol=ATOL, what=name + " letterboxed")


def test_est_gen1_size_vs_oracle(M):
    """the reference's configuration: dim = (6, 240, 304), image_size 640 (yolo.py:56-61), two windows of 50 k events"""
    import torch
    from oracle import est as oest
    eb, est = M
    g = load(golden("est_small")[0][1])
    ws, bs = [g[f"w{i}"] for i in range(3)], [g[f"b{i}"] for i in range(3)]
    C, H, W, S = 6, 240, 304, 640
    rng = np.random.default_rng(12)
    rows = []
    for b, n in enumerate((50000, 31000)):
        t = np.sort(rng.integers(0, 100000, n)).astype(np.float32)
        rows.append(np.stack([rng.integers(0, W, n), rng.integers(0, H, n), t, rng.integers(0, 2, n), np.full(n, b)], 1))
    events = np.concatenate(rows).astype(np.float32)
    with torch.no_grad():
        vox_w, out_w = oest.est_forward(torch.tensor(events), ws, bs, (C, H, W), S)
    ev = _batch(eb, events, 2)
    tables = _tables(est, ws, bs)
    assert_close(est.quantize(ev, H, W, C, tables).permute(0, 3, 1, 2).cpu().numpy(), vox_w.numpy(), rtol=1e-5, atol=ATOL, what="vox")
    assert_close(est.forward(ev, H, W, C, tables, image_size=S).cpu().numpy(), out_w.numpy(), rtol=1e-5, atol=ATOL, what="out")


def test_est_argument_checks(M):
    import torch
    from event_representation_study_b200._lib import EvrepError
    eb, est = M
    ev = _batch(eb, np.array([[1, 1, 5, 1, 0], [2, 2, 9, 0, 0]], np.float32), 1)
    tab = tuple(torch.zeros(k, dtype=torch.float64, device="cuda") for k in (0, 1, 1))
    with pytest.raises(EvrepError):
        est.quantize(ev, 8, 8, 1, tab)  # C - 1 == 0: the reference divides by zero
    out = est.quantize(ev, 8, 8, 2, tab)  # f == 0 everywhere
    assert float(out.abs().max()) == 0.0


def test_est_batch_with_an_empty_window(M):
    """an empty window gives an all-zero grid (the reference would fail on max() of an empty tensor)"""
    import torch
    eb, est = M
    wins = [{"x": np.array([3, 3], np.uint16), "y": np.array([1, 1], np.uint16), "t": np.array([10, 20], np.int64), "p": np.array([1, 0], np.int8)},
            {"x": np.zeros(0, np.uint16), "y": np.zeros(0, np.uint16), "t": np.zeros(0, np.int64), "p": np.zeros(0, np.int8)}]
    ev = eb.pack_events(wins, "cuda")
    tab = (torch.zeros(0, dtype=torch.float64, device="cuda"), torch.tensor([0.0], dtype=torch.float64, device="cuda"),
           torch.tensor([2.0], dtype=torch.float64, device="cuda"))  # f(u) = 2
    out = est.quantize(ev, 4, 6, 2, tab)
    assert float(out[1].abs().max()) == 0.0
    # window 0: tn = 0.5 and 1.0; every bin adds tn * 2 at (y = 1, x = 3), polarity 1 -> channels 2, 3; polarity 0 -> channels 0, 1
    assert out[0, 1, 3].cpu().numpy().tolist() == [2.0, 2.0, 1.0, 1.0]
    assert float(out[0].sum()) == 6.0


def test_est_backward_matches_autograd_through_the_reference_forward(M):
    """training the layer (learned_repr.py:9-77 under autograd): gradients of a random linear functional of the quantised
    tensor with respect to every ValueLayer weight, from evrep_est_backward_batched + the segment surrogate, against
    torch autograd through the reference's forward restated in float64 (C MLP evaluations per event, put_ accumulate)"""
    import torch
    eb, est = M
    torch.manual_seed(0)
    rng = np.random.default_rng(3)
    C, H, W, B, n = 6, 12, 16, 2, 700
    mlp = torch.nn.ModuleList([torch.nn.Linear(1, 20), torch.nn.Linear(20, 20), torch.nn.Linear(20, 1)])

    class VL(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mlp = mlp
    vl = VL()
    ev_np = np.stack([rng.integers(0, W, B * n), rng.integers(0, H, B * n), np.concatenate([np.sort(rng.integers(1, 90000, n)) for _ in range(B)]),
                      rng.integers(0, 2, B * n), np.repeat(np.arange(B), n)], 1).astype(np.float32)
    ev = _batch(eb, ev_np, B)
    proj = torch.as_tensor(rng.standard_normal((B, H, W, 2 * C)), dtype=torch.float32, device="cuda")
    out = est.quantize_trainable(ev, H, W, C, vl, 0.1, t_float=torch.as_tensor(ev_np[:, 2]).cuda())
    (out * proj).sum().backward()
    got = [p.grad.detach().double().cpu().clone() for p in vl.parameters()]
    for p in vl.parameters():
        p.grad = None
    # the reference forward, float64, on the CPU
    e = torch.as_tensor(ev_np, dtype=torch.float64)
    x, y, t, pol, b = e.t()
    t = t.clone()
    t32 = torch.as_tensor(ev_np[:, 2])
    for bi in range(B):
        m = e[:, -1] == bi
        t[m] = (t32[m] / t32[m].max()).double()  # the layer normalises in float32
    vox = torch.zeros(B * H * W * 2 * C, dtype=torch.float64)
    params = [q.double() for q in vl.parameters()]
    for i_bin in range(C):
        u = (t.float() - i_bin / (C - 1)).double()  # float32 tensor minus a Python float, as in the reference
        vals = t * est._mlp_double(list(vl.parameters()), u, 0.1)
        idx = (((b * H + y) * W + x) * (2 * C) + pol * C + i_bin).long()  # (B, H, W, 2C) layout of the CUDA output
        vox = vox.index_put((idx,), vals, accumulate=True)
    (vox * proj.double().cpu().reshape(-1)).sum().backward()
    want = [p.grad.detach().double().cpu() for p in vl.parameters()]
    for g_, w_ in zip(got, want):
        scale = max(float(w_.abs().max()), 1e-12)
        assert float((g_ - w_).abs().max()) <= 2e-5 * scale, (float((g_ - w_).abs().max()), scale)
    # and the forward of the trainable path is the inference path
    tables = est.value_layer_tables(vl)
    assert torch.allclose(out.detach(), est.quantize(ev, H, W, C, tables, t_float=torch.as_tensor(ev_np[:, 2]).cuda()), rtol=1e-6, atol=1e-7)
