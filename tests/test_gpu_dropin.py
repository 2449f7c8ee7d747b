"""The drop-in mirrors of the reference's Python interfaces (event_representation_study_b200.representations,
.evlicious, .tonic_compat) against the golden fixtures produced by the reference files: same call, same
shapes / dtypes / side effects, values within the parity bars of test_gpu_parity.py.  These read like the calls
in representations/gen1_transforms.py and gen1_compute.py."""
import numpy as np
import pytest

from conftest import assert_close, golden, load
from event_representation_study_b200.synth import structured

pytestmark = pytest.mark.gpu
RTOL, VAR_ATOL, TORE_ATOL = 1e-5, 2e-7 * 255, 1e-6 * 255


@pytest.fixture(scope="module", autouse=True)
def _need_cuda(cuda_device):
    return cuda_device


def ids(cases):
    return [c[0] for c in cases]


def ev_of(g, dtype="<i4"):
    return structured({k: g[k] for k in "xytp"}, dtype)


def test_get_item_transform_all_branches():
    """representations/gen1_transforms.py:12-89, every branch, including the x255 and the in-place polarity rewrites."""
    from event_representation_study_b200.representations.gen1_transforms import get_item_transform
    from event_representation_study_b200.representations import gen4_transforms
    from event_representation_study_b200.representations.event_stack import EventStack
    from event_representation_study_b200.representations.time_surface import ToTimesurface
    from event_representation_study_b200.representations.representation_search.mixed_density_event_stack import MixedDensityEventStack
    from event_representation_study_b200 import tonic_compat as tt
    G = {n[len("dispatch_"):]: load(p) for n, p in golden("dispatch_*")}
    tol = {"ToVoxelGrid": dict(rtol=RTOL, atol=2e-6 * 255), "MixedDensityEventStack": dict(rtol=RTOL, atol=VAR_ATOL),
           "EventStack": None, "ToImage": None, "tore": dict(rtol=RTOL, atol=TORE_ATOL), "ToTimesurface": dict(rtol=RTOL, atol=1e-30)}
    for name, tr in [("ToVoxelGrid", tt.ToVoxelGrid), ("MixedDensityEventStack", MixedDensityEventStack), ("EventStack", EventStack),
                     ("ToImage", tt.ToImage), ("tore", None), ("ToTimesurface", ToTimesurface)]:
        g = G[name]
        H, W = int(g["H"]), int(g["W"])
        data = ev_of(g)
        out = get_item_transform(data, name if tr is None else str(tr), tr, H, W, len(data), None)
        assert out.shape == g["out"].shape and out.dtype == g["out"].dtype, (name, out.shape, out.dtype, g["out"].dtype)
        if tol[name] is None:
            assert np.array_equal(out, g["out"]), name
        else:
            assert_close(out, g["out"], what=name, **tol[name])
        if name in ("EventStack", "ToImage", "ToTimesurface"):  # the reference rewrites the caller's polarity field to {0,1}
            assert set(np.unique(data["p"])) <= {0, 1}
        if name == "ToVoxelGrid":  # tonic rewrites 0 -> -1 in place
            assert set(np.unique(data["p"])) <= {-1, 1}
        out4 = gen4_transforms.get_item_transform(ev_of(g), name if tr is None else str(tr), tr, H, W, len(data))
        if name == "ToVoxelGrid":  # float L2 reductions: order dependent in the last ulp (voxel grids)
            assert_close(out4, out, rtol=1e-6, atol=1e-4)
        else:
            assert np.array_equal(out4, out, equal_nan=True)


ES = golden("eventstack_*")


@pytest.mark.parametrize("name,path", ES, ids=ids(ES))
def test_event_stack_class(name, path):
    from event_representation_study_b200.representations.event_stack import EventStack
    g = load(path)
    H, W = int(g["H"]), int(g["W"])
    d = ev_of(g)
    d["p"] = (d["p"] + 1) // 2
    tr = EventStack(12, len(d), H, W)
    pre = tr.pre_stack(d, d[-1]["t"])
    post = tr.post_stack(pre)
    assert post.shape == (H, W, 1, 12) and post.dtype == np.float32
    assert np.array_equal(post.transpose(0, 1, 3, 2)[..., 0], g["out"])
    # the sparse encoding round-trips through post_stack like the reference's own dicts do
    sparse = [dict(pre[0])]
    assert set(sparse[0]) == set(EventStack.STACK_LIST) and len(sparse[0]["index"]) == 12
    assert np.array_equal(tr.post_stack(sparse), post)


def test_event_stack_future_branch():
    """pre_stack with last_timestamp in the middle of the stream (event_stack.py:28-41) against the oracle semantics."""
    from event_representation_study_b200.representations.event_stack import EventStack
    from oracle import representations as orep
    g = load(golden("eventstack_n2000_pm1")[0][1])
    H, W = int(g["H"]), int(g["W"])
    d = ev_of(g)
    d["p"] = (d["p"] + 1) // 2
    t_mid = d["t"][1200]
    post = EventStack(12, len(d), H, W).post_stack(EventStack(12, len(d), H, W).pre_stack(d, t_mid))
    assert post.shape == (H, W, 2, 12)
    past = d["t"] <= t_mid
    fut = ~past
    want_p = orep.event_stack(d["x"][past], d["y"][past], d["t"][past], d["p"][past], H, W, 12)
    want_f = orep.event_stack(d["x"][fut][::-1], d["y"][fut][::-1], np.arange(fut.sum()), 1 - d["p"][fut][::-1], H, W, 12)[:, :, ::-1]
    assert np.array_equal(post[:, :, 0], want_p) and np.array_equal(post[:, :, 1], want_f)


TS = golden("timesurface_*")


@pytest.mark.parametrize("name,path", TS, ids=ids(TS))
def test_time_surface_class(name, path):
    from event_representation_study_b200.representations.time_surface import ToTimesurface, to_timesurface_numpy
    g = load(path)
    H, W = int(g["H"]), int(g["W"])
    d = ev_of(g)
    d["p"] = ((d["p"] + 1) / 2).astype(np.int8)
    out = ToTimesurface(sensor_size=(W, H, 2), surface_dimensions=None, tau=50000, decay="exp")(d, g["indices"])
    assert out.shape == g["out"].shape and out.dtype == np.float64
    assert_close(out, g["out"], rtol=RTOL, atol=1e-30, what=name)
    mem = np.zeros((2, H, W)) - (50000 * 3 + 1)
    surf = np.zeros_like(out)
    to_timesurface_numpy(d["x"], d["y"], d["t"], d["p"], g["indices"], mem, surf, tau=50000)
    assert np.array_equal(surf, out)
    assert mem.max() <= d["t"].max() and (mem > -(50000 * 3 + 1)).any()


TG = golden("tore_gen1_*")


@pytest.mark.parametrize("name,path", TG, ids=ids(TG))
def test_tore_function_gen1_call(name, path):
    """The gen1_transforms.py:51-67 call: 1-based coordinates from the data minimum, data-dependent frame size."""
    from event_representation_study_b200.representations.tore import events2ToreFeature
    g = load(path)
    d = ev_of(g)
    x1, y1 = d["x"] - min(d["x"]) + 1, d["y"] - min(d["y"]) + 1
    out = events2ToreFeature(x1, y1, d["t"], d["p"], d["t"][-1], 6, (max(y1), max(x1)))
    assert out.shape == g["out"].shape and out.dtype == np.float32
    assert_close(out, g["out"], rtol=RTOL, atol=1e-6, what=name)


def test_tore_zero_coordinate_wraps():
    """imagenet.py:1080-1107 passes 0-based x, y: pixel -1 wraps to the last row / column like numpy."""
    from event_representation_study_b200.representations.tore import events2ToreFeature
    from oracle import representations as orep
    g = load(golden("tore_fixed_n2000_pm1")[0][1])
    H, W = int(g["H"]), int(g["W"])
    t = g["t"].astype(np.int32)
    x, y, p = g["x"].astype(np.int32), g["y"].astype(np.int32), g["p"].astype(np.int32)
    out = events2ToreFeature(x, y, t, p, t[-1], 4, (H, W))
    assert_close(out, orep.tore(x, y, t, p, t[-1], 4, (H, W)), rtol=RTOL, atol=1e-6)
    with pytest.raises(IndexError):
        events2ToreFeature(x + W, y, t, p, t[-1], 4, (H, W))


def test_optimized_representation_f8_seconds():
    from event_representation_study_b200.representations.optimized_representation import N_CHANNELS, get_optimized_representation
    g = load(golden("ergo12_f8_seconds")[0][1])
    s8 = structured({"x": g["x"], "y": g["y"], "t": g["t_seconds"], "p": g["p"]}, "<f8")
    out = get_optimized_representation(s8, len(s8), 30, 40)
    assert N_CHANNELS == 12 and out.shape == (30, 40, 12) and out.dtype == np.float64
    assert_close(out, g["out"], rtol=RTOL, atol=2e-7)
    with pytest.raises(ValueError):
        get_optimized_representation(s8[:0], 0, 30, 40)  # the reference raises on an empty window (t.min())


MD = golden("mdes_S*_n2000_pm1") + golden("mdes_S*_n7_pm1")


@pytest.mark.parametrize("name,path", MD, ids=ids(MD))
def test_mixed_density_class(name, path):
    from event_representation_study_b200.representations.representation_search.mixed_density_event_stack import MixedDensityEventStack
    g = load(path)
    H, W = int(g["H"]), int(g["W"])
    spec = (g["win"].tolist(), g["func"].tolist(), g["agg"].tolist())
    tr = MixedDensityEventStack(len(spec[0]), len(g["x"]), H, W, spec, str(g["stacking"]))
    out = tr.stack(ev_of(g))
    assert out.shape == g["out"].shape and out.dtype == np.float64
    assert_close(out, g["out"], rtol=RTOL, atol=2e-7, what=name)
    stacked = tr.make_stack(g["x"].astype(np.int32), g["y"].astype(np.int32), g["p"].astype(np.int32), g["t"].astype(np.int64))
    assert len(stacked) == len(spec[0])
    assert next(iter(stacked[0])) == "_".join([spec[1][0].capitalize(), spec[2][0].capitalize()])
    wins = tr.create_windows(g["x"], g["y"], g["p"], (g["t"] - g["t"].min()) / max(1, g["t"].max() - g["t"].min()))
    assert len(wins) == (7 if str(g["stacking"]) == "SBN" else 8)


def test_mixed_density_record_fast_path_equals_the_general_path():
    """`<i4` records go to the GPU in one copy and are split there (_single.one_window_structured); every other layout takes
    the host path: same values, same exceptions, zeros for an out-of-range pixel (reference :120-127)"""
    from event_representation_study_b200.representations.optimized_representation import get_optimized_representation
    from event_representation_study_b200 import _single
    from event_representation_study_b200.synth import poisson_window
    H, W, N = 240, 304, 120_000  # 5.5 MB of float64 output: the pinned read-back path
    w = poisson_window(77, N, H, W)
    fast, slow = structured(w, "<i4"), structured(w, "<i8")
    assert _single.one_window_structured(fast, H, W) is not None and _single.one_window_structured(slow, H, W) is None
    a, b = get_optimized_representation(fast, N, H, W), get_optimized_representation(slow, N, H, W)
    assert a.dtype == np.float64 and a.flags.writeable and np.array_equal(a, b)
    assert np.array_equal(get_optimized_representation(fast, N, H, W, _scale=255.0), b * 255)
    bad = fast.copy()
    bad["x"][5] = W
    assert not get_optimized_representation(bad, N, H, W).any()
    bad = fast.copy()
    bad["p"][5] = 2
    with pytest.raises(ValueError):
        get_optimized_representation(bad, N, H, W)
    bad = fast.copy()
    bad["t"][-1] = bad["t"][0] + 2**30
    with pytest.raises(ValueError):
        get_optimized_representation(bad, N, H, W)


def test_operations_exec():
    from event_representation_study_b200.representations.representation_search.operations import Operations
    from oracle import representations as orep
    rng = np.random.default_rng(3)
    H, W, n = 20, 30, 1500
    ev = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), np.sort(rng.random(n)) * 0.4 + 0.3, rng.integers(-1, 2, n)], 1).astype(np.float64)
    for f in orep.FUNCTIONS:
        for a in orep.AGGREGATIONS:
            out = Operations(f, a, H, W)(ev)
            want = orep._operation(ev[:, 0].astype(int), ev[:, 1].astype(int), ev[:, 2], ev[:, 3], f, a, H, W)
            assert out.shape == (H, W) and out.dtype == np.float64
            assert_close(out, want, rtol=RTOL, atol=5e-7, what=f"{f}/{a}")
    with pytest.raises(UnboundLocalError):
        Operations("nope", "sum", H, W)(ev)


VE = golden("voxel_evlicious_*")


@pytest.mark.parametrize("name,path", VE, ids=ids(VE))
def test_evlicious_operator(name, path):
    import torch
    from event_representation_study_b200.evlicious import Events, tools
    g = load(path)
    E = Events(g["x"].copy(), g["y"].copy(), g["t"].copy(), g["p"].copy(), int(g["W"]), int(g["H"]))
    out = tools.events_to_voxel_grid(E, int(g["bins"]), normalize=bool(g["normalize"]))
    assert out.shape == g["out"].shape and out.dtype == np.float32
    if bool(g["normalize"]):
        assert_close(out, g["out"], rtol=RTOL, atol=1e-6, what=name)
    else:
        assert np.array_equal(out, g["out"])
    out_c = tools.events_to_voxel_grid_cuda(E, int(g["bins"]), normalize=bool(g["normalize"]), device="cuda:0")
    assert isinstance(out_c, torch.Tensor) and out_c.is_cuda and np.array_equal(out_c.cpu().numpy(), out)
    assert not tools.events_to_voxel_grid(Events(g["x"][:1].copy(), g["y"][:1].copy(), g["t"][:1].copy(), g["p"][:1].copy(), int(g["W"]), int(g["H"])), 5).any()


def test_events_type_asserts():
    from event_representation_study_b200.evlicious import Events
    x = np.array([1, 2], np.uint16)
    with pytest.raises(AssertionError):
        Events(x, x, np.array([0, 1], np.int32), np.array([1, 1], np.int8), 10, 10)  # t must be int64
    e = Events(x, x.copy(), np.array([0, 1], np.int64), np.array([0, 1], np.int8), 10, 10)
    assert list(e.p) == [-1, 1] and len(e) == 2 and e.to_array().shape == (2, 4)


def test_otmi_and_OTMI():
    from event_representation_study_b200.representations.representation_search.compute_otmi import OTMI, compute_kernel, otmi
    import torch
    g = load(golden("otmi_small")[0][1])
    ev = torch.tensor(np.stack([g["x"], g["y"], g["t"], g["p"]], 1).astype(np.int32))
    c = otmi(ev, g["rep"].astype(np.float64), int(g["H"]), int(g["W"]), int(g["rep_size"]))
    assert isinstance(c, float)
    assert_close(c, g["out"], rtol=RTOL, what="otmi")
    for _, p in golden("gwd_a_pair_*"):
        q = load(p)
        T, cost = OTMI(q["Xs"], q["Xt"], h=float(q["h"])).solve()
        assert T.shape == (len(q["Xs"]), len(q["Xt"])) and abs(T.sum() - 1) < 1e-12
        assert_close(cost, q["out"], rtol=RTOL, what="OTMI.solve")
    C = np.abs(np.subtract.outer(np.arange(5.0), np.arange(5.0)))
    Kx, Ky = compute_kernel(C, 2 * C, 0.7)
    assert_close(Kx, Ky, rtol=1e-12)  # the bandwidth is relative to the matrix' own scale


def _otmi_case(seed, N, H, W, S, C):
    rng = np.random.default_rng(seed)
    ev = np.stack([rng.integers(0, W, N), rng.integers(0, H, N), np.sort(rng.integers(0, 300000, N)), rng.choice([-1, 1], N)], 1).astype(np.int32)
    ev[: N // 3, 0] = rng.integers(0, W // 2 - 1, N // 3)  # one quadrant clearly the densest
    ev[: N // 3, 1] = rng.integers(0, H // 2 - 1, N // 3)
    rep = rng.random((S, S, C)) * 255.0
    rep[rng.random((S, S)) < 0.4] = 0.0  # empty pixels are dropped from Xt
    return ev, rep


@pytest.mark.parametrize("seed,N,H,W,S,C", [(1, 3000, 60, 80, 64, 3), (2, 50000, 240, 304, 240, 12), (3, 700, 33, 47, 31, 1)])
def test_otmi_prepare_matches_the_oracle_pairs(seed, N, H, W, S, C):
    """evrep_otmi_prepare against oracle/gwd.py::otmi_pairs (itself pinned on compute_otmi.py run unmodified, fixture otmi_small):
    the reference's own call shape - a torch int32 tensor, so float32 quotients - bit exact, rows in stream / row-major order"""
    import torch
    import event_representation_study_b200.batched as eb
    from oracle import gwd as ogwd
    ev, rep = _otmi_case(seed, N, H, W, S, C)
    want = ogwd.otmi_pairs(ev, rep, H, W, S)
    pairs, info = eb.otmi_prepare(torch.tensor(ev), rep, H, W, S)
    assert len(pairs) == 3 == len(want) and info["dropped"] == 0 and sum(info["events_per_quadrant"]) <= N
    for (xs, xt), (ws, wt) in zip(pairs, want):
        assert xs.dtype == torch.float64 and xs.is_cuda and tuple(xs.shape) == ws.shape and tuple(xt.shape) == wt.shape
        assert np.array_equal(xs.cpu().numpy(), ws.astype(np.float64), equal_nan=True)
        assert np.array_equal(xt.cpu().numpy(), wt.astype(np.float64), equal_nan=True)


def test_otmi_prepare_float64_events_and_errors():
    """float64 events divide in float64 (numpy semantics of compute_otmi.py:164-169); an empty quadrant raises like min() does"""
    import torch
    import event_representation_study_b200.batched as eb
    ev, rep = _otmi_case(5, 4000, 60, 80, 64, 2)
    H, W, S = 60, 80, 64
    pairs, info = eb.otmi_prepare(ev.astype(np.float64), rep, H, W, S)
    e = ev.astype(np.float64)
    X, Y = e[:, 0], e[:, 1]
    w2, h2 = W / 2 - 1, H / 2 - 1
    quads = [e[(X >= 0) & (X <= w2) & (Y >= 0) & (Y <= h2)], e[(X > w2) & (X <= W - 1) & (Y >= 0) & (Y <= h2)],
             e[(X >= 0) & (X <= w2) & (Y > h2) & (Y <= H - 1)], e[(X > w2) & (X <= W - 1) & (Y > h2) & (Y <= H - 1)]]
    kept = [q for q in range(4) if q != info["dropped"]]
    assert info["dropped"] == int(np.argmax([len(q) for q in quads]))
    for (xs, _), qi in zip(pairs, kept):
        q = quads[qi].copy()
        if qi:
            q[:, 0] -= q[:, 0].min()
            q[:, 1] -= q[:, 1].min()
        t, p = q[:, 2], q[:, 3]
        m = (q[:, 0] < (W - 1) // 2) & (q[:, 1] < (H - 1) // 2)
        want = np.stack([q[:, 0] / ((W - 1) // 2), q[:, 1] / ((H - 1) // 2), (t - t[0]) / (t[-1] - t[0]), (p - p.min()) / (p.max() - p.min())], -1)[m]
        assert np.array_equal(xs.cpu().numpy(), want)
    lone = ev[(ev[:, 0] <= W / 2 - 1) | (ev[:, 1] <= H / 2 - 1)]  # nothing in the fourth quadrant
    with pytest.raises(ValueError):
        eb.otmi_prepare(torch.tensor(lone), rep, H, W, S)


def test_compute_repr():
    from event_representation_study_b200.representations.representation_search.gromov_wasserstein import compute_repr
    g = load(golden("voxel_gwd_small")[0][1])
    out = compute_repr(g["x"].astype(int), g["y"].astype(int), g["t01"], g["p"].astype(float), 40, 30, bins=5)
    assert out.shape == g["out"].shape and out.dtype == np.float64
    assert_close(out, g["out"], rtol=RTOL, atol=2e-6)


def test_tonic_compat_shapes():
    from event_representation_study_b200 import tonic_compat as tt
    g = load(golden("voxel_tonic_n2000_01")[0][1])
    d = ev_of(g)
    out = tt.ToVoxelGrid((int(g["W"]), int(g["H"]), 2), n_time_bins=12)(d)
    assert out.shape == g["out"].shape and out.dtype == np.float64
    assert_close(out, g["out"], rtol=RTOL, atol=2e-6)
    assert "ToVoxelGrid" in str(tt.ToVoxelGrid) and "ToImage" in str(tt.ToImage)


@pytest.mark.parametrize("name,path", golden("voxel_subpixel_*"), ids=[n for n, _ in golden("voxel_subpixel_*")])
def test_events_to_voxel_grid_subpixel(name, path):
    """ev-licious Events with divider > 1 (float32 sub-pixel coordinates): the 4-tap bilinear scatter of
    evlicious/tools/utils.py:93-103 against fixtures made by the reference's own events_to_voxel_grid.  float32 sums of
    signed bilinear weights: 1e-5 relative + 2e-6 absolute per accumulated event weight."""
    from event_representation_study_b200.evlicious.io.utils.events import Events
    from event_representation_study_b200.evlicious.tools.utils import events_to_voxel_grid, events_to_voxel_grid_cuda
    g = load(path)
    E = Events(g["x"].copy(), g["y"].copy(), g["t"].copy(), g["p"].copy(), int(g["W"]), int(g["H"]), divider=int(g["divider"]))
    out = events_to_voxel_grid(E, int(g["bins"]), normalize=bool(g["normalize"]))
    assert out.dtype == np.float32 and out.shape == g["out"].shape
    assert_close(out, g["out"], rtol=RTOL, atol=2e-5, what=name)
    out_t = events_to_voxel_grid_cuda(E, int(g["bins"]), normalize=bool(g["normalize"]))
    assert_close(out_t.cpu().numpy(), g["out"], rtol=RTOL, atol=2e-5, what=name + " (cuda entry point)")
