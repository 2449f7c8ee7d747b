"""ev-licious' stateful per-pixel filters on the GPU against fixtures made with the reference's own numba kernels
(tests/golden/filter_*.npz: masks and states after feeding a stream in two pieces) and against the plain-Python oracle.
Masks and integer states are exact; the float32 change map of the resize filter is exact too (same float32 / float64
operations in the same per-cell order)."""
import numpy as np
import pytest

from conftest import golden, load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(cuda_device):
    import event_representation_study_b200.batched as eb
    return eb


def _pieces(E, g, lo, hi):
    sl = slice(lo, hi)
    return E.pack_events([{"x": g["x"][sl], "y": g["y"][sl], "t": g["t"][sl], "p": g["p"][sl]}], "cuda")


@pytest.mark.parametrize("name,path", golden("filter_*"), ids=[n for n, _ in golden("filter_*")])
def test_filters_match_reference_numba_kernels(E, name, path):
    g = load(path)
    H, W, n, cut = int(g["H"]), int(g["W"]), len(g["x"]), int(g["cut"])
    fx, fy = int(g["fx"]), int(g["fy"])
    for kind, key, param, shape, kw in [("refractory", "refr", float(g["refr_period"]), (H, W), {}),
                                        ("contrast", "ctc", float(g["ctc_factor"]), (H, W), {}),
                                        ("resize", "rsz", 0.0, (H // fy, W // fx), {"fx": fx, "fy": fy})]:
        st = E.filter_state(kind, 1, shape[0], shape[1])
        masks = []
        for k, (lo, hi) in enumerate(((0, cut), (cut, n))):
            m, st = E.filter_events(_pieces(E, g, lo, hi), shape[0], shape[1], kind, param, st, **kw)
            masks.append(m.cpu().numpy().astype(bool))
            assert np.array_equal(st[0].cpu().numpy(), g[f"{key}_state{k}"]), f"{name} {kind} state after piece {k}"
        assert np.array_equal(np.concatenate(masks), g[f"{key}_mask"]), f"{name} {kind} mask"


def test_filters_batched_long_streams_vs_oracle(E):
    """several streams per call, sizes around the super-chunk boundary, one hot pixel with thousands of events"""
    from oracle import filters as ofil
    H, W = 40, 56
    rng = np.random.default_rng(5)
    wins = []
    for n in (20000, 8193, 1, 50000):
        x = rng.integers(0, W, n).astype(np.uint16)
        y = rng.integers(0, H, n).astype(np.uint16)
        hot = rng.random(n) < 0.3
        x[hot], y[hot] = 7, 9
        wins.append({"x": x, "y": y, "t": np.cumsum(rng.integers(0, 30, n)).astype(np.int64), "p": rng.choice(np.array([-1, 1], np.int8), n)})
    ev = E.pack_events(wins, "cuda")
    offs = ev.offsets
    m_r, s_r = E.filter_events(ev, H, W, "refractory", 500.0)
    m_c, s_c = E.filter_events(ev, H, W, "contrast", 2.0)
    m_z, s_z = E.filter_events(ev, H // 2, W // 2, "resize", 0.0, fx=2, fy=2)
    for b, w in enumerate(wins):
        n = len(w["x"])
        sl = slice(int(offs[b]), int(offs[b + 1]))
        last = np.full((H, W), -np.inf)
        want = ofil.refractory_period(np.ones(n, bool), w["x"], w["y"], w["t"], 500.0, last)
        assert np.array_equal(m_r[sl].cpu().numpy().astype(bool), want) and np.array_equal(s_r[b].cpu().numpy(), last)
        act = np.zeros((H, W), np.int32)
        want = ofil.contrast_threshold_control(act, np.zeros(n, bool), w["x"], w["y"], w["p"], 2.0)
        assert np.array_equal(m_c[sl].cpu().numpy().astype(bool), want) and np.array_equal(s_c[b].cpu().numpy(), act)
        cm = np.zeros((H // 2, W // 2), np.float32)
        want, cm = ofil.filter_events_resize(w["x"], w["y"], w["p"], np.zeros(n, bool), cm, 2, 2)
        assert np.array_equal(m_z[sl].cpu().numpy().astype(bool), want) and np.array_equal(s_z[b].cpu().numpy(), cm)


def test_evlicious_filter_function_mirrors(E):
    """the reference's positional signatures, arrays updated in place (utils.py:143-158, 184-200)"""
    from event_representation_study_b200.evlicious.tools import utils as U
    from oracle import filters as ofil
    g = load(golden("filter_uniform")[0][1])
    H, W = int(g["H"]), int(g["W"])
    x, y, t, p = g["x"][:5000], g["y"][:5000], g["t"][:5000], g["p"][:5000]
    last, last_w = np.full((H, W), -np.inf), np.full((H, W), -np.inf)
    got = U._refractory_period(np.ones(5000, bool), x, y, t, 2000, last)
    assert np.array_equal(got, ofil.refractory_period(np.ones(5000, bool), x, y, t, 2000, last_w)) and np.array_equal(last, last_w)
    act, act_w = np.zeros((H, W), np.int32), np.zeros((H, W), np.int32)
    got = U._contrast_threshold_control(act, np.zeros(5000, bool), x, y, p, 3)
    assert np.array_equal(got, ofil.contrast_threshold_control(act_w, np.zeros(5000, bool), x, y, p, 3)) and np.array_equal(act, act_w)
    cm, cm_w = np.zeros((H // 3, W // 2), np.float32), np.zeros((H // 3, W // 2), np.float32)
    got, cm2 = U._filter_events_resize(x, y, p, np.zeros(5000, bool), cm, 2, 3)
    want, _ = ofil.filter_events_resize(x, y, p, np.zeros(5000, bool), cm_w, 2, 3)
    assert np.array_equal(got, want) and np.array_equal(cm, cm_w) and cm2 is cm


def test_filters_empty_windows_and_out_of_range_events(E):
    """a batch with empty streams; events outside the sensor get mask 0, raise the flag and leave the state alone"""
    import torch
    H, W = 8, 8
    wins = [{"x": np.zeros(0, np.uint16), "y": np.zeros(0, np.uint16), "t": np.zeros(0, np.int64), "p": np.zeros(0, np.int8)},
            {"x": np.array([1, 1, 200, 1], np.uint16), "y": np.array([2, 2, 2, 2], np.uint16), "t": np.array([0, 5, 6, 20], np.int64),
             "p": np.array([1, 1, 1, -1], np.int8)},
            {"x": np.zeros(0, np.uint16), "y": np.zeros(0, np.uint16), "t": np.zeros(0, np.int64), "p": np.zeros(0, np.int8)}]
    ev = E.pack_events(wins, "cuda")
    m, st = E.filter_events(ev, H, W, "refractory", 10.0)
    assert m.cpu().numpy().tolist() == [1, 0, 0, 1]
    assert int(E.window_flags(ev)[1]) & 0x100
    assert torch.isinf(st[0]).all() and torch.isinf(st[2]).all() and float(st[1, 2, 1]) == 20.0
    none = E.pack_events([wins[0]], "cuda")
    m0, _ = E.filter_events(none, H, W, "contrast", 2.0)
    assert m0.numel() == 0


@pytest.mark.parametrize("name,path", golden("filter_*"), ids=[n for n, _ in golden("filter_*")])
@pytest.mark.parametrize("radius", [1, 2])
def test_background_activity_matches_reference_numba_kernel(E, name, path, radius):
    """utils.py:169-178 run by the reference itself (oracle/gen_golden_filters.py), stream fed in two pieces"""
    g = load(path)
    H, W, n, cut = int(g["H"]), int(g["W"]), len(g["x"]), int(g["cut"])
    st = E.filter_state("background", 1, H, W)
    masks = []
    for k, (lo, hi) in enumerate(((0, cut), (cut, n))):
        m, st = E.filter_events(_pieces(E, g, lo, hi), H, W, "background", float(g["ba_depth"]), st, fx=radius)
        masks.append(m.cpu().numpy().astype(bool))
        assert np.array_equal(st[0].cpu().numpy(), g[f"ba{radius}_state{k}"]), f"{name} state after piece {k}"
    assert np.array_equal(np.concatenate(masks), g[f"ba{radius}_mask"]), name


@pytest.mark.parametrize("radius", [1, 3])
def test_background_activity_batched_vs_oracle(E, radius):
    """several streams, events on every border and corner (clipped blocks), a hot pixel, int32 and int64 timestamps"""
    from oracle import filters as ofil
    H, W = 37, 53
    rng = np.random.default_rng(11 + radius)
    wins = []
    for n, t0 in ((20000, 0), (8193, 5_000_000_000), (1, 7), (30000, 100)):
        x = rng.integers(0, W, n).astype(np.uint16)
        y = rng.integers(0, H, n).astype(np.uint16)
        edge = rng.random(n) < 0.2
        x[edge] = rng.choice(np.array([0, W - 1], np.uint16), int(edge.sum()))
        edge = rng.random(n) < 0.2
        y[edge] = rng.choice(np.array([0, H - 1], np.uint16), int(edge.sum()))
        hot = rng.random(n) < 0.2
        x[hot], y[hot] = 5, 6
        wins.append({"x": x, "y": y, "t": t0 + np.cumsum(rng.integers(0, 25, n)).astype(np.int64), "p": np.ones(n, np.int8)})
    for tdt in (np.int64, None):
        use = wins if tdt is not None else [wins[0], wins[2], wins[3]]  # the second stream needs 64-bit stamps
        ev = E.pack_events(use, "cuda") if tdt is None else E.pack_events(use, "cuda", np.int64)
        m, s = E.filter_events(ev, H, W, "background", 150.0, fx=radius)
        offs = ev.offsets
        for b, w in enumerate(use):
            n = len(w["x"])
            ts = np.full((H, W), -np.inf)
            want = ofil.background_activity_filter(np.ones(n, bool), ts, w["x"], w["y"], w["t"], 150.0, radius)
            got = m[int(offs[b]):int(offs[b + 1])].cpu().numpy().astype(bool)
            assert np.array_equal(got, want), f"stream {b}: {int((got != want).sum())} of {n} differ"
            assert np.array_equal(s[b].cpu().numpy(), ts), f"stream {b} state"
        assert not (E.window_flags(ev) & 0x100).any()  # clipped blocks are not out-of-range events


def test_evlicious_filter_objects(E):
    """tools/filters.py:23-129: the filter classes, state kept as numpy arrays across insert() calls"""
    from event_representation_study_b200.evlicious import Events
    from event_representation_study_b200.evlicious.tools import filters as F
    from oracle import filters as ofil
    g = load(golden("filter_hot")[0][1])
    H, W, n = int(g["H"]), int(g["W"]), 12000
    mk = lambda lo, hi: Events(g["x"][lo:hi].copy(), g["y"][lo:hi].copy(), g["t"][lo:hi].copy(), g["p"][lo:hi].copy(), W, H)  # noqa: E731
    ba = F.from_flags(type("Flags", (), {"filter_type": int(F.Filtering_Type.BackgroundActivity), "depth_us": 300, "radius": 1}))
    kept = [ba.insert(mk(0, 5000)), ba.insert(mk(5000, n))]
    ts = np.full((H, W), -np.inf)
    want = ofil.background_activity_filter(np.ones(n, bool), ts, g["x"][:n], g["y"][:n], g["t"][:n], 300, 1)
    assert np.array_equal(np.concatenate([k.t for k in kept]), g["t"][:n][want]) and np.array_equal(ba.timestamps, ts)
    rp = F.RefractoryPeriod(depth_us=2000)
    last = np.full((H, W), -np.inf)
    want = ofil.refractory_period(np.ones(n, bool), g["x"][:n], g["y"][:n], g["t"][:n], 2000, last)
    assert len(rp.insert(mk(0, n))) == int(want.sum()) and np.array_equal(rp.timestamps, last)
    ct = F.ContrastThresholdIncrease(contrast_threshold_multiplier=3)
    act = np.zeros((H, W), np.int32)
    want = ofil.contrast_threshold_control(act, np.zeros(n, bool), g["x"][:n], g["y"][:n], g["p"][:n], 3)
    assert np.array_equal(ct.insert(mk(0, n)).t, g["t"][:n][want]) and np.array_equal(ct.counter_map, act)
    hp = F.HotPixel()
    out = hp.insert(mk(0, n))  # half of this stream sits on six pixels: they are calibrated away
    want_mask = ofil.hot_pixel_mask(g["x"][:n], g["y"][:n], H, W)
    assert np.array_equal(hp.hot_pixel_mask, want_mask) and not want_mask.all()
    assert np.array_equal(out.t, g["t"][:n][want_mask[g["y"][:n], g["x"][:n]]])
    u = load(golden("filter_uniform")[0][1])
    hp2 = F.HotPixel()
    ev_u = Events(u["x"][:n].copy(), u["y"][:n].copy(), u["t"][:n].copy(), u["p"][:n].copy(), W, H)
    assert len(hp2.insert(ev_u)) == n and hp2.hot_pixel_mask.all()  # no pixel stands out: everything passes
    assert F.Filtering_Type.summary().count("=") == 5 and len(F.Random(4).insert(mk(0, 1000))) == 250


def test_background_activity_argument_checks(E):
    g = load(golden("filter_tiny")[0][1])
    ev = _pieces(E, g, 0, len(g["x"]))
    with pytest.raises(ValueError):
        E.filter_events(ev, int(g["H"]), int(g["W"]), "background", 10.0, fx=5)
    m, st = E.filter_events(E.pack_events([{k: g[k][:0] for k in "xytp"}], "cuda"), 6, 6, "background", 10.0, fx=1)
    assert m.numel() == 0 and np.isinf(st.cpu().numpy()).all()


def test_index_keyed_ops_keep_events_beyond_the_time_span_limit(E):
    """a stream spanning more than 2^30 us (18 min): the filters and EventStack key on the stream index, so every event is
    kept and EVREP_WF_T_RANGE is only informative"""
    from oracle import filters as ofil
    from oracle import representations as orep
    H, W, n = 24, 32, 20000
    rng = np.random.default_rng(77)
    w = {"x": rng.integers(0, W, n).astype(np.uint16), "y": rng.integers(0, H, n).astype(np.uint16),
         "t": np.cumsum(rng.integers(0, 300_000, n)).astype(np.int64), "p": rng.choice(np.array([-1, 1], np.int8), n)}
    assert int(w["t"][-1] - w["t"][0]) > 2 * 2**30
    ev = E.pack_events([w], "cuda", np.int64)
    m, s = E.filter_events(ev, H, W, "refractory", 5e5)
    assert int(E.window_flags(ev)[0]) & 0x400
    last = np.full((H, W), -np.inf)
    want = ofil.refractory_period(np.ones(n, bool), w["x"], w["y"], w["t"], 5e5, last)
    assert np.array_equal(m.cpu().numpy().astype(bool), want) and np.array_equal(s[0].cpu().numpy(), last)
    m, s = E.filter_events(ev, H, W, "background", 4e5, fx=1)
    ts = np.full((H, W), -np.inf)
    want = ofil.background_activity_filter(np.ones(n, bool), ts, w["x"], w["y"], w["t"], 4e5, 1)
    assert np.array_equal(m.cpu().numpy().astype(bool), want) and np.array_equal(s[0].cpu().numpy(), ts)
    es = E.event_stack(ev, H, W, 12).cpu().numpy()[0]
    assert np.array_equal(es, orep.event_stack(w["x"], w["y"], w["t"], (w["p"].astype(np.int32) + 1) // 2, H, W, 12))
