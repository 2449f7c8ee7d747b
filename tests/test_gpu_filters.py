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
