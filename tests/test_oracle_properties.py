"""Size-independent properties of the CPU oracle (test infrastructure checking itself beyond the golden fixtures): relations
that must hold for ANY stream, tried on hypothesis-drawn windows.  The same relations are what the GPU tests use at the
BASELINE sizes, where the oracle is too slow to be the checker (tests/test_gpu_parity.py)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import filters as ofil
from oracle import representations as orep

H, W = 12, 16


@st.composite
def windows(draw, min_n=1, max_n=300, zero_polarity=False):
    n = draw(st.integers(min_n, max_n))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    x = rng.integers(0, W, n).astype(np.uint16)
    y = rng.integers(0, H, n).astype(np.uint16)
    t = np.cumsum(rng.integers(0, draw(st.sampled_from([1, 2, 50, 5000])) + 1, n)).astype(np.int64) + draw(st.sampled_from([0, 10**9]))
    p = rng.choice(np.array([-1, 0, 1] if zero_polarity else [-1, 1], np.int8), n)
    return {"x": x, "y": y, "t": t, "p": p}


@settings(max_examples=40, deadline=None, derandomize=True)
@given(windows(zero_polarity=True), st.integers(0, 6))
def test_mixed_density_relations(w, win):
    fu = ["count", "count_pos", "count_neg", "polarity", "timestamp", "timestamp", "timestamp", "timestamp", "count"]
    ag = ["sum", "sum", "sum", "sum", "min", "mean", "max", "variance", "max"]
    with np.errstate(all="ignore"):
        out = orep.mixed_density_event_stack(w["x"], w["y"], w["t"], w["p"], H, W, [win] * len(fu), fu, ag, "SBN")
    lo, hi = orep.sbn_window_bounds(len(w["x"]))[win]
    cnt, cpos, cneg, pol, tmin, tmean, tmax, tvar, touched = (out[:, :, c] for c in range(9))
    assert np.all(cnt == np.round(cnt)) and np.all(cpos + cneg <= cnt)           # classes partition (p == 0 may belong to neither)
    assert np.array_equal(touched, (cnt > 0).astype(np.float64))                 # max of ones = "touched"
    if np.isfinite(tmean).all():                                                 # a constant-time window gives NaN everywhere
        on = cnt > 0
        assert np.all(tmin[on] <= tmean[on] + 1e-12) and np.all(tmean[on] <= tmax[on] + 1e-12)
        assert np.all(tvar[on] >= -1e-9) and np.all(tvar[cnt == 1] == 0)
        assert np.all(tmin[~on] == 0) and np.all(tmax[~on] == 0)                 # torch_scatter leaves untouched pixels at 0
    assert np.all(np.abs(pol) <= cnt) and cnt.sum() == hi - lo                   # every event of the index window is counted once
    assert pol.sum() == int(w["p"][lo:hi].astype(np.int64).sum())                # the polarity sum sees the raw values


@settings(max_examples=25, deadline=None, derandomize=True)
@given(windows(min_n=2))
def test_ergo12_count_channels_are_consistent(w):
    with np.errstate(all="ignore"):
        out = orep.ergo12(w["x"], w["y"], w["t"], w["p"], H, W)
    assert out.shape == (H, W, 12)
    # v2 channel 5 = (window 6, count, sum) and channel 3 = (window 6, polarity, sum): same events
    assert np.all(np.abs(out[:, :, 3]) <= out[:, :, 5]) and np.all((out[:, :, 3] - out[:, :, 5]) % 2 == 0)
    n6 = len(w["x"]) - (len(w["x"]) // 2 + len(w["x"]) // 4 + len(w["x"]) // 8)
    assert out[:, :, 5].sum() == n6


@settings(max_examples=25, deadline=None, derandomize=True)
@given(windows(), st.integers(1, 12))
def test_event_stack_values_and_nesting(w, k):
    p01 = (w["p"].astype(np.int32) + 1) // 2
    es = orep.event_stack(w["x"], w["y"], w["t"], p01, H, W, k)
    assert es.shape == (H, W, k) and set(np.unique(es)).issubset({-1.0, 0.0, 1.0})
    # slice j + 1 looks at a suffix of slice j's events: a pixel set there is set in every earlier slice
    for j in range(k - 1):
        assert np.all((es[:, :, j + 1] != 0) <= (es[:, :, j] != 0))
    assert np.array_equal(es[:, :, 0] != 0, orep.to_image(w["x"], w["y"], p01, H, W).sum(0) > 0)  # slice 0 = every event


@settings(max_examples=25, deadline=None, derandomize=True)
@given(windows(), st.integers(1, 3))
def test_background_activity_limits(w, radius):
    n = len(w["x"])
    t = w["t"] + 1  # strictly positive stamps: "t_last > 0" is then "the block was written before"
    keep_all = ofil.background_activity_filter(np.ones(n, bool), np.full((H, W), -np.inf), w["x"], w["y"], t, np.inf, radius)
    assert keep_all.all()                                                        # nothing is ever "too old"
    ts = np.full((H, W), -np.inf)
    m = ofil.background_activity_filter(np.ones(n, bool), ts, w["x"], w["y"], t, -1.0, radius)
    # depth < 0: an event passes iff no earlier event wrote its pixel; the first event always passes
    assert m[0] and np.isfinite(ts[w["y"][-1], w["x"][-1]]) and ts.max() == t[-1]
    seen = np.zeros((H, W), bool)
    for i in range(n):
        xi, yi = int(w["x"][i]), int(w["y"][i])
        assert m[i] == (not seen[yi, xi])
        seen[max(yi - radius, 0):yi + radius, max(xi - radius, 0):xi + radius] = True


@settings(max_examples=25, deadline=None, derandomize=True)
@given(windows())
def test_refractory_with_zero_period_keeps_everything_and_tracks_the_last_stamp(w):
    n = len(w["x"])
    last = np.full((H, W), -np.inf)
    assert ofil.refractory_period(np.ones(n, bool), w["x"], w["y"], w["t"], 0, last).all()
    want = np.full((H, W), -np.inf)
    np.maximum.at(want, (w["y"], w["x"]), w["t"].astype(np.float64))
    assert np.array_equal(last, want)
