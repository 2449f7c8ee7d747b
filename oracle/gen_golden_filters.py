"""Writes tests/golden/filter_*.npz with the reference's own numba filter kernels (ev-licious/src/evlicious/tools/utils.py,
loaded by path exactly like oracle/gen_golden.py does).  Runs only where /root/reference exists.  Every fixture feeds the
stream in two pieces, like two `insert` calls of a filter object (tools/filters.py:57-109), and stores the state after
each piece.
TEST INFRASTRUCTURE ONLY (fixture generator; nothing in the product package imports it)."""
import os
import sys
import tempfile

import numpy as np

# _filter_events_resize is jitted with cache=True; its module is loaded by path under a synthetic name, and a cache entry
# written by one run cannot be unpickled by the next ("No module named '<dynamic>'"): give every run a fresh cache
os.environ["NUMBA_CACHE_DIR"] = tempfile.mkdtemp(prefix="numba_cache_")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  (module import only: its main() is not run)
from oracle import filters as ofil  # noqa: E402

U, _Events = gg._load_evlicious_utils()


def stream(seed, n, H, W, hot=False):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, W, n)
    y = rng.integers(0, H, n)
    if hot:  # a few very active pixels: long per-pixel runs
        k = rng.random(n) < 0.5
        x[k] = rng.integers(0, 3, k.sum())
        y[k] = rng.integers(0, 2, k.sum())
    t = np.cumsum(rng.integers(0, 40, n)).astype(np.int64)  # ties included
    p = rng.choice(np.array([-1, 1], np.int8), n)
    return x.astype(np.uint16), y.astype(np.uint16), t, p


def main():
    out = os.path.join(ROOT, "tests", "golden")
    for name, seed, n, H, W, hot in [("uniform", 1, 30000, 48, 64, False), ("hot", 2, 30000, 48, 64, True), ("tiny", 3, 9, 6, 6, False)]:
        x, y, t, p = stream(seed, n, H, W, hot)
        cut = n // 3
        d = {"x": x, "y": y, "t": t, "p": p, "H": H, "W": W, "cut": cut}
        # refractory
        last = np.full((H, W), -np.inf)
        masks, states = [], []
        for sl in (slice(0, cut), slice(cut, n)):
            m = np.ones(len(x[sl]), bool)
            masks.append(U._refractory_period(m, x[sl], y[sl], t[sl], 2000, last))
            states.append(last.copy())
        d.update(refr_mask=np.concatenate(masks), refr_state0=states[0], refr_state1=states[1], refr_period=2000)
        chk = ofil.refractory_period(np.ones(n, bool), x, y, t, 2000, np.full((H, W), -np.inf))
        assert np.array_equal(chk, d["refr_mask"]), name
        # contrast threshold
        act = np.zeros((H, W), np.int32)
        masks, states = [], []
        for sl in (slice(0, cut), slice(cut, n)):
            m = np.zeros(len(x[sl]), bool)
            masks.append(U._contrast_threshold_control(act, m, x[sl], y[sl], p[sl], 3))
            states.append(act.copy())
        d.update(ctc_mask=np.concatenate(masks), ctc_state0=states[0], ctc_state1=states[1], ctc_factor=3)
        assert np.array_equal(ofil.contrast_threshold_control(np.zeros((H, W), np.int32), np.zeros(n, bool), x, y, p, 3), d["ctc_mask"]), name
        # resize filter, 2 x 3 cells
        fx, fy = 2, 3
        cm = np.zeros((H // fy, W // fx), np.float32)
        masks, states = [], []
        for sl in (slice(0, cut), slice(cut, n)):
            m = np.zeros(len(x[sl]), bool)
            m, cm = U._filter_events_resize(x[sl], y[sl], p[sl], m, cm, fx, fy)
            masks.append(m)
            states.append(cm.copy())
        d.update(rsz_mask=np.concatenate(masks), rsz_state0=states[0], rsz_state1=states[1], fx=fx, fy=fy)
        m2, _ = ofil.filter_events_resize(x, y, p, np.zeros(n, bool), np.zeros((H // fy, W // fx), np.float32), fx, fy)
        assert np.array_equal(m2, d["rsz_mask"]), name
        # background activity, radius 1 to 4 (timestamps start at -inf like tools/filters.py:64-65)
        for r in (1, 2, 3, 4):
            ts = np.full((H, W), -np.inf)
            masks, states = [], []
            for sl in (slice(0, cut), slice(cut, n)):
                m = np.ones(len(x[sl]), bool)
                masks.append(U._background_activity_filter(m, ts, x[sl], y[sl], t[sl], 300, r).copy())
                states.append(ts.copy())
            d.update({f"ba{r}_mask": np.concatenate(masks), f"ba{r}_state0": states[0], f"ba{r}_state1": states[1], "ba_depth": 300})
            chk = ofil.background_activity_filter(np.ones(n, bool), np.full((H, W), -np.inf), x, y, t, 300, r)
            assert np.array_equal(chk, d[f"ba{r}_mask"]), (name, r)
        np.savez_compressed(os.path.join(out, f"filter_{name}.npz"), **d)
        print(name, n, "kept", int(d["refr_mask"].sum()), int(d["ctc_mask"].sum()), int(d["rsz_mask"].sum()), int(d["ba1_mask"].sum()),
              int(d["ba2_mask"].sum()))


if __name__ == "__main__":
    main()
