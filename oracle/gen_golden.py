"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE FILES UNMODIFIED.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container, where /root/reference exists:

    python oracle/gen_golden.py            # rewrites tests/golden/

The reference files are imported from /root/reference (never copied).  Three third-party imports that
are not installable offline are satisfied by oracle/ref_shims (torch_scatter, tonic, ot).  ev-licious'
`tools/utils.py` is loaded by file path with a stub `evlicious` package that exposes the real
`io/utils/events.py::Events` semantics (its package __init__ pulls h5py / matplotlib, absent here).

Each fixture stores the inputs (x, y, t, p, sizes, parameters) and the reference output, so the GPU box
needs neither /root/reference nor this script.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("EVREP_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

sys.path[:] = [q for q in sys.path if os.path.abspath(q or ".") != HERE]  # oracle/representations.py must not shadow the reference package
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "ref_shims"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "representations"))

from event_representation_study_b200.synth import poisson_window, structured  # noqa: E402


def _load_evlicious_utils():
    """ev-licious/src/evlicious/tools/utils.py with a stub parent package."""
    pkg = types.ModuleType("evlicious")

    class Events:  # behaviourally what io/utils/events.py:11-45 provides to tools/utils.py
        def __init__(self, x, y, t, p, width, height, divider=1):
            self._x, self._y, self.t, self.p = x, y, t, p
            self.width, self.height, self.divider = width, height, divider
            if self._x.size > 0:
                self.p[self.p == 0] = -1

        @property
        def x(self):
            return self._x.astype("float32") / self.divider if self.divider > 1 else self._x

        @property
        def y(self):
            return self._y.astype("float32") / self.divider if self.divider > 1 else self._y

        def __len__(self):
            return len(self.x)

    pkg.Events = Events
    sys.modules["evlicious"] = pkg
    spec = importlib.util.spec_from_file_location(
        "evlicious_tools_utils", os.path.join(REF, "ev-licious/src/evlicious/tools/utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, Events


def save(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in kw.items() if k == "out"})


def streams():
    """(tag, H, W, event dict) - the edge cases SURVEY.md 8c lists."""
    H, W = 30, 40
    cases = []
    for n in [1, 2, 3, 7, 8, 9, 100, 2000]:
        cases.append((f"n{n}_pm1", H, W, poisson_window(100 + n, n, H, W, "pm1")))
    cases.append(("n2000_01", H, W, poisson_window(7, 2000, H, W, "01")))
    cases.append(("n3000_clustered", H, W, poisson_window(8, 3000, H, W, "pm1", clustered=True)))
    e = poisson_window(9, 500, H, W, "pm1")
    e["p"][:] = 1
    cases.append(("n500_allpos", H, W, e))
    e = poisson_window(10, 600, H, W, "pm1")
    e["t"] = (e["t"] // 20000) * 20000  # heavy ties, including at t[-1]
    cases.append(("n600_ties", H, W, e))
    e = poisson_window(11, 50, H, W, "pm1")
    e["t"][:] = 1234  # t.max() == t.min()
    cases.append(("n50_tconst", H, W, e))
    e = poisson_window(12, 4000, 24, 64, "pm1")
    e["x"][:] = e["x"] % 3  # duplicate-pixel heavy
    cases.append(("n4000_dups", 24, 64, e))
    e = poisson_window(13, 1500, H, W, "pm1")
    e["t"] += 1_700_000_000  # large absolute timestamps (not rebased)
    cases.append(("n1500_abs_t", H, W, e))
    return cases


def main():
    import torch  # noqa: F401  (operations.py needs it)
    from representations.event_stack import EventStack
    from representations.time_surface import ToTimesurface
    from representations.tore import events2ToreFeature
    from representations.optimized_representation import get_optimized_representation
    from representations.representation_search.mixed_density_event_stack import MixedDensityEventStack
    from representations.gen1_transforms import get_item_transform
    import tonic.transforms as tt
    from representations.representation_search import compute_otmi as ref_otmi
    from representations.representation_search import gromov_wasserstein as ref_gw

    rng = np.random.default_rng(2024)
    FUNCS = ["timestamp", "polarity", "count", "timestamp_pos", "timestamp_neg", "count_pos", "count_neg"]
    AGGS = ["sum", "mean", "max", "variance"]

    with np.errstate(all="ignore"):
        for tag, H, W, ev in streams():
            n = len(ev["x"])
            s4 = structured(ev, "<i4")
            base = dict(x=ev["x"], y=ev["y"], t=ev["t"], p=ev["p"], H=H, W=W)
            # ERGO-12 (v2)
            save(f"ergo12_{tag}", out=get_optimized_representation(s4.copy(), n, H, W), **base)
            # generic MixedDensityEventStack, random spec, both stacking types
            for st in ["SBN", "SBT"]:
                C = 9
                wi = rng.integers(0, 7 if st == "SBN" else 8, C).tolist()
                fu = [FUNCS[i] for i in rng.integers(0, 7, C)]
                ag = [AGGS[i] for i in rng.integers(0, 4, C)]
                out = MixedDensityEventStack(C, n, H, W, (wi, fu, ag), st).stack(s4.copy())
                save(f"mdes_{st}_{tag}", out=out, win=np.array(wi), func=np.array(fu), agg=np.array(ag), stacking=st, **base)
            # EventStack, gen1_transforms.py:33-42 call (without the *255)
            d = s4.copy()
            d["p"] = (d["p"] + 1) // 2
            tr = EventStack(12, n, H, W)
            out = tr.post_stack(tr.pre_stack(d, d[-1]["t"])).transpose(0, 1, 3, 2)[..., 0]
            save(f"eventstack_{tag}", out=out, **base)
            # TimeSurface, gen1_transforms.py:69-87 call (without the *255)
            if n >= 2:
                d = s4.copy()
                d["p"] = ((d["p"] + 1) / 2).astype(np.int8)
                t = d["t"]
                idx = np.searchsorted((t - t[0]) / (t[-1] - t[0]) * 6, np.arange(6) + 1)
                out = ToTimesurface(sensor_size=(W, H, 2), surface_dimensions=None, tau=50000, decay="exp")(d, idx)
                save(f"timesurface_{tag}", out=out, indices=idx, **base)
            # TORE: gen1 call (1-based, data-dependent frame) and fixed-frame call (imagenet.py:1080-1107 style)
            d = s4.copy()
            x1, y1 = d["x"] - min(d["x"]) + 1, d["y"] - min(d["y"]) + 1
            out = events2ToreFeature(x1, y1, d["t"], d["p"], d["t"][-1], 6, (max(y1), max(x1)))
            save(f"tore_gen1_{tag}", out=out, **base)
            out = events2ToreFeature(d["x"] + 1, d["y"] + 1, d["t"], d["p"], d["t"][-1], 4, (H, W))
            save(f"tore_fixed_{tag}", out=out, k=4, **base)
            # tonic voxel grid (shim == restatement; dispatch shape only) and 2-channel histogram
            if n >= 2:
                out = tt.ToVoxelGrid((W, H, 2), n_time_bins=12)(s4.copy())
                save(f"voxel_tonic_{tag}", out=out, **base)

        # get_item_transform dispatch, every branch, one stream, WITH the *255 (gen1_transforms.py:12-89)
        tag, H, W, ev = "dispatch", 30, 40, poisson_window(77, 2500, 30, 40, "pm1")
        base = dict(x=ev["x"], y=ev["y"], t=ev["t"], p=ev["p"], H=H, W=W)

        class MixedDensityEventStackName:  # only str(transform) matters for dispatch
            pass

        for name, tr in [("ToVoxelGrid", tt.ToVoxelGrid), ("MixedDensityEventStack", MixedDensityEventStack),
                         ("EventStack", EventStack), ("ToImage", tt.ToImage), ("tore", None), ("ToTimesurface", ToTimesurface)]:
            out = get_item_transform(structured(ev, "<i4"), name, tr, H, W, 2500, None)
            save(f"dispatch_{name}", out=np.asarray(out), **base)

        # n_imagenet-style f8 structured input with t in SECONDS (imagenet.py:1002-1006): int64 truncation quirk
        ev = poisson_window(78, 1200, 30, 40, "pm1")
        s8 = structured(ev, "<f8")
        s8["t"] = ev["t"] * 1e-6 + 3.0
        save("ergo12_f8_seconds", out=get_optimized_representation(s8.copy(), 1200, 30, 40), x=ev["x"], y=ev["y"],
             t_seconds=s8["t"], p=ev["p"], H=30, W=40)

        # Gen1-sized ERGO-12, float32 storage
        ev = poisson_window(1, 50_000, 240, 304, "pm1")
        out = get_optimized_representation(structured(ev, "<i4"), 50_000, 240, 304)
        save("ergo12_gen1_50k", out=out.astype(np.float32), x=ev["x"], y=ev["y"], t=ev["t"], p=ev["p"], H=240, W=304)

        # ev-licious voxel grid (real reference code) - BASELINE config 1 shape and a small one
        utils, Events = _load_evlicious_utils()
        for tag, H, W, n, bins in [("small", 30, 40, 3000, 5), ("gen1_50k", 240, 304, 50_000, 5)]:
            ev = poisson_window(200 + n, n, H, W, "pm1")
            for norm in [False, True]:
                E = Events(ev["x"].copy(), ev["y"].copy(), ev["t"].copy(), ev["p"].copy(), W, H)
                out = utils.events_to_voxel_grid(E, bins, normalize=norm)
                save(f"voxel_evlicious_{tag}_norm{int(norm)}", out=out, x=ev["x"], y=ev["y"], t=ev["t"], p=ev["p"], H=H, W=W,
                     bins=bins, normalize=norm)

        # compute_repr (gromov_wasserstein.py:72-82) and compute_kernel
        ev = poisson_window(300, 3000, 30, 40, "pm1")
        t01 = ev["t"] / ev["t"][-1]
        out = ref_gw.compute_repr(ev["x"].astype(int), ev["y"].astype(int), t01, ev["p"].astype(float), 40, 30, bins=5)
        save("voxel_gwd_small", out=out, x=ev["x"], y=ev["y"], t01=t01, p=ev["p"], H=30, W=40, bins=5)

        # GWD-A: OTMI.solve and otmi() (compute_otmi.py) through the ot shim
        for tag, n, m, dt in [("a", 300, 200, 14), ("b", 150, 400, 7), ("c", 256, 256, 14)]:
            Xs = rng.random((n, 4))
            Xt = np.concatenate([rng.random((m, dt - 2)) * 255, rng.random((m, 2))], axis=1)
            _, cost = ref_otmi.OTMI(Xs, Xt, h=0.7).solve()
            save(f"gwd_a_pair_{tag}", out=np.float64(cost), Xs=Xs, Xt=Xt, h=0.7)
        H, W, S = 48, 64, 64
        ev = poisson_window(400, 4000, H, W, "pm1", clustered=True)
        rep = get_optimized_representation(structured(ev, "<i4"), 4000, H, W) * 255
        sq = np.full((S, S, 12), 114.0)
        sq[(S - H) // 2:(S - H) // 2 + H, :, :] = rep  # letterbox to S x S, pad value 114
        evt = torch.tensor(np.stack([ev["x"], ev["y"], ev["t"], ev["p"]], 1).astype(np.int32))
        cost = ref_otmi.otmi(evt.clone(), sq, H, W, S)
        save("otmi_small", out=np.float64(cost), x=ev["x"], y=ev["y"], t=ev["t"], p=ev["p"], H=H, W=W, rep=sq.astype(np.float32), rep_size=S)


if __name__ == "__main__":
    main()
