"""Writes tests/golden/nimg_*.npz by running the reference's N-ImageNet wrappers
(n_imagenet/real_cnn_model/data/imagenet.py:1002-1134) UNMODIFIED on a small seeded event tensor.  Runs only where
/root/reference exists.  Third-party imports absent offline come from oracle/ref_shims (tonic, torch_scatter - the wrappers
only need the module to import); `np.int` (removed from numpy, used at imagenet.py:1125-1126) is restored as `int`; the
last line of reshape_then_to_image calls `.float()` on a numpy array, so that wrapper is run up to that line by giving
numpy's ndarray subclass a `float` method returning a float32 tensor.
TEST INFRASTRUCTURE ONLY (fixture generator; nothing in the product package imports it)."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:] = [q for q in sys.path if os.path.abspath(q or ".") != HERE]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "ref_shims"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "representations"))
np.int = int  # noqa: removed alias used by the reference

import torch_scatter  # noqa: E402,F401  (the shim: scatter, scatter_max, scatter_min)

spec = importlib.util.spec_from_file_location("ref_imagenet", os.path.join(REF, "n_imagenet", "real_cnn_model", "data", "imagenet.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


class _FloatArray(np.ndarray):
    def float(self):
        return torch.tensor(np.ascontiguousarray(self)).float()


def main():
    H = W = 64
    rng = np.random.default_rng(4242)
    n = 4000
    t = np.sort(rng.random(n) * 0.05) + 1.5                      # seconds, like parse_event delivers them
    ev = np.stack([rng.integers(0, W, n).astype(np.float64), rng.integers(0, H, n).astype(np.float64), t, rng.choice([-1.0, 1.0], n)], 1)
    # the time-based wrappers need microsecond-like stamps to be meaningful (tau = 50000, TORE's 150 us floor)
    ev_us = ev.copy()
    ev_us[:, 2] = np.floor((t - t[0]) * 2e6)
    out = {"events_s": ev, "events_us": ev_us, "H": H, "W": W}
    for name, data in [("voxel_grid", ev), ("optimized", ev_us), ("event_stack", ev), ("tore", ev_us)]:
        rep = getattr(ref, "reshape_then_" + name)(torch.tensor(data.copy()), height=H, width=W)
        out[name] = rep.numpy()
        print(name, tuple(rep.shape), rep.dtype)
    # upstream count / latest-time representations (imagenet.py:169-343), on the seconds-stamped tensor as the loader gives it
    for name in ("acc_count", "acc", "acc_count_pol", "acc_count_only", "acc_time", "acc_all", "acc_time_pol", "acc_exp", "flat", "flat_pol", "acc_intensity"):
        rep = getattr(ref, "reshape_then_" + name)(torch.tensor(ev.copy()), height=H, width=W)
        out["up_" + name] = rep.numpy()
        print(name, tuple(rep.shape), rep.dtype)
    # reshape_then_time_surface cannot run: it writes `.astype(int)` back into the f8 fields, which stay float, and
    # numba refuses float indices in to_timesurface_numpy (time_surface.py:67).  Expected output = the same lines with the
    # casts taking effect (an int-typed structured array), through the reference's own ToTimesurface.
    try:
        ref.reshape_then_time_surface(torch.tensor(ev_us.copy()), height=H, width=W)
        raise SystemExit("the reference wrapper ran: regenerate the fixture from it")
    except Exception as e:  # noqa: BLE001
        print("reference reshape_then_time_surface fails as expected:", type(e).__name__)
    d = np.zeros(n, dtype=[("x", "<i8"), ("y", "<i8"), ("t", "<f8"), ("p", "i1")])
    d["x"], d["y"], d["t"] = ev_us[:, 0].astype(int), ev_us[:, 1].astype(int), ev_us[:, 2]
    d["p"] = ((ev_us[:, 3] + 1) / 2).astype(np.int8)
    tt = d["t"]
    idx = np.searchsorted((tt - tt[0]) / (tt[-1] - tt[0]) * 6, np.arange(6) + 1)
    rep = ref.ToTimesurface(sensor_size=(W, H, 2), surface_dimensions=None, tau=50000, decay="exp")(d, idx)
    rep = rep.reshape((-1, rep.shape[-2], rep.shape[-1])).transpose(1, 2, 0)
    out["time_surface"] = np.ascontiguousarray(rep).astype(np.float32)
    print("time_surface", out["time_surface"].shape)
    # to_image: tonic's ToImage result viewed as an array that has .float()
    orig = ref.tonic_transforms.ToImage

    class ToImageF(orig):
        def __call__(self, e):
            return orig.__call__(self, e).view(_FloatArray)
    ref.tonic_transforms.ToImage = ToImageF
    rep = ref.reshape_then_to_image(torch.tensor(ev.copy()), height=H, width=W)
    ref.tonic_transforms.ToImage = orig
    out["to_image"] = rep.numpy()
    print("to_image", tuple(rep.shape), rep.dtype)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nimg_wrappers.npz"), **out)


if __name__ == "__main__":
    main()
