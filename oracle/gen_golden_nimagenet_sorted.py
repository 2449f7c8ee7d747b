"""Writes tests/golden/nimg_sorted.npz by running the reference's two rank-based N-ImageNet loaders UNMODIFIED:
reshape_then_acc_sort ("sorted time surface", imagenet.py:513-838) under the option combinations that can run
(`denoise_image` / `denoise_sort` call density_filter_event_image, which the reference never defines) and
reshape_then_acc_adj_sort (DiST, imagenet.py:873-999).  Runs only where /root/reference exists; torch_scatter comes from
oracle/ref_shims.  TEST INFRASTRUCTURE ONLY."""
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

spec = importlib.util.spec_from_file_location("ref_imagenet", os.path.join(REF, "n_imagenet", "real_cnn_model", "data", "imagenet.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

BASE = dict(neglect_polarity=False, global_time=True, strict=False, use_image=False, denoise_sort=False, denoise_image=False,
            filter_flash=False, filter_noise=False, quantize_sort=None)
SORT_CASES = {
    "default": {},
    "neglect": dict(neglect_polarity=True),
    "strict": dict(strict=True),
    "strict_neglect": dict(strict=True, neglect_polarity=True),
    "image": dict(use_image=True),
    "image_neglect_strict": dict(use_image=True, neglect_polarity=True, strict=True),
    "quant8": dict(quantize_sort=8),
    "quant_list": dict(quantize_sort=[4, 16], use_image=True),
    "quant_list_neglect": dict(quantize_sort=[4, 16], neglect_polarity=True, strict=True),
    "local_time": dict(global_time=False),
    "local_time_strict": dict(global_time=False, strict=True),
}


def events(seed, n, H, W, dup_frac=0.3):
    """seconds-stamped sample like parse_event delivers; a share of the stamps repeats (microsecond ties)"""
    rng = np.random.default_rng(seed)
    t = np.sort(np.floor(rng.random(n) * 40000.0)) / 1e6 + 1.25  # integer microseconds above 1.25 s: ties are common
    if dup_frac:
        k = rng.random(n) < dup_frac
        t[1:][k[1:]] = t[:-1][k[1:]]
        t = np.sort(t)
    return np.stack([rng.integers(0, W, n).astype(np.float64), rng.integers(0, H, n).astype(np.float64), t, rng.choice([-1.0, 1.0], n)], 1)


def main():
    out = {}
    for tag, (seed, n, H, W) in {"a": (11, 3000, 32, 48), "b": (12, 8000, 64, 64), "c": (13, 40, 16, 16)}.items():
        ev = events(seed, n, H, W)
        out[f"{tag}_events"], out[f"{tag}_H"], out[f"{tag}_W"] = ev, H, W
        for name, kw in SORT_CASES.items():
            rep = ref.reshape_then_acc_sort(torch.tensor(ev.copy()), height=H, width=W, **{**BASE, **kw})
            out[f"{tag}_sort_{name}"] = rep.numpy()
            print(tag, "sort", name, tuple(rep.shape), rep.dtype)
        rep = ref.reshape_then_acc_adj_sort(torch.tensor(ev.copy()), height=H, width=W, **BASE)
        out[f"{tag}_dist"] = rep.numpy()
        print(tag, "dist", tuple(rep.shape), rep.dtype, float(rep.max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nimg_sorted.npz"), **out)


if __name__ == "__main__":
    main()
