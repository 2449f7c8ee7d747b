"""Fixtures for the sub-pixel branch of ev-licious' voxel grid (Events.divider > 1 -> float32 coordinates -> 4-tap bilinear
scatter, ev-licious/src/evlicious/tools/utils.py:70-76, 93-108), made by executing the reference's own
tools/utils.py::events_to_voxel_grid (loaded by path, like oracle/gen_golden.py does).

    python oracle/gen_golden_subpixel.py          # needs /root/reference; writes tests/golden/voxel_subpixel_*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_golden import _load_evlicious_utils  # noqa: E402


def main():
    utils, Events = _load_evlicious_utils()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for seed, (tag, H, W, div, n, bins, norm) in enumerate([("d2_small", 30, 40, 2, 3000, 5, False), ("d2_small_norm", 30, 40, 2, 3000, 5, True),
                                          ("d4_gen1", 60, 76, 4, 20000, 5, True), ("d3_edges", 8, 8, 3, 500, 3, False)]):
        rng = np.random.default_rng(7100 + seed)
        # raw sub-pixel integers: the scaled coordinate must satisfy max(x) <= width - 1 (events.py:30-33)
        x = rng.integers(0, (W - 1) * div + 1, n).astype(np.uint16)
        y = rng.integers(0, (H - 1) * div + 1, n).astype(np.uint16)
        t = np.sort(rng.integers(0, 100_000, n)).astype(np.int64)
        p = np.where(rng.random(n) < 0.5, 1, -1).astype(np.int8)
        E = Events(x.copy(), y.copy(), t.copy(), p.copy(), W, H, divider=div)
        out = utils.events_to_voxel_grid(E, bins, normalize=norm)
        np.savez_compressed(os.path.join(out_dir, f"voxel_subpixel_{tag}.npz"), out=out, x=x, y=y, t=t, p=p, H=H, W=W, divider=div, bins=bins,
                            normalize=norm)
        print(tag, out.shape, float(np.abs(out).sum()))


if __name__ == "__main__":
    main()
