"""Where the per-window drop-in call spends its time (cProfile over get_item_transform on one 1 Mpx window)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from event_representation_study_b200 import synth  # noqa: E402
from event_representation_study_b200.representations.gen1_transforms import get_item_transform  # noqa: E402
from event_representation_study_b200.representations.representation_search.mixed_density_event_stack import MixedDensityEventStack  # noqa: E402

h, w, N = 720, 1280, 1_000_000
wdw = synth.poisson_window(4242, N, h, w)
data = synth.structured(wdw, "<i4")
call = lambda: get_item_transform(data, str(MixedDensityEventStack), MixedDensityEventStack, h, w, N)
for _ in range(3):
    call()
torch.cuda.synchronize()
ts = []
for _ in range(7):
    t0 = time.perf_counter()
    call()
    ts.append(time.perf_counter() - t0)
print("median ms", np.median(ts) * 1e3, "all", [round(t * 1e3, 1) for t in ts])
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    call()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
