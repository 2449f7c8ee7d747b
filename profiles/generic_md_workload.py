"""Interpreted (run-time plan) mixed-density kernel, the compile-time ERGO-12 kernel and run-time specialised kernels
(evrep_mixed_density_specialize, NVRTC) on the headline workload: 32 windows x 1 M events at 1280 x 720.  The representation
search of the reference evaluates arbitrary (window, function, aggregation) tuples (mixed_density_event_stack.py:25-151)."""
import time
import json
import random
import sys
import torch
import event_representation_study_b200.batched as eb
from event_representation_study_b200 import _lib
from event_representation_study_b200.synth import device_batch

H, W, B, N = 720, 1280, 32, 1_000_000
dev = torch.device("cuda", 0)
d = device_batch(B, N, H, W, dev, seed=3)
ev = eb.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
FUNCS = ["timestamp", "polarity", "count", "timestamp_pos", "timestamp_neg", "count_pos", "count_neg"]
AGGS = ["sum", "mean", "max", "variance"]
ergo_w = [0, 3, 2, 6, 5, 6, 2, 5, 1, 0, 4, 1]
ergo_f = ["polarity", "timestamp_neg", "count_neg", "polarity", "count_pos", "count", "timestamp_pos", "count_neg", "timestamp_neg", "timestamp_pos", "timestamp", "count"]
ergo_a = ["variance", "variance", "mean", "sum", "mean", "sum", "mean", "mean", "max", "max", "max", "mean"]


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


cases = [("ergo12 v2 (static)", ergo_w, ergo_f, ergo_a, "SBN")]
w2 = list(ergo_w); w2[11] = 2
cases.append(("ergo12 v2, one window changed (generic)", w2, ergo_f, ergo_a, "SBN"))
rng = random.Random(7)
for k in range(4):
    cases.append((f"random tuple {k}", [rng.randrange(7) for _ in range(12)], [rng.choice(FUNCS) for _ in range(12)], [rng.choice(AGGS) for _ in range(12)], "SBN"))
cases.append(("ergo12 tuple, SBT windows", ergo_w, ergo_f, ergo_a, "SBT"))
out = torch.empty((B, H, W, 12), device=dev, dtype=torch.float32)
for name, w, f, a, st in cases:
    fn = lambda: eb.mixed_density(ev, H, W, w, f, a, st, out=out, specialize=False)
    _lib.lib.evrep_profile_enable(0)
    ms = timed(fn)
    rec = {"case": name, "ms_per_step": round(ms, 4), "gev_s": round(B * N / ms / 1e6, 2)}
    if "static" not in name:
        ref = out.clone()
        t0 = time.time()
        ok = eb.specialize_mixed_density(w, f, a, st, max_events_per_window=N)
        rec["specialize_s"] = round(time.time() - t0, 2)
        rec["specialized"] = ok
        if ok:
            ms2 = timed(fn)
            rec.update({"specialized_ms_per_step": round(ms2, 4), "specialized_gev_s": round(B * N / ms2 / 1e6, 2), "speedup": round(ms / ms2, 2)})
            d = (out - ref).abs()
            tol = 2e-7 + 1e-5 * ref.abs()
            rec["max_err_over_tol_vs_interpreted"] = round(float((d / tol)[~torch.isnan(ref)].max()), 4)
    rec.update({"windows": w, "functions": f, "aggregations": a, "stacking": st})
    print(json.dumps(rec))
    sys.stdout.flush()
