"""Writes tests/golden/mdmin_*.npz: the reference's MixedDensityEventStack (mixed_density_event_stack.py:25-151 with
operations.py:15-89, both UNMODIFIED) run with the aggregation "min" among the others.  Operations.run hands the
aggregation string straight to torch_scatter's `scatter(..., reduce=...)` (operations.py:30-35), so "min" is a legal
fifth aggregation of the reference even though the study's search space stops at four; the N-ImageNet scatter_min planes
(imagenet.py:241-244, 383-386) are the same reduction.  TEST INFRASTRUCTURE ONLY; runs where /root/reference exists.
torch_scatter comes from oracle/ref_shims (parity with the real package unpinned, as for the other aggregations)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  (sets up sys.path for the reference and the shims)


def main():
    import torch  # noqa: F401
    from representations.representation_search.mixed_density_event_stack import MixedDensityEventStack

    rng = np.random.default_rng(77)
    FUNCS = ["timestamp", "polarity", "count", "timestamp_pos", "timestamp_neg", "count_pos", "count_neg"]
    AGGS = ["sum", "mean", "max", "variance", "min"]
    keep = {"n1_pm1", "n7_pm1", "n2000_pm1", "n2000_01", "n500_allpos", "n600_ties", "n50_tconst", "n4000_dups"}
    with np.errstate(all="ignore"):
        for tag, H, W, ev in gg.streams():
            if tag not in keep:
                continue
            n = len(ev["x"])
            s4 = gg.structured(ev, "<i4")
            for st in ["SBN", "SBT"]:
                # every function under "min" on random windows, then two random other channels
                C = 9
                wi = rng.integers(0, 7 if st == "SBN" else 8, C).tolist()
                fu = FUNCS + [FUNCS[i] for i in rng.integers(0, 7, 2)]
                ag = ["min"] * 7 + [AGGS[i] for i in rng.integers(0, 5, 2)]
                out = MixedDensityEventStack(C, n, H, W, (wi, fu, ag), st).stack(s4.copy())
                gg.save(f"mdmin_{st}_{tag}", out=out, win=np.array(wi), func=np.array(fu), agg=np.array(ag), stacking=st,
                        x=ev["x"], y=ev["y"], t=ev["t"], p=ev["p"], H=H, W=W)


if __name__ == "__main__":
    main()
