#!/usr/bin/env python
"""BASELINE.json configs[4]: the Gromov-Wasserstein ranking matrix, 12 representations x 1k samples, on N GPUs.

    python bench_gwd.py [--gpus 1]                                              # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_gwd.py --gpus N

Every (representation r, sample s) entry is one GWD-A pair (compute_otmi.py:50-93: event sample Xs_s vs representation
pixels Xt_{r,s}, each subsampled to n = m = 1000 points, SURVEY.md 8d reading (ii)).  Samples are sharded in contiguous
column blocks over the ranks (sharding.shard_range); each rank evaluates its (R x S/N) block with evrep_gwd_kernel_l1 and
ONE all-gather over NCCL assembles the R x S matrix on every rank (sharding.gather_cost_matrix) - the only collective of the
whole design.  One JSON line from rank 0: pairs/s over all ranks (max-over-ranks device time, gather included), the
per-representation mean the ranking uses, and the numpy port timed on a bounded sample of the same pairs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def pair_inputs(r, s, n, rng_seed=900):
    """deterministic synthetic pair: what otmi() would hand to OTMI after its quadrant split (shapes and ranges only)"""
    rng = np.random.default_rng(rng_seed + s)
    Xs = rng.random((n, 4))                                     # events [x, y, t, p] in [0, 1]
    rr = np.random.default_rng(rng_seed * 7 + r * 100003 + s)
    Xt = np.concatenate([rr.random((n, 12)) * 255 * (rr.random((n, 12)) < 0.4), rr.random((n, 2))], 1)  # pixels [12 channels x 255, row, col]
    return Xs, Xt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reps", type=int, default=12)
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--points", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import event_representation_study_b200.batched as eb
    from event_representation_study_b200 import sharding
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    R, S, n = a.reps, a.samples, a.points
    lo, hi = sharding.shard_range(S, world, rank)
    # device-resident inputs of this rank's column block (generated on the host once, not timed)
    Xs_list, Xt_list = [], []
    for s in range(lo, hi):
        xs = None
        for r in range(R):
            Xs, Xt = pair_inputs(r, s, n)
            if xs is None:
                xs = torch.as_tensor(Xs, device=dev)
            Xs_list.append(xs)
            Xt_list.append(torch.as_tensor(Xt, device=dev))

    def step():
        local_costs = eb.gwd_kernel_l1(Xs_list, Xt_list, 0.7, device=dev).reshape(hi - lo, R).t().contiguous()  # (R, S_local)
        return sharding.gather_cost_matrix(local_costs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(max(a.warmup, 1)):
        M = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        M = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sec = float(ms.item()) * 1e-3 / a.steps
    assert tuple(M.shape) == (R, S)
    if rank == 0:
        from oracle import gwd as ogwd
        t0 = time.perf_counter()
        k, errs = 0, []
        while k < 8 and time.perf_counter() - t0 < 10.0:
            r, s = k % R, (k * 131) % S
            Xs, Xt = pair_inputs(r, s, n)
            want = ogwd.gwd_a_cost(Xs, Xt, 0.7)
            errs.append(abs(float(M[r, s]) - want) / abs(want))
            k += 1
        cpu = (time.perf_counter() - t0) / k
        print(json.dumps({
            "metric": "GWD pairs/s (GWD-A, 12 representations x 1k samples, n = m = 1000 points per pair)", "value": R * S / sec, "unit": "pairs/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "data": "synthetic", "config": {"workload": f"{R} x {S} GWD-A matrix, columns sharded over {world} GPU(s), one NCCL all-gather", "points": n},
            "ranking": [float(v) for v in M.mean(1).tolist()],
            "parity_spot_check": {"pairs": k, "max_rel_err_vs_oracle": max(errs)},
            "cpu_baseline": {"value": 1.0 / cpu, "unit": "pairs/s", "cores": "numpy/BLAS threads", "kind": "port",
                             "sample": f"{k} of the same pairs, oracle gwd_a_cost (closed form of POT's estimate; POT not installable offline)"},
            "speedup_vs_cpu_port": (R * S / sec) * cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
