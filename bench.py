#!/usr/bin/env python
"""Headline benchmark: Gevents/s into ERGO-12 at 1 Mpx (BASELINE.json configs[3], one rank's shard).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm on the host cores

A step = one pass of the hot path (raw events of a batch of windows -> ERGO-12 tensors) over one synthetic
batch: 32 windows x 1,000,000 events on a 1280x720 sensor per GPU (config 4 shards its 256 windows 32 per GPU,
so per-GPU work is fixed: weak scaling, no data-path collective).  One JSON line is printed by rank 0.

value     events of all ranks / max-over-ranks device time, inputs resident in HBM
e2e       same through the public batched API with HOST (pinned) event arrays: per step the host->device copy
          of the events, the kernels, and the device->host read of a per-window checksum of the output
roofline  the kernel that writes the output (k_md_tile), timed with CUDA events inside the timed region
cpu_baseline  the numpy oracle (a port of the reference algorithm) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, C = 720, 1280, 12
BYTES_PER_EVENT = 9  # x u16 + y u16 + t i32 + p i8: SURVEY.md 8(d)
METRIC = "Gevents/s into ERGO-12 @1Mpx 1280x720"
# kernels of one evrep_ergo12_batched call: k_init, k_hist, k_colscan, k_scan, k_bin, k_md_tile_static, k_md_tile_heavy
KERNELS_PER_CALL = 7


def algorithmic_bytes(n_windows, n_events):
    return n_windows * (n_events * BYTES_PER_EVENT + H * W * C * 4)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_per_launch(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, uuid):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", uuid, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference algorithm)
# ------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def _cpu_one(args):
    """One window through the oracle's ERGO-12; the synthetic window is generated once per worker slot and reused, so
    timed steps contain the representation work only (like the GPU arm, whose inputs are generated before timing)."""
    slot, n = args
    from oracle import representations as orep
    from event_representation_study_b200.synth import poisson_window
    key = (slot, n) if slot >= 0 else n  # slot < 0: one cached window per worker process
    w = _CPU_CACHE.get(key)
    if w is None:
        w = _CPU_CACHE[key] = poisson_window(5000 + (slot if slot >= 0 else os.getpid() % 1000), n, H, W)
    t0 = time.perf_counter()
    out = orep.ergo12(w["x"], w["y"], w["t"], w["p"], H, W)
    return time.perf_counter() - t0, float(out[:, :, 5].sum())


def cpu_baseline_scalar(n_events, budget_s=12.0, max_windows=64):
    """Single-process oracle on a bounded sample of the same workload (windows of the bench's size)."""
    done, spent = 0, 0.0
    while done < max_windows and spent < budget_s:
        dt, _ = _cpu_one((done, n_events))
        spent += dt
        done += 1
    return {"value": done * n_events / spent / 1e9, "unit": "Gevents/s", "cores": 1, "kind": "port",
            "sample": f"{done} windows of {n_events} events at {W}x{H}, numpy oracle (oracle/representations.py::ergo12), "
                      f"generation excluded, {spent:.1f} s"}


def run_reference_arm(a):
    """The reference's CPU implementation of the path.  /root/reference is pure Python and absent on the GPU box,
    and three of its imports (torch_scatter, tonic, POT) are not installable offline, so this arm times the numpy
    port of its algorithm (oracle/), one window per worker process on every host core - the shape of the
    reference's own 8-process TaskManager pool (ev-YOLOv6/yolov6/data/gen4/precompute_reps.py:444)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    per_step = procs  # one window per worker per step
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        chunk = lambda: pool.map(_cpu_one, [(-1, a.events)] * per_step, chunksize=1)
        for _ in range(max(a.warmup, 1)):  # the first pass also generates each worker's window
            chunk()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            chunk()
        dt = time.perf_counter() - t0
    val = a.steps * per_step * a.events / dt / 1e9
    sample = (f"{per_step} windows of {a.events} events per step, one per worker process ({procs} processes = all host cores), "
              f"numpy port of the reference algorithm (oracle/representations.py::ergo12), window generation excluded")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gevents/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(a, a.windows),
        "cpu_baseline": {"value": val, "unit": "Gevents/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def config_dict(a, windows):
    return {"workload": f"ERGO-12 v2, {W}x{H}, {a.events} ev/window, {windows} windows/GPU (BASELINE configs[3] shard: 256 windows over 8 GPUs)",
            "windows_per_gpu": windows, "events_per_window": a.events, "stream": "poisson-uniform" + ("-clustered" if a.clustered else ""),
            "l2": "inputs (288 MB) and outputs (1.4 GB) per step exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(a):
    import torch
    import torch.distributed as dist
    import event_representation_study_b200.batched as eb
    from event_representation_study_b200 import _lib
    from event_representation_study_b200.synth import device_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, N = a.windows, a.events
    d = device_batch(B, N, H, W, dev, seed=1000 * 4 + rank, clustered=a.clustered)
    ev = eb.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
    out = torch.empty((B, H, W, C), dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # clocks: nvidia-smi needs ~100 ms to start reporting, the timed region is ~15 ms - start it before the warm-up (the
    # same work as the timed steps, so every sample is taken under load) and keep the GPU busy until the region ends
    uuid = str(torch.cuda.get_device_properties(local).uuid)
    uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
    eb.ergo12(ev, H, W, out=out)
    torch.cuda.synchronize()
    flags = eb.window_flags(ev)
    assert (flags == 0).all(), f"window flags {flags}"
    sampler = ClockSampler(uuid) if rank == 0 else None
    t_w = time.perf_counter()
    n_w = 0
    while n_w < max(a.warmup, 3) or time.perf_counter() - t_w < 0.25:
        eb.ergo12(ev, H, W, out=out)
        n_w += 1
        if n_w % 8 == 0:
            torch.cuda.synchronize()

    # ---- device-resident throughput, with per-kernel events ------------------------------------
    _lib.profile_enable(a.steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        eb.ergo12(ev, H, W, out=out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kern = {k: _lib.profile_read(k) for k in (_lib.K_COUNT, _lib.K_SCAN, _lib.K_BIN, _lib.K_TILE)}
    _lib.profile_enable(0)
    clocks = sampler.stop() if sampler else None
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all = float(tmax.item())
    value = world * B * N * a.steps / (ms_all * 1e-3) / 1e9

    # ---- end to end: pinned host events -> H2D -> kernels -> per-window checksum -> D2H ----------
    # Public API only (EventBatch + ergo12).  Every step copies its own 288 MB of events from pinned host memory and reads
    # its own result back inside the timed region.  The step is cut into groups of windows; a copy stream runs ahead of the
    # compute stream (one device buffer per group, recycled when the group's kernels are done), so the copies of step s + 1
    # are already queued when the host waits for the result of step s - the prefetch any input pipeline does.
    host = {k: d[k].cpu().pin_memory() for k in ("x", "y", "t", "p")}
    n_groups = min(a.e2e_groups, B)
    bounds = [B * g // n_groups for g in range(n_groups + 1)]
    offs = ev.offsets
    res_host = [torch.empty(B, dtype=torch.float64).pin_memory() for _ in range(2)]
    copy_st, comp_st = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    dbuf = [{k: torch.empty(int(offs[bounds[g + 1]] - offs[bounds[g]]), dtype=d[k].dtype, device=dev) for k in ("x", "y", "t", "p")}
            for g in range(n_groups)]
    ready = [torch.cuda.Event() for _ in range(n_groups)]   # group's events are on the device
    done = [torch.cuda.Event() for _ in range(n_groups)]    # group's kernels no longer read its buffer

    def enqueue_copies(first):
        with torch.cuda.stream(copy_st):
            for g_ in range(n_groups):
                if not first:
                    copy_st.wait_event(done[g_])
                e0, e1 = int(offs[bounds[g_]]), int(offs[bounds[g_ + 1]])
                for k in ("x", "y", "t", "p"):
                    dbuf[g_][k].copy_(host[k][e0:e1], non_blocking=True)
                ready[g_].record(copy_st)

    def enqueue_compute(step):
        rh = res_host[step & 1]
        with torch.cuda.stream(comp_st):
            for g_ in range(n_groups):
                w0, w1 = bounds[g_], bounds[g_ + 1]
                comp_st.wait_event(ready[g_])
                sub = eb.EventBatch(dbuf[g_]["x"], dbuf[g_]["y"], dbuf[g_]["t"], dbuf[g_]["p"], offs[w0:w1 + 1] - int(offs[w0]))
                o = eb.ergo12(sub, H, W, out=out[w0:w1])
                done[g_].record(comp_st)
                rh[w0:w1].copy_(o.view(w1 - w0, -1).sum(1, dtype=torch.float64), non_blocking=True)
            fin = torch.cuda.Event()
            fin.record(comp_st)
        return fin, rh

    def e2e_run(steps):
        enqueue_copies(first=True)
        last = None
        for s_ in range(steps):
            fin, rh = enqueue_compute(s_)
            if s_ + 1 < steps:
                enqueue_copies(first=False)   # step s + 1's events: queued before the host blocks on step s's result
            fin.synchronize()                 # the caller reads this step's result
            last = float(rh.sum())
        return last

    sampler2 = ClockSampler(uuid) if rank == 0 else None
    e2e_run(max(2, min(40, int(0.25 / max(ms_all / a.steps * 8e-3, 1e-4)))))  # ~0.25 s of the same work while nvidia-smi starts
    barrier()
    t_e2e0 = time.perf_counter()
    e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0_.record()
    e2e_run(a.steps)
    torch.cuda.current_stream(dev).wait_stream(comp_st)
    e1_.record()
    barrier()
    wall_ms = (time.perf_counter() - t_e2e0) * 1e3
    ms_e2e = torch.tensor([max(e0_.elapsed_time(e1_), wall_ms)], device=dev, dtype=torch.float64)  # events on another stream: trust the slower clock
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * N * a.steps / (float(ms_e2e.item()) * 1e-3) / 1e9
    if sampler2 is not None:  # merge the samples of both timed regions (the first one is only ~15 ms long)
        c2 = sampler2.stop()
        if clocks and c2.get("samples"):
            n1, n2 = clocks.get("samples", 0), c2["samples"]
            clocks = {"sm_mhz": c2["sm_mhz"] if n2 >= n1 else clocks["sm_mhz"], "sm_max_mhz": max(clocks["sm_max_mhz"] or 0, c2["sm_max_mhz"] or 0),
                      "reasons": sorted(set(clocks["reasons"]) | set(c2["reasons"])), "samples": n1 + n2,
                      "regions": {"device_resident": {"sm_mhz": clocks["sm_mhz"], "samples": n1}, "end_to_end": {"sm_mhz": c2["sm_mhz"], "samples": n2}}}
    h2d = int(sum(host[k].numel() * host[k].element_size() for k in host))
    e2e_launches = a.steps * n_groups * KERNELS_PER_CALL

    if rank == 0:
        peak, peak_src = measured_peak()
        tile_ms, tile_n = kern[_lib.K_TILE]
        tile_avg_s = tile_ms / max(tile_n, 1) * 1e-3
        alg = algorithmic_bytes(B, N)
        achieved = alg / tile_avg_s / 1e9
        step_s = ms_all / a.steps * 1e-3
        line = {
            "metric": METRIC, "value": value, "unit": "Gevents/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "warmup_note": "warm-up runs at least --warmup steps and at least 0.25 s (clock sampler start-up)",
            "ms_per_step": ms_all / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/f64->f32",
            "data": "synthetic", "config": config_dict(a, B),
            "e2e": {"value": e2e_value, "unit": "Gevents/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * 8,
                    "gpu_launches": e2e_launches,
                    "note": f"PCIe-bound: 9 B/event over the host link; {n_groups} window groups per step, copy stream prefetching the next step's events; the dense output stays on the GPU for the model"},
            "gpu_launches": a.steps * KERNELS_PER_CALL,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_md_tile (per-tile reduction + finalise, writes the output)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_per_launch("ergo12_1mpx_b32"),
                         "algorithmic_bytes_per_launch": alg, "launch_ms": tile_avg_s * 1e3, "peak_source": peak_src,
                         "whole_step": {"achieved": alg / step_s / 1e9, "frac": alg / step_s / 1e9 / peak,
                                        "note": "same algorithmic bytes over the whole 7-kernel step"},
                         "kernel_ms": {_lib.KERNEL_NAMES[k]: (v[0] / max(v[1], 1)) for k, v in kern.items()}},
        }
        if world == 1 and not a.no_cpu:
            line["cpu_baseline"] = cpu_baseline_scalar(N)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=32, help="windows per GPU")
    ap.add_argument("--events", type=int, default=1_000_000, help="events per window")
    ap.add_argument("--clustered", action="store_true", help="80%% of the events on 5%% of the pixels (contention stress; not the headline)")
    ap.add_argument("--e2e-groups", type=int, default=4, help="window groups per step in the end-to-end leg (copy/compute overlap)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_gpu_arm(a)


if __name__ == "__main__":
    main()
