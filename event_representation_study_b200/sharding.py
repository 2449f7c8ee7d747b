"""Multi-GPU plumbing: one process per GPU, windows sharded in contiguous blocks, no data-path
collective for representations; the only exchange is one all-gather of the small GWD cost matrix.

The reference fans windows out over an 8-process pool (ev-YOLOv6/yolov6/data/gen4/precompute_reps.py:444-463,
evlicious/tools/task_manager.py:8-33); here the unit of distribution is the same (one window), the workers are GPUs."""
import numpy as np


def shard_range(n_items, world_size, rank):
    """Contiguous block [lo, hi) of `n_items` for `rank`: the first n_items % world_size ranks get one extra item."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_offsets(offsets, world_size, rank):
    """CSR offsets of the whole batch -> (window range, event range, local offsets) of this rank's shard."""
    offsets = np.asarray(offsets, np.int64)
    lo, hi = shard_range(len(offsets) - 1, world_size, rank)
    e0, e1 = int(offsets[lo]), int(offsets[hi])
    return (lo, hi), (e0, e1), offsets[lo:hi + 1] - e0


def gather_cost_matrix(local_costs, group=None):
    """local_costs: (R, S_local) tensor of this rank's GWD costs (columns = its block of samples, see shard_range).
    -> (R, S) on every rank.  One all_gather of a few KB (NCCL over NVLink on GPUs, gloo in the CPU tests);
    ragged shards are padded to the widest one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_costs
    world = dist.get_world_size(group)
    R, s_loc = local_costs.shape
    widths = [torch.zeros(1, dtype=torch.int64, device=local_costs.device) for _ in range(world)]
    dist.all_gather(widths, torch.tensor([s_loc], dtype=torch.int64, device=local_costs.device), group=group)
    widths = [int(w.item()) for w in widths]
    wmax = max(widths)
    pad = torch.zeros((R, wmax), dtype=local_costs.dtype, device=local_costs.device)
    pad[:, :s_loc] = local_costs
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:, :w] for p, w in zip(parts, widths)], dim=1)


def bind_to_gpu_cpus(device_index=None):
    """Pin the calling process to the CPUs that NVML reports as local to its GPU, so that the pinned host buffers it allocates
    afterwards are first-touched on the GPU's NUMA node and its copy / launch threads run next to the root complex the GPU hangs
    on.  One process per GPU (torchrun): call it right after torch.cuda.set_device.  Returns a small report
    ({"cpus": n, "of": total, "numa_nodes": [...]}) or {"skipped": reason}; never raises - on a single-node host (every GPU
    local to all CPUs) it changes nothing."""
    import os
    try:
        import pynvml
        import torch
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(idx).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return {"skipped": "NVML reports no local CPU inside this process's affinity mask"}
        os.sched_setaffinity(0, cpus)
        nodes = set()
        for c in cpus:
            base = f"/sys/devices/system/cpu/cpu{c}"
            try:
                nodes |= {int(n[4:]) for n in os.listdir(base) if n.startswith("node") and n[4:].isdigit()}
            except OSError:
                pass
        return {"cpus": len(cpus), "of": len(allowed), "numa_nodes": sorted(nodes)}
    except Exception as e:  # noqa: BLE001  (placement is an optimisation: report, never fail the job)
        return {"skipped": f"{type(e).__name__}: {e}"}
