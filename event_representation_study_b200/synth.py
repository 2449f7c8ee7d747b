"""Seeded synthetic event streams.

Homogeneous spatial Poisson stream per window: x ~ U{0..W-1}, y ~ U{0..H-1}, exponential
inter-arrival times (rate N / duration) cumulated and floored to integer microseconds (t[0] = 0,
sorted, ties allowed), p ~ Bernoulli(0.5) mapped to {-1,+1} (or {0,1}).  A "clustered" variant puts
80 % of the events on a random 5 % of the pixels (atomic-contention stress).

numpy versions are used for parity tests and the host-resident end-to-end bench; `device_batch`
generates the same distribution directly in HBM with torch's Philox generator for the large configs.
"""
import numpy as np

DURATION_US = 300_000


def poisson_window(seed, n, H, W, polarity="pm1", clustered=False, duration_us=DURATION_US):
    """-> dict of x (u16), y (u16), t (i64), p (i8), time-sorted."""
    rng = np.random.default_rng(seed)
    if clustered and n > 0:
        n_hot = max(1, int(0.05 * H * W))
        hot = rng.choice(H * W, size=n_hot, replace=False)
        on_hot = rng.random(n) < 0.8
        lin = np.where(on_hot, hot[rng.integers(0, n_hot, n)], rng.integers(0, H * W, n))
        x, y = (lin % W).astype(np.uint16), (lin // W).astype(np.uint16)
    else:
        x = rng.integers(0, W, n).astype(np.uint16)
        y = rng.integers(0, H, n).astype(np.uint16)
    if n > 0:
        dt = rng.exponential(duration_us / max(n, 1), n)
        t = np.floor(np.cumsum(dt)).astype(np.int64)
        t -= t[0]
    else:
        t = np.zeros(0, np.int64)
    b = rng.random(n) < 0.5
    p = np.where(b, 1, -1 if polarity == "pm1" else 0).astype(np.int8)
    return {"x": x, "y": y, "t": t, "p": p}


def structured(ev, dtype="<i4"):
    """The reference's `fix_events_training` layout (gen1_2yolo.py:567-571 `<i4` x4; imagenet.py:1002-1006 `<f8` x4)."""
    out = np.zeros(len(ev["x"]), dtype=[("x", dtype), ("y", dtype), ("t", dtype), ("p", dtype)])
    for k in "xytp":
        out[k] = ev[k]
    return out


def pack_batch(windows):
    """List of event dicts -> CSR-packed SoA batch (x u16, y u16, t i32, p i8, offsets i64[B+1])."""
    offs = np.zeros(len(windows) + 1, np.int64)
    offs[1:] = np.cumsum([len(w["x"]) for w in windows])
    cat = lambda k, dt: (np.concatenate([np.asarray(w[k]) for w in windows]).astype(dt) if windows else np.zeros(0, dt))
    return {"x": cat("x", np.uint16), "y": cat("y", np.uint16), "t": cat("t", np.int32), "p": cat("p", np.int8), "offsets": offs}


def device_batch(B, n_per_window, H, W, device, seed=0, duration_us=DURATION_US, clustered=False):
    """Same distribution generated directly on `device` (torch Philox); returns torch tensors
    x,y (int16 storage of u16 values), t (int32, sorted per window), p (int8 in {-1,+1}), offsets (int64)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    N = B * n_per_window
    if clustered:
        n_hot = max(1, int(0.05 * H * W))
        hot = torch.randperm(H * W, generator=g, device=device)[:n_hot]
        on_hot = torch.rand(N, generator=g, device=device) < 0.8
        lin = torch.where(on_hot, hot[torch.randint(0, n_hot, (N,), generator=g, device=device)],
                          torch.randint(0, H * W, (N,), generator=g, device=device))
        x, y = (lin % W).to(torch.int16), (lin // W).to(torch.int16)
    else:
        x = torch.randint(0, W, (N,), generator=g, device=device, dtype=torch.int32).to(torch.int16)
        y = torch.randint(0, H, (N,), generator=g, device=device, dtype=torch.int32).to(torch.int16)
    u = torch.rand(B, n_per_window, generator=g, device=device, dtype=torch.float64).clamp_min(1e-300)
    dt = -torch.log(u) * (duration_us / max(n_per_window, 1))
    t = torch.floor(torch.cumsum(dt, dim=1))
    t = (t - t[:, :1]).to(torch.int32).reshape(-1)
    p = (torch.randint(0, 2, (N,), generator=g, device=device, dtype=torch.int32) * 2 - 1).to(torch.int8)
    offsets = torch.arange(B + 1, device=device, dtype=torch.int64) * n_per_window
    return {"x": x, "y": y, "t": t, "p": p, "offsets": offsets}
