"""CPU ORACLE (numpy) for the event -> dense-representation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it, as the checker.

Every function restates one reference function and cites it (paths relative to /root/reference).
Pinning: the reference ships no tests / golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE FILES THEMSELVES, executed unmodified in the build
container by oracle/gen_golden.py (three third-party packages the reference imports are absent
offline and are replaced by the restatements in oracle/ref_shims/: torch_scatter, tonic, ot; parity
with those third-party packages is therefore "unpinned" and says so in DESIGN.md).  The resulting
fixtures live in tests/golden/*.npz and tests/test_oracle_golden.py checks this module against them.
"""
import numpy as np

# ----------------------------------------------------------------------------------------------
# MixedDensityEventStack / ERGO-12
# ----------------------------------------------------------------------------------------------
FUNCTIONS = ["timestamp", "polarity", "count", "timestamp_pos", "timestamp_neg", "count_pos", "count_neg"]
AGGREGATIONS = ["sum", "mean", "max", "variance", "min"]  # "min": torch_scatter's fifth reduce, legal through operations.py:30-35

# representations/optimized_representation.py:86-115 (v2, active) and :16-66 (v1, commented)
ERGO12_V2 = (
    [0, 3, 2, 6, 5, 6, 2, 5, 1, 0, 4, 1],
    ["polarity", "timestamp_neg", "count_neg", "polarity", "count_pos", "count", "timestamp_pos", "count_neg",
     "timestamp_neg", "timestamp_pos", "timestamp", "count"],
    ["variance", "variance", "mean", "sum", "mean", "sum", "mean", "mean", "max", "max", "max", "mean"],
)
ERGO12_V1 = (
    [0, 2, 2, 3, 5, 0, 0, 4, 2, 6, 1, 1],
    ["timestamp", "timestamp_pos", "timestamp_neg", "count_neg", "count_pos", "polarity", "timestamp", "count",
     "timestamp_pos", "count", "timestamp_pos", "timestamp_neg"],
    ["max", "sum", "mean", "sum", "mean", "variance", "variance", "sum", "mean", "sum", "sum", "sum"],
)


def sbn_window_bounds(n):
    """Index windows of create_windows, SBN branch (mixed_density_event_stack.py:48-74).
    Returns 7 half-open [lo, hi) pairs over the n events of the window."""
    n3 = n // 3
    b = [(0, n), (0, n3), (n3, 2 * n3), (2 * n3, 3 * n3)]
    c, s = n, 0
    for _ in range(3):
        c //= 2
        s += c
        b.append((min(s, n), n))
    return b


def _sbt_window_masks(t_s):
    """SBT branch (mixed_density_event_stack.py:76-107): time-based, 3 equispaced + 4 halvings."""
    masks = [np.ones(t_s.shape, bool)]
    f = 1 / 3
    for i in range(3):
        masks.append(np.logical_and(t_s <= (i + 1) * f, t_s >= i * f))
    cur = np.ones(t_s.shape, bool)
    factor = 1.0
    for _ in range(4):
        factor = factor / 2
        cur = cur & (t_s <= factor)  # successive filtering x = x[t <= factor]
        masks.append(cur.copy())
    return masks


def _scatter(src, index, size, reduce):
    """torch_scatter.scatter semantics (operations.py:15-37): sum / mean / max / min, empty -> 0."""
    if reduce == "sum":
        return np.bincount(index, weights=src, minlength=size).astype(np.float64)
    if reduce == "mean":
        s = np.bincount(index, weights=src, minlength=size)
        c = np.bincount(index, minlength=size).astype(np.float64)
        return s / np.maximum(c, 1.0)
    if reduce == "min":
        return -_scatter(-src, index, size, "max")
    if reduce == "max":
        o = np.full(size, -np.inf)
        nan_hit = np.zeros(size, bool)  # NaN sources propagate (torch amax semantics of the shim)
        isn = np.isnan(src)
        if isn.any():
            nan_hit[index[isn]] = True
        np.maximum.at(o, index[~isn], src[~isn])
        o[nan_hit] = np.nan
        touched = np.bincount(index, minlength=size) > 0
        o[~touched] = 0.0
        return o
    raise ValueError(reduce)


def _operation(x, y, t_s, p, func, agg, H, W):
    """Operations.exec + run (operations.py:39-89, 15-37) on one window subset; returns (H, W) f64."""
    pf = p.astype(np.float64)
    if func == "timestamp":
        sel = np.ones(len(x), bool)
    elif func == "polarity":
        sel = np.ones(len(x), bool)
    elif func == "count":
        sel = np.ones(len(x), bool)
    elif func in ("timestamp_pos", "count_pos"):
        sel = pf == 1
    elif func in ("timestamp_neg", "count_neg"):
        sel = pf == -1
        if sel.sum() == 0:  # operations.py:59-61,78-80
            sel = pf == 0
    else:
        raise ValueError(func)
    index = (x[sel].astype(np.float64) + y[sel].astype(np.float64) * W).astype(np.int64)
    if index.size and (index.min() < 0 or index.max() >= H * W):
        raise IndexError("scatter index out of range")  # torch_scatter raises -> swallowed by make_stack
    if func.startswith("timestamp"):
        src = t_s[sel].astype(np.float64)
    elif func == "polarity":
        src = pf[sel]
    else:
        src = np.ones(int(sel.sum()), np.float64)
    size = H * W
    with np.errstate(invalid="ignore"):
        if agg == "variance":
            m = _scatter(src, index, size, "mean")
            m2 = _scatter(src ** 2, index, size, "mean")
            out = m2 - m ** 2
        else:
            out = _scatter(src, index, size, agg)
    return out.reshape(H, W)


def mixed_density_event_stack(x, y, t, p, H, W, window_indexes, functions, aggregations, stacking_type="SBN"):
    """MixedDensityEventStack.stack (mixed_density_event_stack.py:25-46) -> float64 (H, W, C).
    Raises ValueError for an empty window exactly like the reference (`t.min()` on a size-0 array)."""
    x = np.asarray(x).astype(np.int32)
    y = np.asarray(y).astype(np.int32)
    p = np.asarray(p).astype(np.int32)
    t = np.asarray(t).astype(np.int64)
    assert len(x) == len(y) == len(p) == len(t)
    t = t - t.min()  # :33  (ValueError when empty)
    t = t - t.min()  # :112
    interval = t.max() - t.min()
    with np.errstate(divide="ignore", invalid="ignore"):
        t_s = t / interval  # :114, nan when interval == 0
    n = len(x)
    C = len(window_indexes)
    rep = np.zeros((H, W, C), np.float64)
    if stacking_type == "SBN":
        bounds = sbn_window_bounds(n)
        subsets = [slice(lo, hi) for lo, hi in bounds]
    elif stacking_type == "SBT":
        subsets = _sbt_window_masks(t_s)
    else:
        return rep  # create_windows returns only W0; any other index raises -> zero channel
    for c in range(C):
        w = window_indexes[c]
        try:
            s = subsets[w]
            rep[:, :, c] = _operation(x[s], y[s], t_s[s], p[s], functions[c], aggregations[c], H, W)
        except Exception:
            rep[:, :, c] = 0.0  # mixed_density_event_stack.py:120-127 swallow-and-zero
    return rep


def ergo12(x, y, t, p, H, W, version=2):
    """get_optimized_representation (optimized_representation.py:86-134)."""
    spec = ERGO12_V2 if version == 2 else ERGO12_V1
    return mixed_density_event_stack(x, y, t, p, H, W, *spec, stacking_type="SBN")


# ----------------------------------------------------------------------------------------------
# EventStack
# ----------------------------------------------------------------------------------------------
def event_stack_starts(n, stack_size):
    """Start indices s_k of the nested suffix windows (event_stack.py:70-82)."""
    c, s, out = n, 0, []
    for _ in range(stack_size):
        out.append(min(s, n))
        c //= 2
        s += c
    return out


def event_stack(x, y, t, p01, H, W, stack_size=12):
    """EventStack.pre_stack(ev, ev[-1].t) -> post_stack -> .transpose(0,1,3,2)[...,0]
    (event_stack.py:15-63, gen1_transforms.py:33-42), past branch only (t <= last timestamp).
    p01 is the polarity AFTER the caller's (p+1)//2 remap; pre_stack maps it to 2*p-1 in int8.
    out[y, x, k] = polarity of the latest past event at the pixel if its index >= s_k else 0."""
    x = np.asarray(x).astype(np.int32)
    y = np.asarray(y).astype(np.int32)
    t = np.asarray(t).astype(np.int64)
    pol = (2 * np.asarray(p01).astype(np.int8) - 1).astype(np.int8)
    if len(t) == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    past = t <= t[-1]
    x, y, pol = x[past], y[past], pol[past]
    n = len(x)
    lin = y.astype(np.int64) * W + x
    latest = np.full(H * W, -1, np.int64)
    np.maximum.at(latest, lin, np.arange(n, dtype=np.int64))
    out = np.zeros((H * W, stack_size), np.float32)
    has = latest >= 0
    lp = np.zeros(H * W, np.float32)
    lp[has] = pol[latest[has]].astype(np.float32)
    for k, s in enumerate(event_stack_starts(n, stack_size)):
        out[:, k] = np.where(latest >= s, lp, 0.0)
    return out.reshape(H, W, stack_size)


# ----------------------------------------------------------------------------------------------
# TimeSurface
# ----------------------------------------------------------------------------------------------
def time_surface_indices(t, n_surfaces=6):
    """gen1_transforms.py:78-80."""
    t = np.asarray(t)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_norm = (t - t[0]) / (t[-1] - t[0]) * n_surfaces
    return np.searchsorted(t_norm, np.arange(n_surfaces) + 1)


def time_surface(x, y, t, p01, indices, H, W, tau=50000.0, n_pol=2):
    """ToTimesurface.__call__ + to_timesurface_numpy (time_surface.py:25-74) -> f64 (S, P, H, W).
    Surface s is emitted only while `indices` is strictly increasing from the start and in range."""
    x = np.asarray(x).astype(np.int64)
    y = np.asarray(y).astype(np.int64)
    t = np.asarray(t)
    pp = np.asarray(p01).astype(np.int64) % n_pol  # numba negative index wrap-around
    S = len(indices)
    n = len(x)
    out = np.zeros((S, n_pol, H, W), np.float64)
    cell = (pp * H + y) * W + x
    order = np.arange(n, dtype=np.int64)
    pos = 0
    prev = -1
    for s in range(S):
        i = int(indices[s])
        if i <= prev or i >= n:
            break  # equality test can never fire again (time_surface.py:69)
        latest = np.full(n_pol * H * W, -1, np.int64)
        np.maximum.at(latest, cell[: i + 1], order[: i + 1])
        mem = np.full(n_pol * H * W, -(tau * 3 + 1), np.float64)
        has = latest >= 0
        mem[has] = t[latest[has]]
        out[s] = np.exp((mem - t[i]) / tau).reshape(n_pol, H, W)
        prev = i
        pos += 1
    return out


def time_surface_gen1(x, y, t, p_pm1, H, W):
    """The gen1_transforms.py:69-87 branch: -> f64 (H, W, 12) BEFORE the *255."""
    p01 = ((np.asarray(p_pm1) + 1) / 2).astype(np.int8)
    idx = time_surface_indices(t, 6)
    rep = time_surface(x, y, t, p01, idx, H, W, tau=50000.0)
    rep = rep.reshape((-1, H, W)).transpose(1, 2, 0)
    return rep


# ----------------------------------------------------------------------------------------------
# TORE
# ----------------------------------------------------------------------------------------------
def tore(x, y, ts, pol, sample_time, k, frame_size):
    """events2ToreFeature (tore.py:6-83) for time-sorted input -> float32 (Hf, Wf, 2k).
    Pixel = [y-1, x-1] (negative wraps like numpy).  Per pixel and polarity class (pol>0 / pol<=0):
    the k largest timestamps among events with ts < sample_time, as ages, ascending."""
    Hf, Wf = int(frame_size[0]), int(frame_size[1])
    x = np.asarray(x).astype(np.int64)
    y = np.asarray(y).astype(np.int64)
    ts = np.asarray(ts)
    pol = np.asarray(pol)
    keep = ts < sample_time
    ages_all = (sample_time - ts).astype(np.float64)
    out = np.full((Hf * Wf, 2 * k), np.inf, np.float64)
    for c, sel in enumerate([keep & (pol > 0), keep & (pol <= 0)]):
        yy = (y[sel] - 1) % Hf
        xx = (x[sel] - 1) % Wf
        lin = yy * Wf + xx
        ages = ages_all[sel]
        # k smallest ages per pixel, ascending
        order = np.lexsort((ages, lin))
        lin_s, ages_s = lin[order], ages[order]
        if len(lin_s):
            first = np.r_[True, lin_s[1:] != lin_s[:-1]]
            start = np.flatnonzero(first)
            rank = np.arange(len(lin_s)) - np.repeat(start, np.diff(np.r_[start, len(lin_s)]))
            m = rank < k
            out[lin_s[m], c * k + rank[m]] = ages_s[m]
    X = out.astype(np.float32)
    max_time = 500e6
    X[np.isnan(X)] = max_time
    X[X > max_time] = max_time
    X = np.log(X + 1)
    X -= np.log(150 + 1)
    X[X < 0] = 0
    return X.reshape(Hf, Wf, 2 * k)


def tore_gen1(x, y, t, p, k=6):
    """gen1_transforms.py:51-67 branch (1-based coordinates from the data minimum, data-dependent
    frame size) -> float32 (Hf, Wf, 12) BEFORE the *255."""
    x = np.asarray(x)
    y = np.asarray(y)
    x1 = x - x.min() + 1
    y1 = y - y.min() + 1
    return tore(x1, y1, t, p, t[-1], k, (int(y1.max()), int(x1.max())))


# ----------------------------------------------------------------------------------------------
# Voxel grids (three flavours) and the 2-channel histogram
# ----------------------------------------------------------------------------------------------
def voxel_tonic(x, y, t, p, H, W, n_bins):
    """tonic ToVoxelGrid((W,H,2), n_time_bins) as called at gen1_transforms.py:21-25 (restated from
    tonic 1.x to_voxel_grid_numpy; UNPINNED against real tonic) -> f64 (n_bins, H, W)."""
    t = np.asarray(t)
    tf = t.astype(float)
    with np.errstate(divide="ignore", invalid="ignore"):
        ts = n_bins * (tf - t[0]) / (t[-1] - t[0])
    pol = np.asarray(p).astype(np.float64).copy()
    pol[pol == 0] = -1
    with np.errstate(invalid="ignore"):
        ti = ts.astype(int)
    dt = ts - ti
    grid = np.zeros(n_bins * H * W, float)
    lin = np.asarray(x).astype(int) + np.asarray(y).astype(int) * W
    v = ti < n_bins
    np.add.at(grid, lin[v] + ti[v] * W * H, (pol * (1.0 - dt))[v])
    v = (ti + 1) < n_bins
    np.add.at(grid, lin[v] + (ti[v] + 1) * W * H, (pol * dt)[v])
    return grid.reshape(n_bins, H, W)


def voxel_evlicious(x, y, t, p, H, W, num_bins, normalize=True, t0_us=None, t1_us=None, divider=1):
    """evlicious.tools.events_to_voxel_grid (ev-licious/src/evlicious/tools/utils.py:51-85) -> float32 (num_bins, H, W):
    the integer-coordinate branch (uint16 x, y, divider == 1, :105-108) and, with divider > 1, the sub-pixel branch
    (Events.x = _x.astype(float32) / divider, events.py:37-47; 4-tap bilinear scatter, utils.py:93-103).
    Keeps the reference's weight quirk: _bil_w is fed t_norm_int, so the floor bin gets weight p and
    the next bin gets weight 0 (:74)."""
    grid = np.zeros((num_bins, H, W), np.float32)
    t = np.asarray(t).astype(np.int64)
    if len(t) < 2:
        return grid
    t0 = t0_us if t0_us is not None else t[0]
    t1 = t1_us if t1_us is not None else t[-1]
    dT = t1 - t0
    if dT == 0:
        dT = 1.0
    t_norm = (num_bins - 1) * (t - t0) / dT
    ti = t_norm.astype("int32")
    pp = np.asarray(p).astype(np.int8)
    if divider > 1:
        xf = np.asarray(x).astype("float32") / divider
        yf = np.asarray(y).astype("float32") / divider
        xi, yi = xf.astype("int32"), yf.astype("int32")
        for tl in [ti, ti + 1]:
            m = (tl >= 0) & (tl < num_bins)
            val = ((1 - np.abs(tl - ti)) * pp)[m]
            for xl in [xi[m], xi[m] + 1]:
                for yl in [yi[m], yi[m] + 1]:
                    w = (1 - np.abs(xl - xf[m])) * (1 - np.abs(yl - yf[m])) * val
                    mm = (xl >= 0) & (yl >= 0) & (xl < W) & (yl < H)
                    np.add.at(grid, (tl[m][mm], yl[mm], xl[mm]), w[mm])
    else:
        xx = np.asarray(x).astype(np.int64)
        yy = np.asarray(y).astype(np.int64)
        for tl in [ti, ti + 1]:
            m = (tl >= 0) & (tl < num_bins)
            w = (1 - np.abs(tl - ti)) * pp
            mm = m & (xx >= 0) & (yy >= 0) & (xx < W) & (yy < H)
            np.add.at(grid, (tl[mm], yy[mm], xx[mm]), w[mm])
    if normalize:
        nz = np.nonzero(grid)
        if nz[0].size > 0:
            mean, std = grid[nz].mean(), grid[nz].std()
            if std > 0:
                grid[nz] = (grid[nz] - mean) / (1e-5 + std)
    return grid


def voxel_gwd(x, y, t01, p, W, H, bins=5):
    """compute_repr (representation_search/gromov_wasserstein.py:72-82) -> f64 (H, W, bins)."""
    grid = np.zeros((H, W, bins))
    t01 = np.asarray(t01, np.float64)
    b = (bins - 1) * t01
    bi = b.astype("int")
    pp = np.asarray(p)
    for bl in [bi, bi + 1]:
        w = 1 - np.abs(bl - b)
        m = bl < bins
        np.add.at(grid, (np.asarray(y)[m], np.asarray(x)[m], bl[m]), w[m] * pp[m])
    return grid


def to_image(x, y, p01, H, W):
    """tonic ToImage((W,H,2)) as called at gen1_transforms.py:44-49 -> int16 (2, H, W) counts."""
    f = np.zeros((2, H, W), np.int16)
    np.add.at(f, (np.asarray(p01).astype(int), np.asarray(y).astype(int), np.asarray(x).astype(int)), 1)
    return f
