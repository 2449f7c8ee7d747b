"""CPU-only checks: the C-ABI library loads and exports every symbol include/evrep.h declares, host-side logic
(accumulator plans, workspace sizing, argument validation, window sharding) and the world_size-2 exchange over gloo.
No kernel is launched here."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from event_representation_study_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "evrep.h")).read()
    declared = set(re.findall(r"\b(evrep_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 16
    raw = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"libevrep.so does not export {name}"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert L.lib.evrep_version() == 100


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "event_representation_study_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def plan_info(L, H, W, win, func, agg, stacking, nmax):
    w, f, a = (np.array(v, np.int8) for v in (win, func, agg))
    info = np.zeros(5, np.int32)
    rc = L.lib.evrep_mixed_density_plan_info(H, W, w.ctypes.data, f.ctypes.data, a.ctypes.data, len(w), stacking, nmax, info.ctypes.data)
    return rc, info


def test_ergo12_plan(L):
    from event_representation_study_b200.batched import ERGO12_V2
    win, func, agg = ERGO12_V2
    f = [L.FUNCS[x] for x in func]
    a = [L.AGGS[x] for x in agg]
    rc, info = plan_info(L, 720, 1280, win, f, a, 0, 1_000_000)
    assert rc == 0
    bytes_px, tile_px, tiles, words, smem = info
    assert words == 24 and bytes_px == 100 and tile_px == 1024 and tiles == 900 and smem == 102400
    rc, info = plan_info(L, 240, 304, win, f, a, 0, 50_000)  # fewer events -> wider limbs -> fewer words
    assert rc == 0 and info[3] < 24 and info[2] == -(-240 * 304 // info[1])


def test_plan_for_the_min_aggregation(L):
    """EVREP_AGG_MIN: one word per (window, class) like "max"; a (count, min) channel only needs a presence bit"""
    T, CP, TP = L.FUNCS["timestamp"], L.FUNCS["count_pos"], L.FUNCS["timestamp_pos"]
    MIN, MAX = L.AGGS["min"], L.AGGS["max"]
    assert MIN == 4
    rc, one = plan_info(L, 64, 64, [0], [TP], [MAX], 0, 1000)
    rc2, two = plan_info(L, 64, 64, [0, 0], [TP, TP], [MAX, MIN], 0, 1000)
    rc3, mixed = plan_info(L, 64, 64, [0, 0, 1], [TP, T, CP], [MIN, MIN, MIN], 0, 1000)
    assert rc == rc2 == rc3 == 0
    assert one[3] == 1 and two[3] == 2 and mixed[3] == 3  # accumulator words: latest + earliest; two earliest words + presence
    rc, _ = plan_info(L, 64, 64, [0], [T], [5], 0, 1000)   # an unknown aggregation is a zero channel, not an error
    assert rc == 0


def test_plan_rejects_bad_arguments(L):
    rc, _ = plan_info(L, 720, 1280, [0] * 33, [0] * 33, [0] * 33, 0, 1000)
    assert rc == L.EINVAL and b"outside" in L.lib.evrep_last_error()
    rc, _ = plan_info(L, 720, 1280, [0], [0], [0], 7, 1000)
    assert rc == L.EINVAL
    rc, _ = plan_info(L, 0, 1280, [0], [0], [0], 0, 1000)
    assert rc == L.EINVAL
    rc, info = plan_info(L, 30, 40, [9, -1, 0], [2, 2, 99], [0, 0, 0], 0, 1000)  # bad window / function -> zero channels, not errors
    assert rc == 0


def test_workspace_bytes(L):
    f = L.lib.evrep_workspace_bytes
    small = f(L.OP_MIXED_DENSITY, 1, 1000, 240, 304, 12)
    big = f(L.OP_MIXED_DENSITY, 32, 32_000_000, 720, 1280, 12)
    assert 0 < small < big and big >= 32_000_000 * 8
    assert f(L.OP_VOXEL, 32, 32_000_000, 720, 1280, 12) < 1 << 20  # direct-scatter ops need no record buffer
    assert f(0, 1, 1, 1, 1, 1) == 0 and f(L.OP_TORE, -1, 1, 1, 1, 1) == 0


def test_validation_happens_before_any_cuda_call(L):
    offs = np.array([0, 10], np.int64)
    rc = L.lib.evrep_ergo12_batched(None, None, None, 4, None, offs.ctypes.data, 1, 30, 40, 2, None, None, 0, None)
    assert rc == L.EINVAL and b"null" in L.lib.evrep_last_error()
    rc = L.lib.evrep_ergo12_batched(None, None, None, 3, None, offs.ctypes.data, 1, 30, 40, 2, None, None, 0, None)
    assert rc == L.EINVAL and b"t_bytes" in L.lib.evrep_last_error()
    rc = L.lib.evrep_event_stack_batched(None, None, None, 4, None, offs.ctypes.data, 0, 30, 40, 12, None, None, 0, None)
    assert rc == L.OK  # B == 0: nothing to do
    assert L.lib.evrep_gwd_workspace_bytes(None, None, 3) == 0


def test_batched_api_refuses_cpu_tensors():
    import torch
    import event_representation_study_b200.batched as eb
    z = torch.zeros(4, dtype=torch.int16)
    with pytest.raises(ValueError, match="CUDA"):
        eb.EventBatch(z, z, torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.int8), np.array([0, 4]))


def test_shard_range_and_offsets():
    from event_representation_study_b200.sharding import shard_offsets, shard_range
    for n, w in [(256, 8), (10, 4), (3, 8), (0, 2)]:
        blocks = [shard_range(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(256, 8, 3) == (96, 128)  # BASELINE config 4: 32 windows per GPU
    offs = np.array([0, 5, 5, 12, 20, 21])
    (lo, hi), (e0, e1), local = shard_offsets(offs, 2, 1)
    assert (lo, hi) == (3, 5) and (e0, e1) == (12, 21) and local.tolist() == [0, 8, 9]
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, n_samples, q):
    import torch
    import torch.distributed as dist
    from event_representation_study_b200.sharding import gather_cost_matrix, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    R = 12
    full = torch.arange(R * n_samples, dtype=torch.float64).reshape(R, n_samples)
    lo, hi = shard_range(n_samples, world, rank)
    out = gather_cost_matrix(full[:, lo:hi].contiguous())
    q.put((rank, bool(torch.equal(out, full)), tuple(out.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [10, 7])
def test_gwd_matrix_all_gather_world_size_2(n_samples):
    """The one collective of the design: R x S/G blocks gathered into the R x S matrix, ragged included."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_samples, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res) and all(shape == (12, n_samples) for _, _, shape in res)


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference must work without a GPU (tiny sample here)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--events", "20000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_cpu_arms_do_not_load_the_product_library():
    """the reference arm's worker (and the cpu_baseline leg) must import nothing from the product package: its
    __init__ dlopens libevrep.so, which would show up as native code loaded by the CPU arm"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import bench; dt, s = bench._cpu_one((0, 3000)); "
            "bad = [m for m in sys.modules if m.startswith('event_representation_study_b200')]; "
            "maps = open('/proc/self/maps').read(); assert not bad, bad; assert 'libevrep' not in maps; print('clean', dt > 0)") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "clean True" in out.stdout, out.stderr[-2000:]


def test_value_layer_compiles_to_the_same_piecewise_linear_function():
    """host logic of the EST path: the LeakyReLU MLP of one scalar, compiled to breakpoints / slopes / intercepts, is the
    same function as the MLP (float64, 1e-12) - on the reference-trained weights of the fixture and on random ones"""
    from event_representation_study_b200.est import compile_value_layer
    z = np.load(os.path.join(ROOT, "tests", "golden", "est_small.npz"))
    rng = np.random.default_rng(0)
    cases = [([z[f"w{i}"] for i in range(3)], [z[f"b{i}"] for i in range(3)]),
             ([rng.standard_normal((40, 1)), rng.standard_normal((30, 40)), rng.standard_normal((20, 30)), rng.standard_normal((1, 20))],
              [rng.standard_normal(40), rng.standard_normal(30), rng.standard_normal(20), rng.standard_normal(1)])]
    u = np.linspace(-1.5, 1.5, 30001)
    for ws, bs in cases:
        br, sl, ic = compile_value_layer(ws, bs, 0.1)
        assert np.all(np.diff(br) > 0) and len(sl) == len(br) + 1 == len(ic)
        h = u[:, None]
        for w, b in zip(ws[:-1], bs[:-1]):
            h = h @ np.asarray(w, np.float64).T + np.asarray(b, np.float64)
            h = np.where(h > 0, h, 0.1 * h)
        want = (h @ np.asarray(ws[-1], np.float64).T + np.asarray(bs[-1], np.float64))[:, 0]
        j = np.searchsorted(br, u, side="right")
        assert np.abs(sl[j] * u + ic[j] - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def test_n_imagenet_host_helpers():
    """pure host pieces of the N-ImageNet mirrors (imagenet.py:258-262, 1232-1260): no GPU needed"""
    import torch
    from event_representation_study_b200 import n_imagenet as N
    fake = N._empty_guard(torch.zeros((0, 4)))
    assert tuple(fake.shape) == (10, 4) and float(fake[:, 3].min()) == 1.0 and abs(float(fake[-1, 2]) - 0.9) < 1e-6
    keep = torch.ones((3, 4))
    assert N._empty_guard(keep) is keep
    # ImageNetDataset's loader_type dispatch (imagenet.py:1232-1272)
    assert N.loader_for("event_image") is N.reshape_then_acc and N.loader_for("reshape_then_acc_intensity") is N.reshape_then_acc_intensity
    assert N.loader_for("reshape_then_tore") is None and N.loader_for("nope") is None
    assert N.loader_for("sorted_time_surface") is N.reshape_then_acc_sort and N.loader_for("dist") is N.reshape_then_acc_adj_sort
    # the rank-based loaders' image-domain helpers (imagenet.py:567-583, 606-621)
    v = torch.tensor([[5.0, 0.0], [9.0, 5.0]], dtype=torch.float64)
    r = N._dense_rank_normalised(v, v > 0)
    assert r.dtype == torch.float32 and r.tolist() == [[0.0, 0.0], [1.0, 0.0]]
    assert N._dense_rank_normalised(v, v > 100).tolist() == [[0.0, 0.0], [0.0, 0.0]]
    q = N._quantize(torch.tensor([[0.3, 0.8]]), [4, 16])
    assert tuple(q.shape) == (1, 2, 2) and q[0, 0, 0] == 0.25 and q[0, 1, 1] == 0.8125
    with pytest.raises(RuntimeError):
        N._hot_check(torch.zeros(3, 3))


def test_new_entry_points_validate_before_touching_cuda(L):
    """argument checks of the round's new C entry points return EVREP_E* codes without a device"""
    import ctypes
    _lib = L
    L = _lib.lib
    assert L.evrep_gemm_workspace_bytes(0, 4, 4) == 0 and L.evrep_gemm_workspace_bytes(128, 128, 32) == 2 * 32768
    assert L.evrep_gw_kl_workspace_bytes(0, 5) == 0 and L.evrep_gw_kl_workspace_bytes(100, 100) > 0
    assert L.evrep_est_workspace_bytes(-1) == 0 and L.evrep_est_workspace_bytes(4) > 0
    assert L.evrep_image_pipeline_batched(None, 1, 8, 8, 12, 16, 7, 0, 255.0, 1 / 255.0, 114.0, 1, None, None) == _lib.EINVAL   # unknown mode
    assert L.evrep_image_pipeline_batched(None, 1, 8, 8, 12, 16, 0, 9, 255.0, 1 / 255.0, 114.0, 1, None, None) == _lib.EINVAL   # unknown interpolation
    assert L.evrep_assignment_auction(None, 4, 1e-9, None, None, None) == _lib.EINVAL
    d = ctypes.c_double(0.0)
    assert L.evrep_gw_kl(None, 4, 4, None, 4, 4, 0.7, 10, 1e-9, 1e-9, 0, ctypes.byref(d), None, None, None, None, 0, None) == _lib.EINVAL  # null arrays
    offs = (ctypes.c_int64 * 2)(0, 0)
    assert L.evrep_filter_batched(None, None, None, 4, None, offs, 1, 8, 8, 9, 1.0, 1, 1, None, None, None, 0, None) == _lib.EINVAL   # unknown filter
    assert L.evrep_est_quantize_batched(None, None, None, None, offs, 1, 8, 8, 1, None, None, None, 0, None, None, 0, None) == _lib.EINVAL  # C < 2
    # background-activity filter: radius outside 1..4, and the workspace query
    assert L.evrep_filter_background_workspace_bytes(1, 1000, 8, 8, 0, 4) == 0 and L.evrep_filter_background_workspace_bytes(1, 1000, 8, 8, 5, 8) == 0
    w1, w2 = L.evrep_filter_background_workspace_bytes(1, 1000, 8, 8, 1, 4), L.evrep_filter_background_workspace_bytes(1, 1000, 8, 8, 2, 4)
    assert 0 < w1 < w2
    assert L.evrep_filter_background_batched(None, None, None, 4, offs, 1, 8, 8, 10.0, 7, None, None, None, 0, None) == _lib.EINVAL


def test_evlicious_filter_module_mirrors_the_reference_names():
    """tools/filters.py:7-129: class names, enum values and the flag dispatch (host logic only)"""
    from event_representation_study_b200.evlicious.tools import filters as F
    assert [int(v) for v in F.Filtering_Type] == [1, 2, 3, 4, 5] and "HotPixel=5" in F.Filtering_Type.summary()
    mk = lambda **kw: type("Flags", (), kw)  # noqa: E731
    assert isinstance(F.from_flags(mk(filter_type=1, depth_us=10, radius=1)), F.BackgroundActivity)
    assert isinstance(F.from_flags(mk(filter_type=2, random_downsampling_factor=3)), F.Random)
    assert isinstance(F.from_flags(mk(filter_type=3, contrast_threshold_multiplier=2)), F.ContrastThresholdIncrease)
    assert isinstance(F.from_flags(mk(filter_type=4, depth_us=10)), F.RefractoryPeriod)
    assert isinstance(F.from_flags(mk(filter_type=5)), F.HotPixel)
    with pytest.raises(ValueError):
        F.from_flags(mk(filter_type=9))
    with pytest.raises(AssertionError):
        F.from_flags(mk(filter_type=1, depth_us=0, radius=1))


def test_evlicious_filter_objects_host_logic(monkeypatch):
    """the filter classes' own logic (lazy state, start mask, events[mask]) with the GPU calls replaced by the oracle loops"""
    from event_representation_study_b200.evlicious import Events
    from event_representation_study_b200.evlicious.tools import filters as F
    from event_representation_study_b200.evlicious.tools import utils as U
    from oracle import filters as ofil
    monkeypatch.setattr(U, "_background_activity_filter", ofil.background_activity_filter)
    monkeypatch.setattr(U, "_refractory_period", ofil.refractory_period)
    monkeypatch.setattr(U, "_contrast_threshold_control", ofil.contrast_threshold_control)
    rng = np.random.default_rng(3)
    H, W, n = 10, 12, 600
    x, y = rng.integers(0, W, n).astype(np.uint16), rng.integers(0, H, n).astype(np.uint16)
    t, p = np.cumsum(rng.integers(1, 40, n)).astype(np.int64), rng.choice(np.array([-1, 1], np.int8), n)
    mk = lambda lo, hi: Events(x[lo:hi].copy(), y[lo:hi].copy(), t[lo:hi].copy(), p[lo:hi].copy(), W, H)  # noqa: E731
    ba = F.BackgroundActivity(depth_us=100, radius=2)
    assert ba.timestamps is None
    kept = np.concatenate([ba.insert(mk(0, 250)).t, ba.insert(mk(250, n)).t])
    ts = np.full((H, W), -np.inf)
    want = ofil.background_activity_filter(np.ones(n, bool), ts, x, y, t, 100, 2)
    assert np.array_equal(kept, t[want]) and np.array_equal(ba.timestamps, ts) and ba.timestamps.dtype == np.float64
    rp = F.RefractoryPeriod(depth_us=300)
    last = np.full((H, W), -np.inf)
    want = ofil.refractory_period(np.ones(n, bool), x, y, t, 300, last)
    assert np.array_equal(rp.insert(mk(0, n)).t, t[want]) and np.array_equal(rp.timestamps, last)
    ct = F.ContrastThresholdIncrease(contrast_threshold_multiplier=2)
    act = np.zeros((H, W), np.int32)
    want = ofil.contrast_threshold_control(act, np.zeros(n, bool), x, y, p, 2)
    got = ct.insert(mk(0, n))
    assert np.array_equal(got.t, t[want]) and np.array_equal(ct.counter_map, act) and ct.counter_map.dtype == np.int32
    assert got.width == W and got.height == H and len(F.Random(3).insert(mk(0, 100))) == 33
    monkeypatch.setattr(F, "_pixel_counts", lambda ev: np.bincount(ev.y.astype(int) * W + ev.x.astype(int), minlength=H * W).reshape(H, W) * 1.0)
    xh = x.copy()
    xh[: n // 2] = 1
    yh = y.copy()
    yh[: n // 2] = 2
    hot = Events(xh, yh, t.copy(), p.copy(), W, H)
    hp = F.HotPixel()
    out = hp.insert(hot)
    assert np.array_equal(hp.hot_pixel_mask, ofil.hot_pixel_mask(xh, yh, H, W)) and not hp.hot_pixel_mask[2, 1] and hp.hot_pixel_mask.sum() == H * W - 1
    assert len(out) == int((~((xh == 1) & (yh == 2))).sum())


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """include/evrep.h is the drop-in boundary: it must compile as C99 (no C++ / torch types) and a C caller must link and
    run against libevrep.so (argument checks only - no device here)"""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "event_representation_study_b200", "lib")
    src = tmp_path / "caller.c"
    src.write_text('''#include <stdio.h>
#include <string.h>
#include "evrep.h"
int main(void) {
  long long offs[2] = {0, 10};
  if (evrep_version() != 100) return 1;
  if (evrep_ergo12_batched(NULL, NULL, NULL, 4, NULL, (const int64_t*)offs, 1, 30, 40, 2, NULL, NULL, 0, NULL) != EVREP_EINVAL) return 2;
  if (!strstr(evrep_last_error(), "null")) return 3;
  if (evrep_workspace_bytes(EVREP_OP_MIXED_DENSITY, 1, 1000, 240, 304, 12) == 0) return 4;
  if (evrep_filter_background_workspace_bytes(1, 1000, 8, 8, 1, 4) == 0) return 5;
  {
    /* a plain C caller compiles kernels for its own tuple (NVRTC runs on the host: no device needed for this step) */
    const int8_t win[4] = {0, 3, 6, 1}, func[4] = {EVREP_FUNC_COUNT, EVREP_FUNC_TIMESTAMP_NEG, EVREP_FUNC_POLARITY, EVREP_FUNC_TIMESTAMP_POS};
    const int8_t agg[4] = {EVREP_AGG_SUM, EVREP_AGG_VARIANCE, EVREP_AGG_MEAN, EVREP_AGG_MAX};
    size_t nbytes = 0;
    int rc = evrep_mixed_density_specialize_compile_only(win, func, agg, 4, EVREP_STACK_SBN, 200000, &nbytes);
    if (rc != EVREP_OK && rc != EVREP_EUNSUPPORTED) return 6;   /* EUNSUPPORTED only where libnvrtc is not installed */
    if (rc == EVREP_OK && (nbytes == 0 || !evrep_mixed_density_is_specialized(win, func, agg, 4, EVREP_STACK_SBN, 1000))) return 7;
  }
  puts("ok");
  return 0;
}
''')
    exe = tmp_path / "caller"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-L", libdir, "-levrep",
                    f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True, capture_output=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)


def test_split_collated_raw_event_batch():
    """the reference's collate for raw events, rows [x, y, t, p, b] (gen1_2yolo.py:433-445) -> SoA fields + CSR offsets"""
    import torch
    from event_representation_study_b200 import batched as eb
    rng = np.random.default_rng(8)
    samples = [np.stack([rng.integers(0, 304, n), rng.integers(0, 240, n), np.sort(rng.integers(0, 90_000, n)), rng.choice([-1, 1], n)], 1).astype(np.float32)
               for n in (5, 0, 3, 7)]
    col = np.concatenate([np.concatenate([d, i * np.ones((len(d), 1), np.float32)], 1) for i, d in enumerate(samples)], 0)  # what collate_fn does
    x, y, t, p, offs = eb.split_collated(torch.from_numpy(col))
    assert offs.tolist() == [0, 5, 5, 8, 15] and x.dtype == torch.int16 and t.dtype == torch.int32 and p.dtype == torch.int8
    cat = np.concatenate(samples, 0)
    assert np.array_equal(x.numpy().view(np.uint16), cat[:, 0].astype(np.uint16)) and np.array_equal(y.numpy(), cat[:, 1].astype(np.int16))
    assert np.array_equal(t.numpy(), cat[:, 2].astype(np.int64)) and np.array_equal(p.numpy(), cat[:, 3].astype(np.int8))
    assert eb.split_collated(torch.from_numpy(col), num_windows=6)[4].tolist() == [0, 5, 5, 8, 15, 15, 15]  # trailing empty samples
    assert eb.split_collated(torch.zeros((0, 5)))[4].tolist() == [0]
    big = col.astype(np.float64).copy()
    big[:, 2] += 3e9
    assert eb.split_collated(torch.from_numpy(big))[2].dtype == torch.int64
    wide = col.copy()
    wide[0, 0] = 40000  # a uint16 value above int16: kept bit for bit
    assert int(eb.split_collated(torch.from_numpy(wide))[0].numpy().view(np.uint16)[0]) == 40000
    with pytest.raises(ValueError):
        eb.split_collated(torch.from_numpy(col[::-1].copy()))       # not grouped by ascending b
    with pytest.raises(ValueError):
        eb.split_collated(torch.zeros((4, 4)))
    with pytest.raises(ValueError):
        eb.split_collated(torch.from_numpy(col), num_windows=2)
    bad = col.copy()
    bad[1, 3] = 2
    with pytest.raises(ValueError):
        eb.split_collated(torch.from_numpy(bad))
    with pytest.raises(ValueError):
        eb.from_collated(torch.from_numpy(col), device="cpu")      # an EventBatch lives on the GPU: there is no CPU path


def test_bind_to_gpu_cpus_never_raises():
    """rank placement is an optimisation: without a GPU / NVML it reports why it was skipped and leaves the mask alone"""
    import os
    from event_representation_study_b200.sharding import bind_to_gpu_cpus
    before = os.sched_getaffinity(0)
    rep = bind_to_gpu_cpus(0)
    assert isinstance(rep, dict) and ("skipped" in rep or rep["cpus"] >= 1)
    if "skipped" in rep:
        assert os.sched_getaffinity(0) == before


def test_otmi_prepare_validates_before_touching_cuda(L):
    """evrep_otmi_prepare: argument checks return EVREP_E* codes without a device"""
    lib = L.lib
    assert lib.evrep_otmi_workspace_bytes(-1, 64) == 0 and lib.evrep_otmi_workspace_bytes(1000, 0) == 0
    assert lib.evrep_otmi_workspace_bytes(50000, 240) > 0
    assert lib.evrep_otmi_prepare(0, 0, 10, 0, 64, 3, 60, 80, 0, 10, 0, 2000, 0, 0, 0, 0) == L.EINVAL  # null pointers
    assert b"null" in lib.evrep_last_error()
