"""Host side of the run-time specialised mixed-density kernels: NVRTC turns a tuple into an sm_100a image without a GPU
(evrep_mixed_density_specialize_compile_only), tuples outside the envelope are refused, nothing is registered as
specialised until it has been compiled."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from event_representation_study_b200 import _lib

lib = _lib.lib


def _nvrtc_available():
    w = np.zeros(4, np.int8)
    f = np.full(4, 2, np.int8)
    nb = ctypes.c_size_t(0)
    rc = lib.evrep_mixed_density_specialize_compile_only(w.ctypes.data, f.ctypes.data, w.ctypes.data, 4, 0, 1000, ctypes.byref(nb))
    return not (rc == _lib.EUNSUPPORTED and b"unavailable" in lib.evrep_last_error())


pytestmark = pytest.mark.skipif(not _nvrtc_available(), reason="libnvrtc.so.12 (CUDA toolkit runtime compiler) is not installed on this machine")


def codes(wi, fu, ag):
    return (np.array(wi, np.int8), np.array([_lib.FUNCS[f] for f in fu], np.int8), np.array([_lib.AGGS[a] for a in ag], np.int8))


def compile_only(wi, fu, ag, stacking=0, n_max=1 << 20):
    w, f, a = codes(wi, fu, ag)
    nb = ctypes.c_size_t(0)
    rc = lib.evrep_mixed_density_specialize_compile_only(w.ctypes.data, f.ctypes.data, a.ctypes.data, len(wi), stacking, n_max, ctypes.byref(nb))
    return rc, nb.value


TUPLE = ([2, 1, 3, 5, 0, 0, 6, 4, 0, 2, 4, 1],
         ["timestamp", "polarity", "count", "timestamp_pos", "timestamp_neg", "count_pos", "count_neg", "polarity", "timestamp", "count", "timestamp_neg", "count_pos"],
         ["variance", "mean", "sum", "max", "mean", "mean", "sum", "variance", "sum", "max", "min", "mean"])


def test_nvrtc_compiles_a_tuple_without_a_gpu(tmp_path):
    dump = tmp_path / "tuple.cubin"
    os.environ["EVREP_JIT_DUMP"] = str(dump)
    try:
        rc, nbytes = compile_only(*TUPLE)
    finally:
        del os.environ["EVREP_JIT_DUMP"]
    assert rc == _lib.OK, lib.evrep_last_error().decode()
    assert nbytes > 10_000
    w, f, a = codes(*TUPLE)
    # compiled, hence known to the library: a later evrep_mixed_density_batched with this tuple would take these kernels
    assert lib.evrep_mixed_density_is_specialized(w.ctypes.data, f.ctypes.data, a.ctypes.data, 12, 0, 1000) == 1
    assert lib.evrep_mixed_density_is_specialized(w.ctypes.data, f.ctypes.data, a.ctypes.data, 12, 0, 1 << 23) == 0  # needs a narrower limb than compiled
    if dump.exists():
        res = subprocess.run(["cuobjdump", "-res-usage", str(dump)], capture_output=True, text=True)
        if res.returncode == 0:
            assert "k_md_tile_static" in res.stdout and "k_md_tile_heavy" in res.stdout


def test_second_compilation_is_served_from_the_cache():
    import time
    compile_only(*TUPLE)
    t0 = time.time()
    rc, nbytes = compile_only(*TUPLE)
    assert rc == _lib.OK and nbytes > 0 and time.time() - t0 < 0.2


def test_sbt_tuples_compile_too():
    rc, nbytes = compile_only(*TUPLE, stacking=1)
    assert rc == _lib.OK and nbytes > 10_000, lib.evrep_last_error().decode()


@pytest.mark.parametrize("case", ["too-many-sums"])
def test_outside_the_envelope_is_refused(case):
    wi, fu, ag = TUPLE
    if True:  # 28 distinct timestamp-variance groups: the packed plan alone exceeds a tile's shared memory
        wi = [k % 7 for k in range(28)]
        fu = ["timestamp", "timestamp_pos", "timestamp_neg", "timestamp"][:1] * 7 + ["timestamp_pos"] * 7 + ["timestamp_neg"] * 7 + ["timestamp"] * 7
        rc, _ = compile_only(wi, fu, ["variance"] * 28)
    assert rc == _lib.EUNSUPPORTED
    assert lib.evrep_last_error()


def test_ergo_tuples_need_no_compilation():
    """evrep_mixed_density_specialize on an ERGO-12 tuple returns at once: those kernels are built ahead of time."""
    wi = [0, 3, 2, 6, 5, 6, 2, 5, 1, 0, 4, 1]
    fu = ["polarity", "timestamp_neg", "count_neg", "polarity", "count_pos", "count", "timestamp_pos", "count_neg", "timestamp_neg", "timestamp_pos", "timestamp", "count"]
    ag = ["variance", "variance", "mean", "sum", "mean", "sum", "mean", "mean", "max", "max", "max", "mean"]
    w, f, a = codes(wi, fu, ag)
    assert lib.evrep_mixed_density_specialize(w.ctypes.data, f.ctypes.data, a.ctypes.data, 12, 0, 1 << 20) == _lib.OK  # no device needed


def test_background_compilation_finishes_without_a_gpu():
    import time
    wi = [6, 5, 4, 3, 2, 1, 0, 6, 5, 4, 3, 2]
    w, f, a = codes(wi, TUPLE[1], TUPLE[2])
    args = (w.ctypes.data, f.ctypes.data, a.ctypes.data, 12, 0, 100_000)
    assert lib.evrep_mixed_density_is_specialized(*args) == 0
    t0 = time.time()
    assert lib.evrep_mixed_density_specialize_async(*args) == _lib.OK
    assert time.time() - t0 < 0.2, "the asynchronous call must not wait for NVRTC"
    assert lib.evrep_mixed_density_specialize_async(*args) == _lib.OK  # already pending: no second compilation
    while lib.evrep_mixed_density_is_specialized(*args) == 0:
        assert time.time() - t0 < 120
        time.sleep(0.05)


def test_mirror_class_host_logic_with_the_oracle_patched_in(monkeypatch):
    """MixedDensityEventStack (the drop-in class): the background specialisation is requested exactly once per tuple, at the
    SPECIALIZE_AFTER_CALLS-th stack() call counted across instances (the reference builds one instance per sample,
    optimized_representation.py:131-134), for SBN and SBT but not for a stacking type the reference does not know; the
    GPU calls are replaced by the numpy oracle here."""
    import torch
    from oracle import representations as orep
    from event_representation_study_b200.representations.representation_search import mixed_density_event_stack as M
    from event_representation_study_b200.synth import poisson_window, structured

    H, W = 12, 16
    asked = []

    class FakeBatch:
        def __init__(self, w):
            self.w = w

    def fake_one_window_structured(records, height, width):
        return FakeBatch({k: np.asarray(records[k]) for k in "xytp"})

    def fake_mixed_density(ev, height, width, wi, fu, ag, stacking="SBN", out=None, specialize="auto"):
        w = ev.w
        with np.errstate(all="ignore"):
            rep = orep.mixed_density_event_stack(w["x"], w["y"], w["t"], w["p"], height, width, wi, fu, ag, stacking)
        return torch.as_tensor(rep[None].astype(np.float32))

    monkeypatch.setattr(M, "one_window_structured", fake_one_window_structured)
    monkeypatch.setattr(M.eb, "mixed_density", fake_mixed_density)
    monkeypatch.setattr(M.eb, "specialize_mixed_density", lambda w, f, a, st, max_events_per_window=0, wait=True: asked.append((st, tuple(w), wait)) or True)
    monkeypatch.setattr(M, "to_host", lambda t, dtype, scale=None: t.to(dtype).numpy() * (1.0 if scale is None else scale))
    monkeypatch.setattr(M, "SPECIALIZE_AFTER_CALLS", 4)
    monkeypatch.setattr(M, "_TUPLE_CALLS", {})
    rec = structured(poisson_window(3, 500, H, W))
    spec = ([0, 1, 2, 3], ["count", "polarity", "timestamp", "count_pos"], ["sum", "mean", "max", "sum"])
    for st in ("SBN", "SBT", "no such stacking"):
        for k in range(9):
            rep = M.MixedDensityEventStack(4, len(rec), H, W, spec, st).stack(rec)
            assert rep.shape == (H, W, 4) and rep.dtype == np.float64
    assert asked == [("SBN", (0, 1, 2, 3), False), ("SBT", (0, 1, 2, 3), False)]
    monkeypatch.setattr(M, "SPECIALIZE_AFTER_CALLS", None)  # switched off
    monkeypatch.setattr(M, "_TUPLE_CALLS", {})
    for k in range(6):
        M.MixedDensityEventStack(4, len(rec), H, W, spec, "SBN").stack(rec)
    assert len(asked) == 2


FUN = ["timestamp", "polarity", "count", "timestamp_pos", "timestamp_neg", "count_pos", "count_neg"]
AGG = ["sum", "mean", "max", "variance", "min"]


def _raw_compile(w, f, a, stacking=0, n_max=1 << 20):
    w, f, a = np.array(w, np.int8), np.array(f, np.int8), np.array(a, np.int8)
    nb = ctypes.c_size_t(0)
    rc = lib.evrep_mixed_density_specialize_compile_only(w.ctypes.data, f.ctypes.data, a.ctypes.data, len(w), stacking, n_max, ctypes.byref(nb))
    return rc, nb.value


@pytest.mark.parametrize("name,w,f,a,stacking", [
    ("one-channel", [0], [2], [0], 0),
    ("one-channel-sbt", [7], [0], [3], 1),
    ("counts-only", [0, 1, 2, 3, 4, 5, 6, 0], [2, 5, 6, 2, 5, 6, 2, 5], [0, 0, 0, 1, 1, 2, 2, 3], 0),
    ("latest-and-earliest", [0, 1, 2, 3, 4, 5], [0, 3, 4, 0, 3, 4], [2, 2, 2, 4, 4, 4], 0),
    ("all-swallowed", [9, 0, 0, 3], [0, 9, 0, -1], [0, 0, 9, 1], 0),
    ("polarity-every-aggregation", [0, 1, 2, 3, 4], [1, 1, 1, 1, 1], [0, 1, 2, 3, 4], 0),
    ("negative-window-index", [-1, -7, 3], [0, 2, 1], [1, 0, 3], 0),
    ("thirty-two-channels", [c % 7 for c in range(32)], [[2, 5, 6, 1][c % 4] for c in range(32)], [c % 3 for c in range(32)], 0),
    ("huge-windows", [0, 3, 2, 6], [0, 4, 6, 1], [3, 3, 1, 0], 0),
])
def test_unusual_tuples_compile(name, w, f, a, stacking):
    """The kernel templates must instantiate for every plan shape a caller can ask for - one channel, zero channels left after the
    reference's swallow-and-zero, presence-only plans, earliest / latest stamps, 32 channels, SBT, limb widths down to 8 bits."""
    rc, nbytes = _raw_compile(w, f, a, stacking, n_max=(1 << 23) if name == "huge-windows" else (1 << 20))
    assert rc == _lib.OK and nbytes > 0, (name, lib.evrep_last_error().decode()[:2000])


def test_random_tuples_compile():
    import random
    rng = random.Random(123)
    for k in range(6):
        C = rng.choice([3, 6, 9, 12, 12, 16])
        w = [rng.randrange(8 if k % 2 else 7) for _ in range(C)]
        f = [rng.randrange(7) for _ in range(C)]
        a = [rng.randrange(5) for _ in range(C)]
        rc, nbytes = _raw_compile(w, f, a, k % 2)
        if rc == _lib.EUNSUPPORTED:  # a plan too wide for a tile's shared memory: refused, never a compiler error
            assert b"does not fit" in lib.evrep_last_error() or b"accumulator words" in lib.evrep_last_error()
            continue
        assert rc == _lib.OK and nbytes > 0, (w, f, a, lib.evrep_last_error().decode()[:2000])
