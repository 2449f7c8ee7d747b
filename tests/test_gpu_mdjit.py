"""Run-time specialised mixed-density kernels (evrep_mixed_density_specialize: NVRTC-compiled instances of the ERGO-12 kernel
templates for an arbitrary tuple) against the reference-generated fixtures, the numpy oracle and the interpreted kernel.
Reference: representations/representation_search/mixed_density_event_stack.py:25-151, operations.py:15-89."""
import numpy as np
import pytest

from conftest import assert_close, golden, load

pytestmark = pytest.mark.gpu

RTOL, VAR_ATOL = 1e-5, 2e-7  # the bars of test_gpu_parity.py
INT_FUNCS = {"count", "count_pos", "count_neg", "polarity"}


@pytest.fixture(scope="module")
def E(cuda_device):
    import event_representation_study_b200.batched as eb
    return eb


def np_(t):
    return t.detach().cpu().numpy()


def batch_of(E, gs):
    wins = [{"x": g["x"], "y": g["y"], "t": np.asarray(g["t"]).astype(np.int64), "p": g["p"]} for g in gs]
    return E.pack_events(wins, "cuda")


def integer_channels(funcs, aggs):
    return [c for c, (f, a) in enumerate(zip(funcs, aggs)) if f in INT_FUNCS and a in ("sum", "max", "min")]


MDES = [c for c in golden("mdes_*") + golden("mdmin_*")]


@pytest.mark.parametrize("name,path", MDES, ids=[c[0] for c in MDES])
def test_specialized_golden(E, name, path):
    g = load(path)
    win, func, agg, st = g["win"].tolist(), g["func"].tolist(), g["agg"].tolist(), str(g["stacking"])
    assert E.specialize_mixed_density(win, func, agg, st, max_events_per_window=1 << 20)
    assert E.mixed_density_is_specialized(win, func, agg, st, len(g["x"]))
    out = np_(E.mixed_density(batch_of(E, [g]), int(g["H"]), int(g["W"]), win, func, agg, st))[0]
    assert_close(out, g["out"], rtol=RTOL, atol=VAR_ATOL, what=name)


def test_specialized_all_pairs_vs_oracle(E):
    """Every (function, aggregation) pair on every SBN and SBT window, mixed {-1,0,+1} polarities, through specialised kernels."""
    from oracle import representations as orep
    from event_representation_study_b200.synth import poisson_window
    H, W = 48, 64
    w = poisson_window(31, 30_000, H, W)
    rng = np.random.default_rng(5)
    w["p"] = rng.integers(-1, 2, len(w["p"])).astype(np.int8)
    w["p"][: len(w["p"]) // 3] = np.abs(w["p"][: len(w["p"]) // 3])  # first third holds no -1: the p == 0 fallback
    ev = E.pack_events([w, poisson_window(32, 777, H, W)], "cuda")
    for st, nwin in (("SBN", 7), ("SBT", 8)):
        for win in range(nwin):
            full = [(win, f, a) for f in orep.FUNCTIONS for a in orep.AGGREGATIONS]  # 35 channels
            for spec in (full[:20], full[19:]):                                        # 20 + 16
                wi, fu, ag = [s[0] for s in spec], [s[1] for s in spec], [s[2] for s in spec]
                assert E.specialize_mixed_density(wi, fu, ag, st, max_events_per_window=30_000)
                out = np_(E.mixed_density(ev, H, W, wi, fu, ag, st))[0]
                with np.errstate(all="ignore"):
                    want = orep.mixed_density_event_stack(w["x"], w["y"], w["t"], w["p"], H, W, wi, fu, ag, st)
                assert_close(out, want, rtol=RTOL, atol=VAR_ATOL, what=f"{st} window {win}")


def random_tuple(seed, C=12):
    import random
    rng = random.Random(seed)
    F = ["timestamp", "polarity", "count", "timestamp_pos", "timestamp_neg", "count_pos", "count_neg"]
    A = ["sum", "mean", "max", "variance"]
    return [rng.randrange(7) for _ in range(C)], [rng.choice(F) for _ in range(C)], [rng.choice(A) for _ in range(C)]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_specialized_vs_oracle_at_the_headline_size(E, seed):
    """A random search tuple on a 1 M-event window at 1280 x 720 (the benchmarked size) against the oracle, and against the
    interpreted kernel on a ragged batch: integer-valued channels bit exact."""
    from oracle import representations as orep
    from event_representation_study_b200.synth import poisson_window
    H, W = 720, 1280
    wi, fu, ag = random_tuple(seed)
    wins = [poisson_window(5000 + seed, 1_000_000, H, W, clustered=(seed == 2)), poisson_window(5100 + seed, 40_001, H, W, polarity="01")]
    ev = E.pack_events(wins, "cuda")
    generic = np_(E.mixed_density(ev, H, W, wi, fu, ag, "SBN")) if not E.mixed_density_is_specialized(wi, fu, ag, "SBN", 1_000_000) else None
    assert E.specialize_mixed_density(wi, fu, ag, "SBN", max_events_per_window=1_000_000)
    out = np_(E.mixed_density(ev, H, W, wi, fu, ag, "SBN"))
    assert (E.window_flags(ev) == 0).all()
    ints = integer_channels(fu, ag)
    for i, w in enumerate(wins):
        with np.errstate(all="ignore"):
            want = orep.mixed_density_event_stack(w["x"], w["y"], w["t"], w["p"], H, W, wi, fu, ag, "SBN")
        assert_close(out[i], want, rtol=RTOL, atol=VAR_ATOL, what=f"tuple {seed} window {i}")
        for c in ints:
            assert np.array_equal(out[i][:, :, c], want[:, :, c].astype(np.float32)), (seed, i, c)
    if generic is not None:
        assert_close(out, generic, rtol=RTOL, atol=VAR_ATOL, what="specialised vs interpreted")
        for c in ints:
            assert np.array_equal(out[..., c], generic[..., c])


@pytest.mark.parametrize("many_sums", [False, True], ids=["1024-pixel-tiles", "512-pixel-tiles"])
def test_specialized_hot_tile(E, many_sums):
    """More than 65535 events inside one 1024-pixel tile: the specialised packed plan hands the bucket to the specialised wide
    plan; with eight timestamp-variance groups the plans need 512-pixel tiles."""
    from oracle import representations as orep
    from event_representation_study_b200.synth import poisson_window
    H, W = 64, 64
    w = poisson_window(77, 180_000, H, W)
    w["x"] = (w["x"] % 8).astype(np.uint16)
    hot = np.arange(len(w["x"])) % 3 == 0
    w["x"][hot] = 3
    w["y"][hot] = 5  # one pixel alone receives 60k events
    if many_sums:
        wi = [0, 1, 2, 3, 4, 5, 6, 0, 0, 1, 2, 3]
        fu = ["timestamp_pos"] * 7 + ["timestamp_neg"] + ["count", "polarity", "count_pos", "timestamp"]
        ag = ["variance"] * 8 + ["sum", "mean", "sum", "max"]
    else:
        wi, fu, ag = random_tuple(11)
    assert E.specialize_mixed_density(wi, fu, ag, "SBN", max_events_per_window=180_000)
    out = np_(E.mixed_density(E.pack_events([w, poisson_window(78, 5000, H, W)], "cuda"), H, W, wi, fu, ag, "SBN"))
    with np.errstate(all="ignore"):
        want = orep.mixed_density_event_stack(w["x"], w["y"], w["t"], w["p"], H, W, wi, fu, ag, "SBN")
    assert_close(out[0], want, rtol=RTOL, atol=VAR_ATOL, what="hot tile")
    for c in integer_channels(fu, ag):
        assert np.array_equal(out[0][:, :, c], want[:, :, c].astype(np.float32))


def test_outside_the_envelope_keeps_the_interpreted_kernel(E):
    from oracle import representations as orep
    from event_representation_study_b200.synth import poisson_window
    H, W = 30, 40
    w = poisson_window(9, 5000, H, W)
    ev = E.pack_events([w], "cuda")
    wi, fu, ag = random_tuple(21)
    assert not E.specialize_mixed_density(wi, fu, ag, "no such stacking")  # the reference builds window 0 only for it
    assert E.specialize_mixed_density(wi[:7], fu[:7], ag[:7], "SBN")        # 7 channels: compiled, but 30 * 41 * 7 is odd ...
    for spec, st, hw in (((wi, fu, ag), "no such stacking", (30, 40)), ((wi[:7], fu[:7], ag[:7]), "SBN", (30, 41))):  # ... so this call is interpreted
        H, W = hw
        w = poisson_window(9, 5000, H, W)
        ev = E.pack_events([w], "cuda")
        out = np_(E.mixed_density(ev, H, W, *spec, st))[0]
        with np.errstate(all="ignore"):
            want = orep.mixed_density_event_stack(w["x"], w["y"], w["t"], w["p"], H, W, *spec, st)
        assert_close(out, want, rtol=RTOL, atol=VAR_ATOL, what=st)
    # swallowed list entries (unknown names / window indices) are zero channels in specialised kernels too
    H, W = 30, 40
    w = poisson_window(9, 5000, H, W)
    ev = E.pack_events([w], "cuda")
    fu2 = list(fu); fu2[3] = "no_such_function"
    wi2 = list(wi); wi2[5] = 9
    assert E.specialize_mixed_density(wi2, fu2, ag, "SBN")
    out = np_(E.mixed_density(ev, H, W, wi2, fu2, ag, "SBN"))[0]
    assert not out[:, :, 3].any() and not out[:, :, 5].any()


def test_mirror_class_specialises_in_the_background(E):
    """The drop-in class hands its tuple to a background compilation after SPECIALIZE_AFTER_CALLS stack() calls: no call waits,
    later calls run the specialised kernels, and every result stays the reference's."""
    import time
    from oracle import representations as orep
    from event_representation_study_b200.synth import poisson_window, structured
    from event_representation_study_b200.representations.representation_search import mixed_density_event_stack as M
    H, W = 30, 40
    wi, fu, ag = random_tuple(33)
    w = poisson_window(10, 3000, H, W)
    rec = structured(w)
    with np.errstate(all="ignore"):
        want = orep.mixed_density_event_stack(w["x"], w["y"], w["t"], w["p"], H, W, wi, fu, ag, "SBN")
    old = M.SPECIALIZE_AFTER_CALLS
    M.SPECIALIZE_AFTER_CALLS = 3
    try:
        for k in range(3):
            assert not E.mixed_density_is_specialized(wi, fu, ag, "SBN", 3000)
            rep = M.MixedDensityEventStack(12, len(rec), H, W, (wi, fu, ag), "SBN").stack(rec)
            assert_close(rep, want, rtol=RTOL, atol=VAR_ATOL, what=f"call {k}")
        t0 = time.time()
        while not E.mixed_density_is_specialized(wi, fu, ag, "SBN", 3000):
            assert time.time() - t0 < 60, "background compilation did not finish"
            time.sleep(0.05)
        rep = M.MixedDensityEventStack(12, len(rec), H, W, (wi, fu, ag), "SBN").stack(rec)
        assert_close(rep, want, rtol=RTOL, atol=VAR_ATOL, what="after the switch")
    finally:
        M.SPECIALIZE_AFTER_CALLS = old


def test_specialised_call_replays_as_a_cuda_graph(E):
    """The driver-API launches of the specialised kernels are captured like the runtime launches of the other kernels."""
    import torch
    from event_representation_study_b200.synth import poisson_window
    H, W = 120, 160
    wi, fu, ag = random_tuple(44)
    ev = E.pack_events([poisson_window(50 + i, 20_000 + 17 * i, H, W) for i in range(4)], "cuda")
    assert E.specialize_mixed_density(wi, fu, ag, "SBN", max_events_per_window=30_000)
    out = torch.empty((4, H, W, 12), device="cuda")
    eager = E.mixed_density(ev, H, W, wi, fu, ag, "SBN").clone()
    call = E.GraphedCall(lambda: E.mixed_density(ev, H, W, wi, fu, ag, "SBN", out=out))
    out.zero_()
    call.replay()
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(out, nan=-7.0), torch.nan_to_num(eager, nan=-7.0))


def test_large_batches_ask_for_their_kernels_by_themselves(E):
    """mixed_density(specialize="auto"): a batch of 4 M events or more starts a background compilation of its tuple; the call that
    asked is served by the interpreted kernel, a later one by the specialised kernels, with the same result."""
    import time
    from event_representation_study_b200.synth import device_batch
    import torch
    H, W = 240, 304
    wi, fu, ag = random_tuple(55)
    d = device_batch(32, 150_000, H, W, torch.device("cuda", 0), seed=8)
    ev = E.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
    assert not E.mixed_density_is_specialized(wi, fu, ag, "SBN", 150_000)
    first = E.mixed_density(ev, H, W, wi, fu, ag, "SBN").clone()
    t0 = time.time()
    while not E.mixed_density_is_specialized(wi, fu, ag, "SBN", 150_000):
        assert time.time() - t0 < 60, "background compilation did not finish"
        time.sleep(0.05)
    second = E.mixed_density(ev, H, W, wi, fu, ag, "SBN")
    a, b = np_(first), np_(second)
    assert_close(b, a, rtol=RTOL, atol=VAR_ATOL, what="specialised vs interpreted")
    for c in integer_channels(fu, ag):
        assert np.array_equal(a[..., c], b[..., c])
