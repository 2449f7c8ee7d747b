"""Packed host wire format (event_representation_study_b200/packed.py, evrep_unpack_events): the numpy packer against its numpy
decoder on the CPU, the CUDA decode kernel against both on the GPU, and ERGO-12 from a packed batch against the SoA batch."""
import numpy as np
import pytest


def _batch(sizes, H, W, seed=0, polarity="pm1", duration_us=300_000):
    from event_representation_study_b200.synth import pack_batch, poisson_window
    wins = [poisson_window(seed + i, n, H, W, polarity=polarity, duration_us=duration_us) for i, n in enumerate(sizes)]
    for i, w in enumerate(wins):
        w["t"] = w["t"] + 1_000_000 * (i + 1)  # absolute stamps: the format stores them relative to the window's first event
    return wins, pack_batch(wins)


@pytest.mark.parametrize("sizes,H,W,fmt,dur", [([5000, 0, 64, 65, 1, 12345], 240, 304, 4, 2_000), ([5000, 0, 256, 257, 1, 12345], 240, 304, 6, 50_000),
                                               ([20000, 777], 720, 1280, None, 2_000), ([3000], 4000, 5000, None, 300_000),
                                               ([5000, 0, 64, 65, 1, 12345, 63, 129], 720, 1280, 3, 2_000), ([4000, 130], 240, 304, 3, 3_000_000)])
def test_pack_roundtrip_on_the_host(sizes, H, W, fmt, dur):
    from event_representation_study_b200 import packed
    wins, b = _batch(sizes, H, W, 3, duration_us=dur)
    pk = packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=fmt)
    assert pk is not None and pk.fmt == (fmt or (3 if W <= 2048 else 6))  # smallest that fits: 3 needs x, y within 21 bits
    x, y, t, p = packed.unpack_numpy(pk)
    assert np.array_equal(x, b["x"]) and np.array_equal(y, b["y"]) and np.array_equal(p, b["p"])
    n = np.diff(b["offsets"])
    first = np.repeat(b["t"].astype(np.int64)[b["offsets"][:-1][n > 0]], n[n > 0])
    assert np.array_equal(t, b["t"].astype(np.int64) - first)
    per_event = pk.nbytes / max(1, int(b["offsets"][-1]))
    assert per_event < {3: 3.3, 4: 4.2, 6: 6.1}[pk.fmt] or int(b["offsets"][-1]) < 5000 or dur > 1_000_000  # sparse streams: escapes


def test_pack_falls_back_when_a_block_spans_too_long():
    from event_representation_study_b200 import packed
    wins, b = _batch([4000], 240, 304, 5, duration_us=300_000_000)  # 75 ms between events: no block of 64 fits 2^13 us
    assert packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], 240, 304, fmt=4) is None
    pk = packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], 240, 304)
    assert pk is not None and pk.fmt == 3  # the delta format takes any gap through its escape table (12 B per escaped event)
    wins, b = _batch([300], 240, 304, 6, duration_us=2_000_000_000, polarity="01")  # p == 0 events: not format 3; gaps too long for 4 and 6
    assert packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], 240, 304) is None  # caller uploads the SoA arrays
    wins, b = _batch([300], 240, 304, 6)
    b["t"][10], b["t"][11] = b["t"][11], b["t"][10] + 0  # unsorted: negative difference
    if b["t"][10] != b["t"][11]:
        assert packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], 240, 304, fmt=3) is None
    with pytest.raises(IndexError):
        packed.pack_host(np.array([400]), np.array([0]), np.array([0]), np.array([1]), np.array([0, 1]), 240, 304)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [4, 6])
@pytest.mark.parametrize("polarity", ["pm1", "01"])
def test_decode_kernel_matches_the_host_decoder(cuda_device, fmt, polarity):
    from event_representation_study_b200 import packed
    H, W = 720, 1280
    # dense enough for the 9 offset bits format 4 has at this sensor size (a block of 64 events may span 511 us)
    wins, b = _batch([100_000, 0, 1, 63, 64, 65, 1023, 1024, 1025, 4097, 250_001], H, W, 11, polarity=polarity, duration_us=400 if fmt == 4 else 40_000)
    pk = packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=fmt, pin=True)
    ev = packed.upload(pk, "cuda")
    x, y, t, p = packed.unpack_numpy(pk)
    assert np.array_equal(ev.x.cpu().numpy().view(np.uint16), x) and np.array_equal(ev.y.cpu().numpy().view(np.uint16), y)
    assert np.array_equal(ev.t.cpu().numpy().astype(np.int64), t) and np.array_equal(ev.p.cpu().numpy(), p)


@pytest.mark.gpu
@pytest.mark.parametrize("dur", [300, 40_000, 30_000_000])
def test_delta_decode_kernel_matches_the_host_decoder(cuda_device, dur):
    """format 3: dense streams (codes 0..2 only), ordinary ones, and sparse ones where nearly every event is an escape; also a
    group of windows decoded from slices of the batch's tables (what the end-to-end leg of bench.py ships per group)"""
    from event_representation_study_b200 import packed
    H, W = 720, 1280
    wins, b = _batch([100_000, 0, 1, 63, 64, 65, 1023, 1024, 1025, 4097, 250_001], H, W, 11, duration_us=dur)
    pk = packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=3, pin=True)
    assert pk is not None and pk.fmt == 3
    x, y, t, p = packed.unpack_numpy(pk)
    n = np.diff(b["offsets"])
    first = np.repeat(b["t"].astype(np.int64)[b["offsets"][:-1][n > 0]], n[n > 0])
    assert np.array_equal(x, b["x"]) and np.array_equal(y, b["y"]) and np.array_equal(p, b["p"]) and np.array_equal(t, b["t"].astype(np.int64) - first)
    ev = packed.upload(pk, "cuda")
    assert np.array_equal(ev.x.cpu().numpy().view(np.uint16), x) and np.array_equal(ev.y.cpu().numpy().view(np.uint16), y)
    assert np.array_equal(ev.t.cpu().numpy().astype(np.int64), t) and np.array_equal(ev.p.cpu().numpy(), p)
    w0, w1 = 4, 11  # a group: slices of the tables, escape prefixes not starting at 0
    parts = {k: v.to("cuda") for k, v in pk.host_parts(w0, w1).items()}
    sub = pk.decode_parts(parts, b["offsets"][w0:w1 + 1] - b["offsets"][w0])
    e0, e1 = int(b["offsets"][w0]), int(b["offsets"][w1])
    assert np.array_equal(sub.t.cpu().numpy().astype(np.int64), t[e0:e1]) and np.array_equal(sub.x.cpu().numpy().view(np.uint16), x[e0:e1])
    assert np.array_equal(sub.p.cpu().numpy(), p[e0:e1])


@pytest.mark.gpu
def test_ergo12_from_a_packed_batch_equals_the_soa_batch(cuda_device):
    """timestamps come back shifted by one constant per window: ERGO-12 (t - t.min() first thing) must not see it"""
    import torch
    import event_representation_study_b200.batched as eb
    from event_representation_study_b200 import packed
    H, W = 240, 304
    wins, b = _batch([60_000, 30_001, 7], H, W, 21, duration_us=20_000)
    want = eb.ergo12(eb.pack_events(wins, "cuda", t_dtype=np.int64), H, W)
    for fmt in (3, 4):
        pk = packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=fmt)
        assert pk is not None and pk.fmt == fmt
        got = eb.ergo12(packed.upload(pk, "cuda"), H, W)
        assert torch.equal(got, want)


def _same(a, b):
    import torch
    assert (a is None) == (b is None)
    if a is None:
        return
    assert a.fmt == b.fmt == 3 and a.x_bits == b.x_bits and a.y_bits == b.y_bits and a.block_shift == b.block_shift
    for k in ("rec3", "tbase", "esc_prefix", "esc_dt"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("sizes,H,W,dur,t_dtype,threads", [
    ([5000, 0, 64, 65, 1, 12345, 63, 129], 720, 1280, 2_000, np.int32, 1),
    ([4000, 130], 240, 304, 3_000_000, np.int64, 1),        # sparse: one event in four escapes (capacity grown on demand)
    ([300_000, 1, 0, 280_001], 720, 1280, 300_000, np.int32, 4),  # > 4096 blocks: the threaded path
    ([0, 0], 240, 304, 1000, np.int32, 2), ([], 240, 304, 1000, np.int32, 1),
])
def test_native_host_encoder_equals_the_numpy_packer(sizes, H, W, dur, t_dtype, threads):
    """evrep_pack_events_delta_host (one fused pass in C++, host threads) writes byte for byte what the numpy passes of
    packed._pack3 write: block records, base timestamps, escape prefix and escape table."""
    from event_representation_study_b200 import packed
    from event_representation_study_b200.synth import pack_batch
    wins, b = _batch(sizes, H, W, 11, duration_us=dur)
    if not sizes:
        b = pack_batch([])
    t = b["t"].astype(t_dtype) if len(b["t"]) else np.zeros(0, t_dtype)
    ref = packed.pack_host(b["x"], b["y"], t, b["p"], b["offsets"], H, W, fmt=3, native=False)
    nat = packed.pack_host(b["x"], b["y"], t, b["p"], b["offsets"], H, W, fmt=3, native=True, threads=threads)
    assert ref is not None
    _same(ref, nat)
    if int(b["offsets"][-1]):
        x, y, tt, p = packed.unpack_numpy(nat)
        assert np.array_equal(x, b["x"]) and np.array_equal(y, b["y"]) and np.array_equal(p, b["p"])


def test_native_host_encoder_refuses_what_the_numpy_packer_refuses():
    from event_representation_study_b200 import packed
    H, W = 240, 304
    wins, b = _batch([3000, 500], H, W, 21)
    args = lambda **kw: (kw.get("x", b["x"]), kw.get("y", b["y"]), kw.get("t", b["t"]), kw.get("p", b["p"]), b["offsets"], H, W)
    t_bad = b["t"].copy()
    t_bad[100], t_bad[101] = t_bad[101] + 5, t_bad[100]  # a negative difference inside a block
    p0 = b["p"].copy()
    p0[7] = 0
    for kw in ({"t": t_bad}, {"p": p0}):
        assert packed.pack_host(*args(**kw), fmt=3, native=True) is None
        assert packed.pack_host(*args(**kw), fmt=3, native=False) is None
    pk, pk_np = packed.pack_host(*args(p=p0), native=True), packed.pack_host(*args(p=p0), native=False)
    assert (pk is None) == (pk_np is None) and (pk is None or pk.fmt == pk_np.fmt != 3)  # falls through to the next formats, like the numpy path
    wins_d, bd = _batch([30000], H, W, 22, duration_us=30_000)
    pd = bd["p"].copy()
    pd[7] = 0
    pk = packed.pack_host(bd["x"], bd["y"], bd["t"], pd, bd["offsets"], H, W, native=True)
    assert pk is not None and pk.fmt == 4
    x_bad = b["x"].copy()
    x_bad[5] = W
    for native in (True, False):
        with pytest.raises(IndexError):
            packed.pack_host(*args(x=x_bad), native=native)
    # arrays in another layout (int64 coordinates) are packed by the numpy passes: same bytes
    _same(packed.pack_host(b["x"].astype(np.int64), b["y"].astype(np.int64), b["t"], b["p"].astype(np.int64), b["offsets"], H, W, fmt=3),
          packed.pack_host(*args(), fmt=3, native=True))
    # a sensor whose coordinates need more than 21 bits
    big = packed.pack_host(np.array([4000], np.uint16), np.array([3000], np.uint16), np.array([5], np.int32), np.array([1], np.int8), np.array([0, 1]), 4000, 5000)
    assert big is not None and big.fmt in (4, 6)


def test_simd_and_scalar_host_encoders_agree_on_adversarial_streams(monkeypatch):
    """The eight-events-per-step path (AVX2, int32 timestamps, full blocks) and the scalar loop must return the same code and
    the same bytes for anything: timestamps at the ends of the int32 range, negative differences, bad polarities, pixels
    outside the sensor, escapes of every size."""
    import ctypes
    from event_representation_study_b200 import _lib
    lib = _lib.lib
    rng = np.random.default_rng(2024)
    H, W = 480, 640

    def run(x, y, t, p, offs, scalar, zero_neg=False):
        if scalar:
            monkeypatch.setenv("EVREP_PACK_SCALAR", "1")
        else:
            monkeypatch.delenv("EVREP_PACK_SCALAR", raising=False)
        B = len(offs) - 1
        nb = int(lib.evrep_pack_delta_host_blocks(offs.ctypes.data, B))
        rec3, tbase, ep, ed = np.full(192 * nb + 8, 0xAB, np.uint8), np.zeros(nb, np.int32), np.zeros(nb + 1, np.uint32), np.zeros(len(x) + 1, np.uint32)
        need = ctypes.c_int64(-1)
        rc = lib.evrep_pack_events_delta_host(x.ctypes.data, y.ctypes.data, t.ctypes.data, 4, p.ctypes.data, offs.ctypes.data, B, H, W, rec3.ctypes.data,
                                              tbase.ctypes.data, ep.ctypes.data, ed.ctypes.data, len(ed), ctypes.byref(need), int(zero_neg), 1)
        return rc, (rec3, tbase, ep, ed[:max(need.value, 0)]) if rc == 0 else None

    kinds = ["clean", "gaps", "unsorted", "polarity", "pixel", "low-end", "high-end", "wrap"]
    seen_ok = seen_bad = 0
    for trial in range(64):
        kind = kinds[trial % len(kinds)]
        n = int(rng.integers(64, 400))
        offs = np.array([0, n], np.int64)
        x = rng.integers(0, W, n).astype(np.uint16)
        y = rng.integers(0, H, n).astype(np.uint16)
        p = rng.choice(np.array([-1, 1], np.int8), n)
        t = np.cumsum(rng.integers(0, 4, n)).astype(np.int64)
        if kind == "gaps":
            t = np.cumsum(rng.choice([0, 1, 2, 3, 1000, 2**20], n)).astype(np.int64)
        if kind == "low-end":
            t = t - 2**31
        if kind == "high-end":
            t = t + (2**31 - 1 - int(t[-1]))
        if kind == "wrap":  # spans more than 2^31: must be refused by both
            t = t - 2**31
            t[n // 2:] += 2**31 + 5
        t = np.clip(t, -2**31, 2**31 - 1).astype(np.int32)
        k = int(rng.integers(1, n))
        if kind == "unsorted":
            t[k] = t[k - 1] - int(rng.integers(1, 50)) if t[k - 1] > -2**31 + 60 else t[k]
        if kind == "polarity":
            p[k] = rng.choice(np.array([0, 2, -2, 127], np.int8))
        if kind == "pixel":
            (x if trial % 2 else y)[k] = (W if trial % 2 else H) + int(rng.integers(0, 3))
        zn = trial % 3 == 0  # every third trial also accepts p == 0 as "negative"
        if zn and kind == "clean":
            p[rng.integers(0, n, 5)] = 0
        a, b = run(x, y, t, p, offs, scalar=False, zero_neg=zn), run(x, y, t, p, offs, scalar=True, zero_neg=zn)
        assert a[0] == b[0], (kind, trial, a[0], b[0])
        if a[0] == 0:
            seen_ok += 1
            for u, v in zip(a[1], b[1]):
                assert np.array_equal(u, v), (kind, trial)
            assert (a[1][0][-8:] == 0xAB).all()  # nothing written past the last block
        else:
            seen_bad += 1
    assert seen_ok >= 16 and seen_bad >= 16


def test_zero_as_negative_lets_01_streams_use_the_three_byte_format():
    """pack_host(zero_as_negative=True): p == 0 travels in format 3 as "negative" and decodes as -1 - native encoder and numpy
    restatement write the same bytes, equal to packing the stream with its zeros replaced by -1."""
    from event_representation_study_b200 import packed
    H, W = 240, 304
    wins, b = _batch([5000, 64, 70], H, W, 31, polarity="01", duration_us=20_000)
    assert (b["p"] == 0).any() and (b["p"] == 1).any()
    args = (b["x"], b["y"], b["t"])
    plain = packed.pack_host(*args, b["p"], b["offsets"], H, W)
    assert plain is not None and plain.fmt in (4, 6)                 # lossless by default: the zeros need the 2-bit polarity code
    mapped = np.where(b["p"] == 0, -1, b["p"]).astype(np.int8)
    want = packed.pack_host(*args, mapped, b["offsets"], H, W, fmt=3, native=False)
    for native in (True, False):
        got = packed.pack_host(*args, b["p"], b["offsets"], H, W, native=native, zero_as_negative=True)
        _same(want, got)
    x, y, t, p = packed.unpack_numpy(got)
    assert np.array_equal(p, mapped) and np.array_equal(x, b["x"]) and np.array_equal(y, b["y"])


def test_host_packer_reuses_its_buffers():
    """packed.HostPacker: buffers allocated once, handed out round robin; every batch equals pack_host's bytes."""
    from event_representation_study_b200 import packed
    H, W = 240, 304
    hp = packed.HostPacker(max_events=20_000, max_windows=4, H=H, W=W, pin=False, slots=2, threads=1, escape_fraction=1.0)  # sparse test streams
    seen = []
    for k, sizes in enumerate(([5000, 0, 70], [64], [9000, 9000], [1, 2, 3, 4])):
        wins, b = _batch(sizes, H, W, 40 + k, duration_us=20_000)
        pk = hp.pack(b["x"], b["y"], b["t"], b["p"], b["offsets"])
        _same(packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=3, native=False), pk)
        seen.append(pk.rec3.data_ptr())
    assert seen[0] == seen[2] and seen[1] == seen[3] and seen[0] != seen[1]  # two slots, alternating
    wins, b = _batch([30_000], H, W, 50)
    with pytest.raises(ValueError):
        hp.pack(b["x"], b["y"], b["t"], b["p"], b["offsets"])               # beyond the capacity it was built for
    wins, b = _batch([300], H, W, 51, polarity="01")
    with pytest.raises(TypeError):
        hp.pack(b["x"].astype(np.int64), b["y"], b["t"], b["p"], b["offsets"])
    assert hp.pack(b["x"], b["y"], b["t"], b["p"], b["offsets"]) is None    # zeros do not fit one polarity bit ...
    hz = packed.HostPacker(1000, 1, H, W, pin=False, zero_as_negative=True, escape_fraction=1.0)
    assert hz.pack(b["x"], b["y"], b["t"], b["p"], b["offsets"]).fmt == 3    # ... unless the caller says they are the negatives
    sparse = packed.HostPacker(1000, 1, H, W, pin=False, escape_fraction=0.0)
    wins, b = _batch([900], H, W, 52, duration_us=3_000_000_000)
    if (np.diff(b["t"].astype(np.int64)) > 2).sum() > 1024:
        with pytest.raises(ValueError):
            sparse.pack(b["x"], b["y"], b["t"], b["p"], b["offsets"])


@pytest.mark.parametrize("fmt,sizes,dur,polarity,t_dtype,threads,shuffle", [
    (4, [5000, 0, 64, 65, 1, 12345], 2_000, "pm1", np.int32, 1, False),
    (4, [30000, 7], 30_000, "01", np.int64, 1, True),          # any order, zeros kept
    (6, [5000, 0, 256, 257, 1, 12345], 50_000, "01", np.int32, 1, True),
    (6, [300_000, 300_001, 0, 5], 300_000, "pm1", np.int32, 4, False),  # > 4096 blocks: the threaded path
    (4, [], 1000, "pm1", np.int32, 1, False),
])
def test_native_word_encoders_equal_the_numpy_packer(fmt, sizes, dur, polarity, t_dtype, threads, shuffle):
    """evrep_pack_events_host (formats 4 / 6) against the numpy passes of pack_host: same words, offsets and block bases"""
    import torch
    from event_representation_study_b200 import packed
    from event_representation_study_b200.synth import pack_batch
    H, W = 240, 304
    wins, b = _batch(sizes, H, W, 61, polarity=polarity, duration_us=dur)
    if not sizes:
        b = pack_batch([])
    t = b["t"].astype(t_dtype) if len(b["t"]) else np.zeros(0, t_dtype)
    if shuffle and len(t) > 10:  # events out of order inside a block: these formats do not care
        t[3], t[9] = t[9], t[3]
    ref = packed.pack_host(b["x"], b["y"], t, b["p"], b["offsets"], H, W, fmt=fmt, native=False)
    nat = packed.pack_host(b["x"], b["y"], t, b["p"], b["offsets"], H, W, fmt=fmt, native=True, threads=threads)
    assert (ref is None) == (nat is None)
    if ref is None:
        return
    assert ref.fmt == nat.fmt == fmt and ref.block_shift == nat.block_shift and ref.x_bits == nat.x_bits
    assert torch.equal(ref.word, nat.word) and torch.equal(ref.tbase, nat.tbase)
    assert (ref.dt16 is None) == (nat.dt16 is None) and (ref.dt16 is None or torch.equal(ref.dt16, nat.dt16))
    if int(b["offsets"][-1]):
        x, y, tt, p = packed.unpack_numpy(nat)
        assert np.array_equal(x, b["x"]) and np.array_equal(p, b["p"])


def test_native_word_encoders_refuse_like_the_numpy_packer():
    from event_representation_study_b200 import packed
    H, W = 240, 304
    wins, b = _batch([4000], H, W, 62, duration_us=300_000_000)  # 75 ms between events
    for native in (True, False):
        assert packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=4, native=native) is None
        assert packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=6, native=native) is None
    bad_p = b["p"].copy()
    bad_p[3] = 2
    for native in (True, False):
        with pytest.raises(ValueError):
            packed.pack_host(b["x"], b["y"], b["t"], bad_p, b["offsets"], H, W, native=native)
    bad_y = b["y"].copy()
    bad_y[0] = H
    for native in (True, False):
        with pytest.raises(IndexError):
            packed.pack_host(b["x"], bad_y, b["t"], b["p"], b["offsets"], H, W, fmt=4, native=native)
    # a sensor too large for format 4: the chain moves on to format 6 in both paths
    big_n = packed.pack_host(np.array([4000], np.uint16), np.array([3000], np.uint16), np.array([5], np.int32), np.array([0], np.int8), np.array([0, 1]), 40000, 50000)
    big_r = packed.pack_host(np.array([4000], np.uint16), np.array([3000], np.uint16), np.array([5], np.int32), np.array([0], np.int8), np.array([0, 1]), 40000, 50000, native=False)
    assert (big_n is None) == (big_r is None) and (big_n is None or big_n.fmt == big_r.fmt)


def test_pinned_branches_of_the_packers(monkeypatch):
    """pin=True takes its own allocation branch in every packer (buffers pinned first, filled through their pointers); with
    pin_memory replaced by a copy - there is no CUDA here - the results must equal the pageable ones."""
    import torch
    from event_representation_study_b200 import packed
    calls = []
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: calls.append(self.numel()) or self.clone())
    H, W = 240, 304
    wins, b = _batch([5000, 0, 300], H, W, 71, duration_us=20_000)
    for fmt in (3, 4, 6):
        for native in (True, False):
            a = packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=fmt, native=native, pin=True)
            c = packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=fmt, native=native, pin=False)
            assert a is not None and c is not None and a.fmt == c.fmt == fmt
            for k in ("word", "dt16", "tbase", "rec3", "esc_prefix", "esc_dt"):
                u, v = getattr(a, k), getattr(c, k)
                assert (u is None) == (v is None) and (u is None or torch.equal(u, v)), (fmt, native, k)
            parts = a.host_parts(0, 3)
            assert all(v.numel() > 0 for v in parts.values())
    hp = packed.HostPacker(10_000, 3, H, W, pin=True, escape_fraction=1.0)
    _same(hp.pack(b["x"], b["y"], b["t"], b["p"], b["offsets"]), packed.pack_host(b["x"], b["y"], b["t"], b["p"], b["offsets"], H, W, fmt=3, native=False))
    assert len(calls) > 10
