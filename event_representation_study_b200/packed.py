"""Packed host wire format for the end-to-end path (include/evrep.h, evrep_unpack_events).

The engine consumes events from HBM ~9x faster than the host link can deliver them as SoA arrays (9 B/event), so the
loader side can pack an event into 3 bytes (format 3: x, y, a polarity bit and the 2-bit difference to the previous event's
timestamp, larger differences in a side table; blocks of 64 events; needs sorted windows, polarities -1 / +1 and x, y within
21 bits together), 4 bytes (format 4: x, y, polarity and a small time offset to the base of its block of 64 events in one
32-bit word) or 6 bytes (format 6: the word plus a 16-bit offset, blocks of 256 events), and one decode kernel restores the
SoA arrays on the GPU.  Packing is a few vectorised numpy passes per window, the kind of work the reference's
DataLoader workers already do per sample (ev-YOLOv6/yolov6/data/gen1_2yolo.py:186-208 slices, concatenates and casts the
same arrays); decoded timestamps equal the originals up to one constant per window, which no representation can see."""
import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import batched as eb
from ._lib import check, lib
from ._lib import EINVAL as _EINVAL, EUNSUPPORTED as _EUNSUPPORTED, EWORKSPACE as _EWORKSPACE

BLOCK_SHIFT = {3: 6, 4: 6, 6: 8}  # events per block: 64 / 64 / 256


def _bits(n):
    return max(1, int(n - 1).bit_length())


@dataclass
class PackedEvents:
    """Host side of a packed batch.  Formats 4 / 6: `word` (uint32 per event, stored as int32), `dt16` (uint16 per event as
    int16, format 6 only), `tbase` (int32 per block).  Format 3: `rec3` (uint8, 192 bytes per block of 64 events), `tbase`,
    `esc_prefix` (escapes before each block, int32 storage, blocks + 1 entries) and `esc_dt` (the escaped differences, int32
    storage).  CSR `offsets` (B + 1) and the format constants.  Tensors may be pinned."""
    word: Optional[torch.Tensor]
    dt16: Optional[torch.Tensor]
    tbase: torch.Tensor
    offsets: np.ndarray
    fmt: int
    x_bits: int
    y_bits: int
    block_shift: int
    rec3: Optional[torch.Tensor] = None
    esc_prefix: Optional[torch.Tensor] = None
    esc_dt: Optional[torch.Tensor] = None

    @property
    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in (self.word, self.dt16, self.tbase, self.rec3, self.esc_prefix, self.esc_dt) if v is not None)

    def block_range(self, w0, w1):
        """[first, last) block of the windows [w0, w1) (blocks never straddle windows)"""
        n = np.diff(self.offsets)
        nb = (n + (1 << self.block_shift) - 1) >> self.block_shift
        pre = np.concatenate([[0], np.cumsum(nb)])
        return int(pre[w0]), int(pre[w1])

    def host_parts(self, w0, w1):
        """The payload of the windows [w0, w1) as a dict of host tensor views (what a loader ships for that group)"""
        e0, e1 = int(self.offsets[w0]), int(self.offsets[w1])
        b0, b1 = self.block_range(w0, w1)
        if self.fmt == 3:
            q0, q1 = int(self.esc_prefix[b0]), int(self.esc_prefix[b1])
            return {"rec3": self.rec3[192 * b0:192 * b1], "tbase": self.tbase[b0:b1], "esc_prefix": self.esc_prefix[b0:b1 + 1],
                    "esc_dt": self.esc_dt[q0:max(q1, q0 + 1)]}  # never empty: a zero-size copy has no device pointer
        parts = {"word": self.word[e0:e1], "tbase": self.tbase[b0:b1]}
        if self.dt16 is not None:
            parts["dt16"] = self.dt16[e0:e1]
        return parts

    def decode_parts(self, dev_parts, local_offsets, out=None):
        """Device copies of host_parts(w0, w1) -> EventBatch of those windows (local_offsets = offsets[w0:w1 + 1] - offsets[w0])"""
        if self.fmt == 3:
            return decode_delta(dev_parts["rec3"], dev_parts["tbase"], dev_parts["esc_prefix"], dev_parts["esc_dt"], local_offsets, self.x_bits, self.y_bits,
                                out=out)
        return decode(dev_parts["word"], dev_parts.get("dt16"), dev_parts["tbase"], local_offsets, self.fmt, self.x_bits, self.y_bits, self.block_shift, out=out)


def _pack3(x, y, p8, rel, offsets, xb, yb, pin):
    """Format 3 or None: sorted windows, polarities -1 / +1, xb + yb <= 21"""
    total = int(offsets[-1])
    if xb + yb > 21 or (total and ((p8 == 0).any())):
        return None
    n = np.diff(offsets)
    nb = (n + 63) >> 6
    n_blocks = int(nb.sum())
    local = np.arange(total, dtype=np.int64) - np.repeat(offsets[:-1], n)
    blk = np.repeat(np.concatenate([[0], np.cumsum(nb)])[:-1], n) + (local >> 6)
    first = (local & 63) == 0
    d = np.zeros(total, np.int64)
    if total:
        d[1:] = rel[1:] - rel[:-1]
        d[first] = 0
        if d.min() < 0 or d.max() >= 2**32 or rel.min() < 0 or rel.max() >= 2**31:
            return None
    esc = d > 2
    code = np.where(esc, 3, d).astype(np.uint32)
    rec = x | (y << xb) | ((p8 > 0).astype(np.uint32) << (xb + yb)) | (code << (xb + yb + 1))
    rec3 = np.zeros((n_blocks * 64, 3), np.uint8)
    slot = blk * 64 + (local & 63)
    rec3[slot, 0], rec3[slot, 1], rec3[slot, 2] = rec & 255, (rec >> 8) & 255, (rec >> 16) & 255
    tbase = rel[first].astype(np.int32) if total else np.zeros(0, np.int32)
    esc_prefix = np.zeros(n_blocks + 1, np.int64)
    if total:
        esc_prefix[1:] = np.cumsum(np.bincount(blk[esc], minlength=n_blocks))
    esc_dt = d[esc].astype(np.uint32)
    if len(esc_dt) == 0:
        esc_dt = np.zeros(1, np.uint32)  # keeps the device pointer valid
    tens = [torch.from_numpy(rec3.reshape(-1)), torch.from_numpy(tbase), torch.from_numpy(esc_prefix.astype(np.uint32).view(np.int32)),
            torch.from_numpy(esc_dt.view(np.int32))]
    if pin:
        tens = [v.pin_memory() for v in tens]
    return PackedEvents(None, None, tens[1], offsets, 3, xb, yb, 6, rec3=tens[0], esc_prefix=tens[2], esc_dt=tens[3])


def _pack3_native(x, y, t, p, offsets, H, W, pin, threads, zero_as_negative=False):
    """Format 3 through the library's host encoder (evrep_pack_events_delta_host: one fused pass on a few threads, byte-identical
    to _pack3) -> PackedEvents, None when the stream does not fit the format, or NotImplemented when the arrays are not in the
    layout the encoder reads (it takes them as they are: no copies)."""
    total, B = int(offsets[-1]), len(offsets) - 1
    arrs = []
    for v, kinds in ((x, (np.uint16, np.int16)), (y, (np.uint16, np.int16)), (t, (np.int32, np.int64)), (p, (np.int8,))):
        v = np.asarray(v)
        if v.dtype.type not in kinds or not v.flags.c_contiguous or len(v) < total:
            return NotImplemented
        arrs.append(v)
    x, y, t, p = arrs
    n_blocks = int(lib.evrep_pack_delta_host_blocks(offsets.ctypes.data, B))
    if n_blocks < 0:
        return NotImplemented
    mk = (lambda n, dt: torch.empty(max(n, 1), dtype=dt).pin_memory()) if pin else (lambda n, dt: torch.empty(max(n, 1), dtype=dt))
    rec3, tbase, esc_prefix = mk(192 * n_blocks, torch.uint8), mk(n_blocks, torch.int32), mk(n_blocks + 1, torch.int32)
    cap = max(1024, total // 256)  # a sorted stream of N events in ~0.3 s escapes ~1e-5 of them; grown on demand
    need = ctypes.c_int64(0)
    while True:
        esc_dt = mk(cap, torch.int32)
        rc = lib.evrep_pack_events_delta_host(x.ctypes.data, y.ctypes.data, t.ctypes.data, t.dtype.itemsize, p.ctypes.data, offsets.ctypes.data, B, H, W,
                                              rec3.data_ptr(), tbase.data_ptr(), esc_prefix.data_ptr(), esc_dt.data_ptr(), cap, ctypes.byref(need),
                                              int(bool(zero_as_negative)), int(threads))
        if rc == _EWORKSPACE:
            cap = int(need.value)
            continue
        break
    if rc == _EUNSUPPORTED:
        return None
    if rc == _EINVAL and b"outside the sensor" in lib.evrep_last_error():
        raise IndexError("event outside the sensor")
    check(rc)
    n_esc = int(need.value)
    if n_esc == 0:
        esc_dt[:1] = 0  # (the numpy packer ships one zero so that the device pointer stays valid)
    return PackedEvents(None, None, tbase[:n_blocks], offsets, 3, _bits(W), _bits(H), 6, rec3=rec3[:192 * n_blocks],
                        esc_prefix=esc_prefix[:n_blocks + 1], esc_dt=esc_dt[:max(n_esc, 1)])


def _native_arrays(x, y, t, p, total):
    """The arrays as the library's host encoders read them (no copies), or None when they are in another layout"""
    out = []
    for v, kinds in ((x, (np.uint16, np.int16)), (y, (np.uint16, np.int16)), (t, (np.int32, np.int64)), (p, (np.int8,))):
        v = np.asarray(v)
        if v.dtype.type not in kinds or not v.flags.c_contiguous or len(v) < total:
            return None
        out.append(v)
    return out


def _pack_words_native(x, y, t, p, offsets, H, W, f, pin, threads):
    """Formats 4 / 6 through evrep_pack_events_host (byte-identical to the numpy passes of pack_host) -> PackedEvents, None when
    the stream does not fit the format, NotImplemented when the arrays are in another layout."""
    total, B = int(offsets[-1]), len(offsets) - 1
    arrs = _native_arrays(x, y, t, p, total)
    if arrs is None:
        return NotImplemented
    x, y, t, p = arrs
    n_blocks = int(lib.evrep_pack_host_blocks(offsets.ctypes.data, B, f))
    if n_blocks < 0:
        return NotImplemented
    mk = (lambda n, dt: torch.empty(max(n, 1), dtype=dt).pin_memory()) if pin else (lambda n, dt: torch.empty(max(n, 1), dtype=dt))
    word, tbase = mk(total, torch.int32), mk(n_blocks, torch.int32)
    dt16 = mk(total, torch.int16) if f == 6 else None
    rc = lib.evrep_pack_events_host(x.ctypes.data, y.ctypes.data, t.ctypes.data, t.dtype.itemsize, p.ctypes.data, offsets.ctypes.data, B, H, W, f,
                                    word.data_ptr(), dt16.data_ptr() if dt16 is not None else None, tbase.data_ptr(), int(threads))
    if rc == _EUNSUPPORTED:
        return None
    if rc == _EINVAL:
        msg = lib.evrep_last_error()
        if b"outside the sensor" in msg:
            raise IndexError("event outside the sensor")
        if b"polarities" in msg:
            raise ValueError("polarities must be in {-1, 0, 1}")
    check(rc)
    return PackedEvents(word[:total], dt16[:total] if dt16 is not None else None, tbase[:n_blocks], offsets, f, _bits(W), _bits(H), BLOCK_SHIFT[f])


class HostPacker:
    """Format-3 encoder with its output buffers allocated ONCE (pinned, if asked): what a loader thread calls per batch.
    `pack_host(pin=True)` pins fresh buffers on every call, and cudaHostAlloc costs milliseconds; this object keeps `slots`
    sets of buffers sized for `max_events` / `max_windows` and hands them out round robin, so batch k + 1 can be packed while
    the copy of batch k is still in flight.  pack() returns a PackedEvents whose tensors are views into the current slot
    (valid until that slot comes round again), or None when the stream does not fit format 3 (see evrep_pack_events_delta_host).
    Arrays must be in the encoder's layout: x, y uint16 / int16, t int32 / int64, p int8, C-contiguous."""

    def __init__(self, max_events, max_windows, H, W, pin=True, slots=2, threads=0, zero_as_negative=False, escape_fraction=1 / 64):
        self.H, self.W, self.threads, self.zero_as_negative = int(H), int(W), int(threads), bool(zero_as_negative)
        self.max_events, self.max_windows = int(max_events), int(max_windows)
        self.max_blocks = (self.max_events + 63) // 64 + self.max_windows  # every window may end in a partial block
        self.esc_capacity = max(1024, int(self.max_events * escape_fraction))
        mk = (lambda n, dt: torch.empty(n, dtype=dt).pin_memory()) if pin else (lambda n, dt: torch.empty(n, dtype=dt))
        self._slots = [{"rec3": mk(192 * self.max_blocks, torch.uint8), "tbase": mk(self.max_blocks, torch.int32),
                        "esc_prefix": mk(self.max_blocks + 1, torch.int32), "esc_dt": mk(self.esc_capacity, torch.int32)} for _ in range(max(1, int(slots)))]
        self._next = 0

    def pack(self, x, y, t, p, offsets):
        offsets = np.ascontiguousarray(offsets, np.int64)
        total, B = int(offsets[-1]), len(offsets) - 1
        if total > self.max_events or B > self.max_windows:
            raise ValueError(f"batch of {total} events in {B} windows exceeds the packer's capacity ({self.max_events}, {self.max_windows})")
        for v, kinds in ((x, (np.uint16, np.int16)), (y, (np.uint16, np.int16)), (t, (np.int32, np.int64)), (p, (np.int8,))):
            if not isinstance(v, np.ndarray) or v.dtype.type not in kinds or not v.flags.c_contiguous or len(v) < total:
                raise TypeError("HostPacker takes C-contiguous numpy arrays: x, y uint16 / int16, t int32 / int64, p int8")
        n_blocks = int(lib.evrep_pack_delta_host_blocks(offsets.ctypes.data, B))
        buf = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        need = ctypes.c_int64(0)
        rc = lib.evrep_pack_events_delta_host(x.ctypes.data, y.ctypes.data, t.ctypes.data, t.dtype.itemsize, p.ctypes.data, offsets.ctypes.data, B, self.H, self.W,
                                              buf["rec3"].data_ptr(), buf["tbase"].data_ptr(), buf["esc_prefix"].data_ptr(), buf["esc_dt"].data_ptr(),
                                              self.esc_capacity, ctypes.byref(need), int(self.zero_as_negative), self.threads)
        if rc == _EUNSUPPORTED:
            return None
        if rc == _EWORKSPACE:
            raise ValueError(f"{int(need.value)} timestamp escapes exceed the packer's table of {self.esc_capacity} (raise escape_fraction: the stream is sparse)")
        if rc == _EINVAL and b"outside the sensor" in lib.evrep_last_error():
            raise IndexError("event outside the sensor")
        check(rc)
        n_esc = int(need.value)
        if n_esc == 0:
            buf["esc_dt"][:1] = 0
        return PackedEvents(None, None, buf["tbase"][:n_blocks], offsets, 3, _bits(self.W), _bits(self.H), 6, rec3=buf["rec3"][:192 * n_blocks],
                            esc_prefix=buf["esc_prefix"][:n_blocks + 1], esc_dt=buf["esc_dt"][:max(n_esc, 1)])


def pack_host(x, y, t, p, offsets, H, W, fmt=None, pin=False, native=True, threads=0, zero_as_negative=False):
    """SoA numpy events of a CSR batch -> PackedEvents, or None when the stream fits none of the formats (sparse or unsorted
    streams: upload the SoA arrays instead).  fmt: 3, 4, 6 or None (= the smallest that fits).  The library's host encoders
    write every format when the arrays are uint16 / int16 x, y, int32 / int64 t and int8 p (native=False: the numpy
    passes, ~30 x slower, kept as the restatement the tests hold the encoders to); threads: host threads of that encoder
    (0 = as many as the machine has, at most 16).  zero_as_negative: format 3 has one polarity bit; with this flag a {0, 1} stream
    fits it too - p == 0 travels as "negative" and comes back as -1, which every representation here treats like 0 when the window
    holds no -1 (operations.py:59-61) - instead of falling back to the 4-byte format (formats 4 / 6 keep the 0: they are lossless)."""
    offsets = np.ascontiguousarray(offsets, np.int64)
    total = int(offsets[-1])
    arrs = _native_arrays(x, y, t, p, total) if native else None
    if arrs is not None:  # the library's host encoders, smallest format first; they read the arrays as they are
        try:
            if total and (arrs[3][:total].min() < -1 or arrs[3][:total].max() > 1):
                raise ValueError("polarities must be in {-1, 0, 1}")
            for f in ((fmt,) if fmt else (3, 4, 6)):
                if f == 3:
                    pk = _pack3_native(*arrs, offsets, H, W, pin, threads, zero_as_negative)
                else:
                    pk = _pack_words_native(*arrs, offsets, H, W, f, pin, threads)
                if pk is NotImplemented:
                    break
                if pk is not None:
                    return pk
            else:
                return None
        except (IndexError, ValueError):
            raise
        except Exception as e:  # e.g. pinned memory unavailable: the numpy passes below write the same bytes
            import warnings
            warnings.warn(f"native event packer unavailable ({type(e).__name__}: {e}); packing with numpy")
    x = np.asarray(x)[:total].astype(np.uint32)
    y = np.asarray(y)[:total].astype(np.uint32)
    if total and (x.max() >= W or y.max() >= H):
        raise IndexError("event outside the sensor")
    p8 = np.asarray(p)[:total].astype(np.int8)
    if total and (p8.min() < -1 or p8.max() > 1):
        raise ValueError("polarities must be in {-1, 0, 1}")
    xb, yb = _bits(W), _bits(H)
    n = np.diff(offsets)
    t64 = np.asarray(t)[:total].astype(np.int64)
    first = np.repeat(t64[offsets[:-1][n > 0]], n[n > 0]) if total else np.zeros(0, np.int64)
    rel = t64 - first
    for f in ((fmt,) if fmt else (3, 4, 6)):
        if f == 3:
            pk3 = _pack3(x, y, np.where(p8 == 0, np.int8(-1), p8) if zero_as_negative else p8, rel, offsets, xb, yb, pin)
            if pk3 is not None:
                return pk3
            continue
        if (f == 4 and xb + yb > 29) or xb + yb > 30:  # no room for the time offset / for the polarity code in a 32-bit word
            continue
        bs = BLOCK_SHIFT[f]
        nb = (n + (1 << bs) - 1) >> bs
        # global index of every block's first event
        if total:
            blk_w = np.repeat(np.arange(len(n)), nb)
            blk_pre = np.concatenate([[0], np.cumsum(nb)])[:-1]
            starts = offsets[:-1][blk_w] + ((np.arange(int(nb.sum())) - blk_pre[blk_w]) << bs)
            base = np.minimum.reduceat(rel, starts)
            lens = np.diff(np.concatenate([starts, [total]]))
            dt = rel - np.repeat(base, lens)
        else:
            base, dt = np.zeros(0, np.int64), np.zeros(0, np.int64)
        limit = (1 << (30 - xb - yb)) if f == 4 else 65536
        if total and (dt.max() >= limit or base.min() < -2**31 or base.max() >= 2**31):
            continue
        word = x | (y << xb) | ((p8.astype(np.uint32) & 3) << (xb + yb))
        dt16 = None
        if f == 4:
            word = word | (dt.astype(np.uint32) << (xb + yb + 2))
        else:
            dt16 = torch.from_numpy(dt.astype(np.uint16).view(np.int16))
        tens = [torch.from_numpy(word.astype(np.uint32).view(np.int32)), dt16, torch.from_numpy(base.astype(np.int32))]
        if pin:
            tens = [v.pin_memory() if v is not None else None for v in tens]
        return PackedEvents(tens[0], tens[1], tens[2], offsets, f, xb, yb, bs)
    return None


def unpack_numpy(pk):
    """Host reference of the decode (tests): -> x, y, t (relative to each window's first timestamp), p."""
    xb, yb = pk.x_bits, pk.y_bits
    n = np.diff(pk.offsets)
    nb = (n + (1 << pk.block_shift) - 1) >> pk.block_shift
    local = np.arange(int(pk.offsets[-1])) - np.repeat(pk.offsets[:-1], n)
    blk = np.repeat(np.concatenate([[0], np.cumsum(nb)])[:-1], n) + (local >> pk.block_shift)
    if pk.fmt == 3:
        b3 = pk.rec3.numpy().reshape(-1, 3).astype(np.uint32)
        rec = (b3[:, 0] | (b3[:, 1] << 8) | (b3[:, 2] << 16))[blk * 64 + (local & 63)]
        code = (rec >> (xb + yb + 1)) & 3
        d = code.astype(np.int64)
        d[code == 3] = pk.esc_dt.numpy().view(np.uint32)[: int((code == 3).sum())]
        run = np.cumsum(d)
        first = np.flatnonzero((local & 63) == 0)
        run = run - np.repeat(run[first], np.diff(np.concatenate([first, [len(d)]])))  # restart at every block (its first code is 0)
        t = pk.tbase.numpy().astype(np.int64)[blk] + run
        p = np.where((rec >> (xb + yb)) & 1, 1, -1).astype(np.int8)
        return (rec & ((1 << xb) - 1)).astype(np.uint16), ((rec >> xb) & ((1 << yb) - 1)).astype(np.uint16), t, p
    word = pk.word.numpy().view(np.uint32)
    x = word & ((1 << xb) - 1)
    y = (word >> xb) & ((1 << yb) - 1)
    pc = (word >> (xb + yb)) & 3
    p = np.where(pc == 3, -1, pc).astype(np.int8)
    dt = (word >> (xb + yb + 2)).astype(np.int64) if pk.fmt == 4 else pk.dt16.numpy().view(np.uint16).astype(np.int64)
    return x.astype(np.uint16), y.astype(np.uint16), pk.tbase.numpy().astype(np.int64)[blk] + dt, p


def decode_delta(rec3, tbase, esc_prefix, esc_dt, offsets, x_bits, y_bits, out=None):
    """Format-3 payload already on the GPU -> EventBatch (evrep_unpack_events_delta)"""
    offsets = np.ascontiguousarray(offsets, np.int64)
    total, B = int(offsets[-1]), len(offsets) - 1
    dev = rec3.device
    eb._require_current(dev)
    if out is None:
        out = {"x": torch.empty(total, dtype=torch.int16, device=dev), "y": torch.empty(total, dtype=torch.int16, device=dev),
               "t": torch.empty(total, dtype=torch.int32, device=dev), "p": torch.empty(total, dtype=torch.int8, device=dev)}
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = eb._workspace(dev, ("unpack", stream), max(int(lib.evrep_unpack_delta_workspace_bytes(B)), 256))
    check(lib.evrep_unpack_events_delta(rec3.data_ptr(), tbase.data_ptr(), esc_prefix.data_ptr(), esc_dt.data_ptr(), offsets.ctypes.data, B, x_bits, y_bits,
                                        out["x"].data_ptr(), out["y"].data_ptr(), out["t"].data_ptr(), out["p"].data_ptr(), ws.data_ptr(), ws.numel(), stream))
    return eb.EventBatch(out["x"][:total], out["y"][:total], out["t"][:total], out["p"][:total], offsets)


def decode(word, dt16, tbase, offsets, fmt, x_bits, y_bits, block_shift, out=None):
    """Packed payload already on the GPU (word / dt16 / tbase CUDA tensors of the windows in `offsets`) -> EventBatch.
    `out`: optional dict of preallocated x, y (int16), t (int32), p (int8) CUDA tensors."""
    offsets = np.ascontiguousarray(offsets, np.int64)
    total, B = int(offsets[-1]), len(offsets) - 1
    dev = word.device
    eb._require_current(dev)
    if out is None:
        out = {"x": torch.empty(total, dtype=torch.int16, device=dev), "y": torch.empty(total, dtype=torch.int16, device=dev),
               "t": torch.empty(total, dtype=torch.int32, device=dev), "p": torch.empty(total, dtype=torch.int8, device=dev)}
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws = eb._workspace(dev, ("unpack", stream), max(int(lib.evrep_unpack_workspace_bytes(B, total)), 256))
    check(lib.evrep_unpack_events(word.data_ptr(), dt16.data_ptr() if dt16 is not None else None, tbase.data_ptr(), offsets.ctypes.data, B, fmt,
                                  x_bits, y_bits, block_shift, out["x"].data_ptr(), out["y"].data_ptr(), out["t"].data_ptr(), out["p"].data_ptr(),
                                  ws.data_ptr(), ws.numel(), stream))
    return eb.EventBatch(out["x"][:total], out["y"][:total], out["t"][:total], out["p"][:total], offsets)


def upload(pk, device="cuda", out=None):
    """PackedEvents on the host -> EventBatch on `device`: a few host-to-device copies and one decode kernel."""
    dev = torch.device(device)
    parts = pk.host_parts(0, len(pk.offsets) - 1)
    dev_parts = {k: v.to(dev, non_blocking=v.is_pinned()) for k, v in parts.items()}
    return pk.decode_parts(dev_parts, pk.offsets, out=out)
