"""Packed host wire format for the end-to-end path (include/evrep.h, evrep_unpack_events).

The engine consumes events from HBM ~9x faster than the host link can deliver them as SoA arrays (9 B/event), so the
loader side can pack an event into 4 bytes (format 4: x, y, polarity and a small time offset to the base of its block of 64
events in one 32-bit word) or 6 bytes (format 6: the word plus a 16-bit offset, blocks of 256 events), and one decode kernel
restores the SoA arrays on the GPU.  Packing is a few vectorised numpy passes per window, the kind of work the reference's
DataLoader workers already do per sample (ev-YOLOv6/yolov6/data/gen1_2yolo.py:186-208 slices, concatenates and casts the
same arrays); decoded timestamps equal the originals up to one constant per window, which no representation can see."""
import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import batched as eb
from ._lib import check, lib

BLOCK_SHIFT = {4: 6, 6: 8}  # events per block: 64 / 256


def _bits(n):
    return max(1, int(n - 1).bit_length())


@dataclass
class PackedEvents:
    """Host side of a packed batch: `word` (uint32 per event, stored as int32), `dt16` (uint16 per event as int16, format 6
    only), `tbase` (int32 per block), CSR `offsets` (B + 1) and the format constants.  Tensors may be pinned."""
    word: torch.Tensor
    dt16: Optional[torch.Tensor]
    tbase: torch.Tensor
    offsets: np.ndarray
    fmt: int
    x_bits: int
    y_bits: int
    block_shift: int

    @property
    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in (self.word, self.dt16, self.tbase) if v is not None)

    def block_range(self, w0, w1):
        """[first, last) block of the windows [w0, w1) (blocks never straddle windows)"""
        n = np.diff(self.offsets)
        nb = (n + (1 << self.block_shift) - 1) >> self.block_shift
        pre = np.concatenate([[0], np.cumsum(nb)])
        return int(pre[w0]), int(pre[w1])


def pack_host(x, y, t, p, offsets, H, W, fmt=None, pin=False):
    """SoA numpy events of a CSR batch -> PackedEvents, or None when some block's time span fits neither format (sparse or
    unsorted streams: upload the SoA arrays instead).  fmt: 4, 6 or None (= the smallest that fits)."""
    offsets = np.ascontiguousarray(offsets, np.int64)
    total = int(offsets[-1])
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
    for f in ((fmt,) if fmt else (4, 6)):
        if f == 4 and xb + yb > 29:
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
    word = pk.word.numpy().view(np.uint32)
    xb, yb = pk.x_bits, pk.y_bits
    x = word & ((1 << xb) - 1)
    y = (word >> xb) & ((1 << yb) - 1)
    pc = (word >> (xb + yb)) & 3
    p = np.where(pc == 3, -1, pc).astype(np.int8)
    dt = (word >> (xb + yb + 2)).astype(np.int64) if pk.fmt == 4 else pk.dt16.numpy().view(np.uint16).astype(np.int64)
    n = np.diff(pk.offsets)
    nb = (n + (1 << pk.block_shift) - 1) >> pk.block_shift
    local = np.arange(int(pk.offsets[-1])) - np.repeat(pk.offsets[:-1], n)
    blk = np.repeat(np.concatenate([[0], np.cumsum(nb)])[:-1], n) + (local >> pk.block_shift)
    return x.astype(np.uint16), y.astype(np.uint16), pk.tbase.numpy().astype(np.int64)[blk] + dt, p


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
    """PackedEvents on the host -> EventBatch on `device`: three host-to-device copies and one decode kernel."""
    dev = torch.device(device)
    nb = pk.word.is_pinned()
    to = lambda v: v.to(dev, non_blocking=nb) if v is not None else None
    return decode(to(pk.word), to(pk.dt16), to(pk.tbase), pk.offsets, pk.fmt, pk.x_bits, pk.y_bits, pk.block_shift, out=out)
