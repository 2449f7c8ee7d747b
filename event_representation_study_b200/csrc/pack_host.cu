// Host-side encoder of packed wire format 3 (the loader's half of evrep_unpack_events_delta): SoA events in host memory ->
// 3 bytes per event + per-block base timestamps + the escape table, byte for byte what packed.py's numpy packer writes.
//
// The numpy packer needs ~10 vectorised passes and manages ~12 M events/s on a core; the engine takes 17 G events/s over the
// link in this format, so a loader that packs on the fly needs a packer that runs at memory speed: one fused pass per block
// of 64 events, blocks spread over a few host threads.  (The reference's loaders do the equivalent slicing / casting per
// sample in Python, ev-YOLOv6/yolov6/data/gen1_2yolo.py:186-208; this is the native replacement for that step.)
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "evrep_common.cuh"

namespace {

int bits_for(int n) {  // packed.py::_bits
  int v = n - 1, b = 0;
  while (v > 0) { ++b; v >>= 1; }
  return b < 1 ? 1 : b;
}

struct Job {
  const uint16_t* x;
  const uint16_t* y;
  const void* t;
  int t_bytes;
  const int8_t* p;
  const int64_t* offs;
  int B, H, W, xb, yb;
  const int64_t* blk_prefix;  // B + 1: first block of every window
  uint8_t* rec3;
  int32_t* tbase;
  uint32_t* esc_count;  // per block (pass 1), then exclusive prefix in place -> esc_prefix
  uint32_t* esc_dt;
  std::atomic<int> status{0};  // bit 0 unsorted, bit 1 polarity not -1 / +1, bit 2 pixel outside the sensor, bit 3 time range
};

// blocks [b0, b1): pass 1 writes rec3 / tbase / escape counts, pass 2 (fill) writes the escapes at their prefix.
// Everything the loop reads is copied into locals first: the byte stores into rec3 may alias anything as far as the compiler
// knows, and would otherwise force a reload of every field of the job after each of them.
template <typename TT>
void run_blocks_t(Job& j, int64_t b0, int64_t b1, bool fill) {
  const uint16_t* __restrict__ x = j.x;
  const uint16_t* __restrict__ y = j.y;
  const TT* __restrict__ t = (const TT*)j.t;
  const int8_t* __restrict__ p = j.p;
  const int64_t* __restrict__ offs = j.offs;
  const int64_t* __restrict__ blk_prefix = j.blk_prefix;
  uint8_t* __restrict__ rec3 = j.rec3;
  const uint32_t Wd = (uint32_t)j.W, Hd = (uint32_t)j.H;
  const int xb = j.xb;
  const uint32_t sh_p = (uint32_t)(j.xb + j.yb), sh_c = sh_p + 1u;
  int w = (int)(std::upper_bound(blk_prefix, blk_prefix + j.B + 1, b0) - blk_prefix) - 1;  // window of block b0
  int bad = 0;
  for (int64_t b = b0; b < b1; ++b) {
    while (b >= blk_prefix[w + 1]) ++w;
    const int64_t ws = offs[w], we = offs[w + 1];
    const int64_t e0 = ws + ((b - blk_prefix[w]) << 6), e1 = std::min(e0 + 64, we);
    const int64_t t_first = (int64_t)t[ws];
    if (fill) {
      uint32_t* dst = j.esc_dt + j.esc_count[b];
      for (int64_t i = e0 + 1; i < e1; ++i) {
        const int64_t d = (int64_t)t[i] - (int64_t)t[i - 1];
        if (d > 2) *dst++ = (uint32_t)d;
      }
      continue;
    }
    uint8_t* out = rec3 + 192 * b;
    const int64_t rel0 = (int64_t)t[e0] - t_first;
    if (rel0 < 0) bad |= 1;
    if (rel0 >= ((int64_t)1 << 31)) bad |= 8;
    j.tbase[b] = (int32_t)rel0;
    uint32_t n_esc = 0;
    // one event -> its 24-bit record (the first event of a block is compared with itself: difference 0)
    auto record = [&](int64_t i) -> uint32_t {
      const int64_t ti = (int64_t)t[i];
      const int64_t d = ti - (int64_t)t[i > e0 ? i - 1 : i];
      const uint32_t xi = x[i], yi = y[i];
      const int pi = p[i];
      bad |= (d < 0) | ((pi != 1 && pi != -1) << 1) | ((xi >= Wd || yi >= Hd) << 2) | ((d >= ((int64_t)1 << 32) || ti - t_first >= ((int64_t)1 << 31)) << 3);
      n_esc += d > 2;
      return xi | (yi << xb) | ((pi > 0 ? 1u : 0u) << sh_p) | ((d > 2 ? 3u : (uint32_t)d) << sh_c);
    };
    int64_t i = e0;
    for (; i + 4 <= e1; i += 4) {  // four records = three 32-bit stores (little endian, like the byte stores below)
      const uint32_t r0 = record(i), r1 = record(i + 1), r2 = record(i + 2), r3 = record(i + 3);
      const uint32_t w0 = r0 | (r1 << 24), w1 = (r1 >> 8) | (r2 << 16), w2 = (r2 >> 16) | (r3 << 8);
      memcpy(out, &w0, 4);
      memcpy(out + 4, &w1, 4);
      memcpy(out + 8, &w2, 4);
      out += 12;
    }
    for (; i < e1; ++i) {
      const uint32_t rec = record(i);
      out[0] = (uint8_t)rec;
      out[1] = (uint8_t)(rec >> 8);
      out[2] = (uint8_t)(rec >> 16);
      out += 3;
    }
    if (e1 - e0 < 64) memset(out, 0, (size_t)(64 - (e1 - e0)) * 3);  // the tail of a window's last block
    j.esc_count[b] = n_esc;
  }
  if (bad) j.status.fetch_or(bad);
}
void run_blocks(Job& j, int64_t b0, int64_t b1, bool fill) {
  if (j.t_bytes == 4) run_blocks_t<int32_t>(j, b0, b1, fill);
  else run_blocks_t<int64_t>(j, b0, b1, fill);
}

void parallel_blocks(Job& j, int64_t n_blocks, int n_threads, bool fill) {
  if (n_threads <= 1 || n_blocks < 4096) {
    run_blocks(j, 0, n_blocks, fill);
    return;
  }
  std::vector<std::thread> th;
  const int64_t per = (n_blocks + n_threads - 1) / n_threads;
  for (int k = 0; k < n_threads; ++k) {
    const int64_t b0 = k * per, b1 = std::min(n_blocks, b0 + per);
    if (b0 >= b1) break;
    th.emplace_back([&j, b0, b1, fill] { run_blocks(j, b0, b1, fill); });
  }
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" int64_t evrep_pack_delta_host_blocks(const int64_t* win_offsets, int B) {
  if (!win_offsets || B < 0) return -1;
  int64_t nb = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t n = win_offsets[b + 1] - win_offsets[b];
    if (n < 0) return -1;
    nb += (n + 63) >> 6;
  }
  return nb;
}

extern "C" int evrep_pack_events_delta_host(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                                            int B, int H, int W, uint8_t* rec3, int32_t* tbase, uint32_t* esc_prefix, uint32_t* esc_dt,
                                            int64_t esc_capacity, int64_t* n_escapes, int n_threads) {
  using evrep::set_error;
  if (B < 0 || !win_offsets || (t_bytes != 4 && t_bytes != 8) || H < 1 || W < 1 || !esc_prefix || !n_escapes) { set_error("bad argument"); return EVREP_EINVAL; }
  const int xb = bits_for(W), yb = bits_for(H);
  if (xb + yb > 21) { set_error("wire format 3 holds x and y in 21 bits; %d x %d needs %d", W, H, xb + yb); return EVREP_EUNSUPPORTED; }
  std::vector<int64_t> blk_prefix((size_t)B + 1, 0);
  for (int b = 0; b < B; ++b) {
    const int64_t n = win_offsets[b + 1] - win_offsets[b];
    if (n < 0 || win_offsets[b] < 0) { set_error("win_offsets must be non-decreasing and non-negative"); return EVREP_EINVAL; }
    blk_prefix[(size_t)b + 1] = blk_prefix[(size_t)b] + ((n + 63) >> 6);
  }
  const int64_t n_blocks = blk_prefix[(size_t)B];
  const int64_t total = B ? win_offsets[B] : 0;
  if (total > 0 && (!x || !y || !t || !p || !rec3 || !tbase)) { set_error("null event / output array"); return EVREP_EINVAL; }
  if (n_blocks >= ((int64_t)1 << 31)) { set_error("too many blocks"); return EVREP_EUNSUPPORTED; }
  Job j;
  j.x = x; j.y = y; j.t = t; j.t_bytes = t_bytes; j.p = p; j.offs = win_offsets;
  j.B = B; j.H = H; j.W = W; j.xb = xb; j.yb = yb;
  j.blk_prefix = blk_prefix.data();
  j.rec3 = rec3; j.tbase = tbase; j.esc_count = esc_prefix; j.esc_dt = esc_dt;
  if (n_threads < 1) n_threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  parallel_blocks(j, n_blocks, n_threads, false);
  const int bad = j.status.load();
  if (bad & 4) { set_error("event outside the sensor"); return EVREP_EINVAL; }  // (the numpy packer raises for this one before anything else)
  if (bad & 1) { set_error("wire format 3 needs time-sorted windows"); return EVREP_EUNSUPPORTED; }
  if (bad & 2) { set_error("wire format 3 needs polarities -1 / +1"); return EVREP_EUNSUPPORTED; }
  if (bad & 8) { set_error("timestamps of a window span 2^31 us or more"); return EVREP_EUNSUPPORTED; }
  // exclusive prefix of the per-block escape counts, in place (blocks + 1 entries)
  uint64_t run = 0;
  for (int64_t b = 0; b < n_blocks; ++b) {
    const uint32_t c = esc_prefix[b];
    esc_prefix[b] = (uint32_t)run;
    run += c;
  }
  esc_prefix[n_blocks] = (uint32_t)run;
  *n_escapes = (int64_t)run;
  if (run >= ((uint64_t)1 << 32)) { set_error("too many escapes"); return EVREP_EUNSUPPORTED; }
  if ((int64_t)run > esc_capacity || (run > 0 && !esc_dt)) {
    set_error("escape table needs %lld entries, %lld given", (long long)run, (long long)esc_capacity);
    return EVREP_EWORKSPACE;
  }
  if (run > 0) parallel_blocks(j, n_blocks, n_threads, true);
  return EVREP_OK;
}
