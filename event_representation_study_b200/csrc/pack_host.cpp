// Host-side encoder of packed wire format 3 (the loader's half of evrep_unpack_events_delta): SoA events in host memory ->
// 3 bytes per event + per-block base timestamps + the escape table, byte for byte what packed.py's numpy packer writes.
//
// The numpy packer needs ~10 vectorised passes and manages ~12 M events/s on a core; the engine takes 17 G events/s over the
// link in this format, so a loader that packs on the fly needs a packer that runs at memory speed: one fused pass per block
// of 64 events (eight events per step with AVX2), blocks spread over a few host threads.  (The reference's loaders do the equivalent slicing / casting per
// sample in Python, ev-YOLOv6/yolov6/data/gen1_2yolo.py:186-208; this is the native replacement for that step.)
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/evrep.h"
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define EVREP_PACK_AVX2 1
#endif

namespace evrep {
void set_error(const char* fmt, ...);  // api.cu: thread-local error text
}

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
  bool zero_neg = false;       // p == 0 is accepted and written like p == -1 (the caller's consumers treat both as "negative")
  bool avx2 = false;           // full blocks of int32 timestamps take the eight-events-per-step path
  std::atomic<int> status{0};  // bit 0 unsorted, bit 1 polarity not -1 / +1, bit 2 pixel outside the sensor, bit 3 time range
};

#ifdef EVREP_PACK_AVX2
// One full block of 64 events with int32 timestamps, eight events per step.  All checks are exact without 64-bit arithmetic:
// t >= t_first is a signed compare, t - t_first < 2^31 is "the wrapped difference is not negative", and once both neighbours
// pass those two, their difference cannot wrap.  Returns the `bad` bits; *n_esc_out = escapes of the block.
__attribute__((target("avx2"))) int block64_avx2(const uint16_t* x, const uint16_t* y, const int32_t* t, const int8_t* p, int32_t t_first, uint32_t Wd,
                                                  uint32_t Hd, int xb, uint32_t sh_p, uint32_t sh_c, bool zero_neg, uint8_t* out, uint32_t* n_esc_out) {
  const __m256i v_first = _mm256_set1_epi32(t_first), v_three = _mm256_set1_epi32(3), v_two = _mm256_set1_epi32(2), v_one = _mm256_set1_epi32(1),
                v_m1 = _mm256_set1_epi32(-1), v_zero = _mm256_setzero_si256(), v_w = _mm256_set1_epi32((int)Wd - 1), v_h = _mm256_set1_epi32((int)Hd - 1);
  const __m128i c_xb = _mm_cvtsi32_si128(xb), c_p = _mm_cvtsi32_si128((int)sh_p), c_c = _mm_cvtsi32_si128((int)sh_c);
  const __m256i pick = _mm256_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1, 0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1);
  const __m256i shift_in = _mm256_setr_epi32(0, 0, 1, 2, 3, 4, 5, 6);
  const __m256i v_neg_alt = _mm256_set1_epi32(zero_neg ? 0 : -1);  // the second value accepted as "negative"
  __m256i f_sort = v_zero, f_pol = v_zero, f_pix = v_zero, f_rng = v_zero;
  uint32_t n_esc = 0;
  for (int g = 0; g < 8; ++g) {
    const __m256i tv = _mm256_loadu_si256((const __m256i*)(t + 8 * g));
    const __m256i tprev = g ? _mm256_loadu_si256((const __m256i*)(t + 8 * g - 1)) : _mm256_permutevar8x32_epi32(tv, shift_in);  // first event: itself
    const __m256i xv = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i*)(x + 8 * g)));
    const __m256i yv = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i*)(y + 8 * g)));
    const __m256i pv = _mm256_cvtepi8_epi32(_mm_loadl_epi64((const __m128i*)(p + 8 * g)));
    const __m256i rel = _mm256_sub_epi32(tv, v_first);
    f_rng = _mm256_or_si256(f_rng, _mm256_cmpgt_epi32(v_zero, rel));   // t - t_first >= 2^31 (or t < t_first, caught below as well)
    f_sort = _mm256_or_si256(f_sort, _mm256_cmpgt_epi32(v_first, tv));  // before the window's first event
    const __m256i d = _mm256_sub_epi32(tv, tprev);
    f_sort = _mm256_or_si256(f_sort, _mm256_cmpgt_epi32(v_zero, d));
    const __m256i pos = _mm256_cmpeq_epi32(pv, v_one);
    f_pol = _mm256_or_si256(f_pol, _mm256_andnot_si256(_mm256_or_si256(_mm256_or_si256(pos, _mm256_cmpeq_epi32(pv, v_m1)), _mm256_cmpeq_epi32(pv, v_neg_alt)), v_m1));
    f_pix = _mm256_or_si256(f_pix, _mm256_or_si256(_mm256_cmpgt_epi32(xv, v_w), _mm256_cmpgt_epi32(yv, v_h)));
    const __m256i esc = _mm256_cmpgt_epi32(d, v_two);
    n_esc += (uint32_t)__builtin_popcount((unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(esc)));
    const __m256i code = _mm256_min_epi32(_mm256_max_epi32(d, v_zero), v_three);
    __m256i rec = _mm256_or_si256(xv, _mm256_sll_epi32(yv, c_xb));
    rec = _mm256_or_si256(rec, _mm256_sll_epi32(_mm256_and_si256(pos, v_one), c_p));
    rec = _mm256_or_si256(rec, _mm256_sll_epi32(code, c_c));
    const __m256i b = _mm256_shuffle_epi8(rec, pick);  // 12 payload bytes at the bottom of each 128-bit half
    const __m128i lo = _mm256_castsi256_si128(b), hi = _mm256_extracti128_si256(b, 1);
    _mm_storeu_si128((__m128i*)out, lo);            // 16 bytes: the 4 spare ones are overwritten next
    _mm_storel_epi64((__m128i*)(out + 12), hi);     // 8 + 4 bytes: exactly the second half's payload
    const int tail = _mm_extract_epi32(hi, 2);
    memcpy(out + 20, &tail, 4);
    out += 24;
  }
  *n_esc_out = n_esc;
  int bad = 0;
  if (!_mm256_testz_si256(f_sort, f_sort)) bad |= 1;
  if (!_mm256_testz_si256(f_pol, f_pol)) bad |= 2;
  if (!_mm256_testz_si256(f_pix, f_pix)) bad |= 4;
  if (!_mm256_testz_si256(f_rng, f_rng)) bad |= 8;
  return bad;
}
template <typename TT>
inline bool try_block64(const uint16_t*, const uint16_t*, const TT*, const int8_t*, int64_t, uint32_t, uint32_t, int, uint32_t, uint32_t, bool, uint8_t*, uint32_t*, int*, bool) {
  return false;  // 64-bit timestamps: the scalar loop
}
template <>
inline bool try_block64<int32_t>(const uint16_t* x, const uint16_t* y, const int32_t* t, const int8_t* p, int64_t t_first, uint32_t Wd, uint32_t Hd, int xb,
                                 uint32_t sh_p, uint32_t sh_c, bool zero_neg, uint8_t* out, uint32_t* n_esc, int* bad, bool have_avx2) {
  if (!have_avx2) return false;
  *bad |= block64_avx2(x, y, t, p, (int32_t)t_first, Wd, Hd, xb, sh_p, sh_c, zero_neg, out, n_esc);
  return true;
}
#endif

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
  const bool zero_neg = j.zero_neg;
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
#ifdef EVREP_PACK_AVX2
    if (e1 - e0 == 64 && try_block64<TT>(x + e0, y + e0, t + e0, p + e0, t_first, Wd, Hd, xb, sh_p, sh_c, zero_neg, out, &n_esc, &bad, j.avx2)) {
      j.esc_count[b] = n_esc;
      continue;
    }
#endif
    // one event -> its 24-bit record (the first event of a block is compared with itself: difference 0)
    auto record = [&](int64_t i) -> uint32_t {
      const int64_t ti = (int64_t)t[i];
      const int64_t d = ti - (int64_t)t[i > e0 ? i - 1 : i];
      const uint32_t xi = x[i], yi = y[i];
      const int pi = p[i];
      bad |= (d < 0) | ((pi != 1 && pi != -1 && !(zero_neg && pi == 0)) << 1) | ((xi >= Wd || yi >= Hd) << 2) | ((d >= ((int64_t)1 << 32) || ti - t_first >= ((int64_t)1 << 31)) << 3);
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
                                            int64_t esc_capacity, int64_t* n_escapes, int zero_is_negative, int n_threads) {
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
  j.zero_neg = zero_is_negative != 0;
#ifdef EVREP_PACK_AVX2
  j.avx2 = __builtin_cpu_supports("avx2") && !getenv("EVREP_PACK_SCALAR");
#endif
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

// ---------------------------------------------------------------------------------------------
// formats 4 and 6 (the loader's half of evrep_unpack_events): one 32-bit word per event - x, y, a 2-bit polarity code and, in
// format 4, the offset to the smallest timestamp of the event's block of 64 - or that word plus a 16-bit offset (format 6,
// blocks of 256).  Any event order, any polarity in {-1, 0, 1}.
// ---------------------------------------------------------------------------------------------
namespace {

struct WordJob {
  const uint16_t* x;
  const uint16_t* y;
  const void* t;
  int t_bytes;
  const int8_t* p;
  const int64_t* offs;
  const int64_t* blk_prefix;
  int B, H, W, xb, yb, fmt, bs;
  uint32_t* word;
  uint16_t* dt16;
  int32_t* tbase;
  std::atomic<int> status{0};  // bit 0 offset too large for the format, bit 1 polarity outside {-1, 0, 1}, bit 2 pixel outside, bit 3 base outside int32
};

template <typename TT>
void word_blocks_t(WordJob& j, int64_t b0, int64_t b1) {
  const uint16_t* __restrict__ x = j.x;
  const uint16_t* __restrict__ y = j.y;
  const TT* __restrict__ t = (const TT*)j.t;
  const int8_t* __restrict__ p = j.p;
  const int64_t* __restrict__ offs = j.offs;
  const int64_t* __restrict__ blk_prefix = j.blk_prefix;
  uint32_t* __restrict__ word = j.word;
  uint16_t* __restrict__ dt16 = j.dt16;
  const uint32_t Wd = (uint32_t)j.W, Hd = (uint32_t)j.H;
  const int xb = j.xb, bs = j.bs, fmt = j.fmt;
  const uint32_t sh_p = (uint32_t)(j.xb + j.yb), sh_t = sh_p + 2u;
  const int64_t limit = fmt == 4 ? ((int64_t)1 << (30 - j.xb - j.yb)) : 65536;
  int w = (int)(std::upper_bound(blk_prefix, blk_prefix + j.B + 1, b0) - blk_prefix) - 1;
  int bad = 0;
  for (int64_t b = b0; b < b1; ++b) {
    while (b >= blk_prefix[w + 1]) ++w;
    const int64_t ws = offs[w], we = offs[w + 1];
    const int64_t e0 = ws + ((b - blk_prefix[w]) << bs), e1 = std::min(e0 + ((int64_t)1 << bs), we);
    const int64_t t_first = (int64_t)t[ws];
    int64_t lo = (int64_t)t[e0];
    for (int64_t i = e0 + 1; i < e1; ++i) lo = std::min(lo, (int64_t)t[i]);
    const int64_t base = lo - t_first;
    if (base < -((int64_t)1 << 31) || base >= ((int64_t)1 << 31)) bad |= 8;
    j.tbase[b] = (int32_t)base;
    for (int64_t i = e0; i < e1; ++i) {
      const int64_t d = (int64_t)t[i] - lo;
      const uint32_t xi = x[i], yi = y[i];
      const int pi = p[i];
      bad |= (d >= limit) | ((pi < -1 || pi > 1) << 1) | ((xi >= Wd || yi >= Hd) << 2);
      uint32_t v = xi | (yi << xb) | (((uint32_t)pi & 3u) << sh_p);
      if (fmt == 4) v |= (uint32_t)d << sh_t;
      else dt16[i] = (uint16_t)d;
      word[i] = v;
    }
  }
  if (bad) j.status.fetch_or(bad);
}

}  // namespace

extern "C" int64_t evrep_pack_host_blocks(const int64_t* win_offsets, int B, int fmt) {
  if (!win_offsets || B < 0 || (fmt != 3 && fmt != 4 && fmt != 6)) return -1;
  const int bs = fmt == 6 ? 8 : 6;
  int64_t nb = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t n = win_offsets[b + 1] - win_offsets[b];
    if (n < 0) return -1;
    nb += (n + ((int64_t)1 << bs) - 1) >> bs;
  }
  return nb;
}

extern "C" int evrep_pack_events_host(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets, int B,
                                      int H, int W, int fmt, uint32_t* word, uint16_t* dt16, int32_t* tbase, int n_threads) {
  using evrep::set_error;
  if (B < 0 || !win_offsets || (t_bytes != 4 && t_bytes != 8) || H < 1 || W < 1 || (fmt != 4 && fmt != 6)) { set_error("bad argument"); return EVREP_EINVAL; }
  const int xb = bits_for(W), yb = bits_for(H);
  if (fmt == 4 && xb + yb > 29) { set_error("wire format 4 holds x, y, the polarity code and a time offset in 32 bits; %d x %d leaves no room", W, H); return EVREP_EUNSUPPORTED; }
  if (xb + yb > 30) { set_error("sensor %d x %d does not fit a 32-bit event word", W, H); return EVREP_EUNSUPPORTED; }
  const int bs = fmt == 6 ? 8 : 6;
  std::vector<int64_t> blk_prefix((size_t)B + 1, 0);
  for (int b = 0; b < B; ++b) {
    const int64_t n = win_offsets[b + 1] - win_offsets[b];
    if (n < 0 || win_offsets[b] < 0) { set_error("win_offsets must be non-decreasing and non-negative"); return EVREP_EINVAL; }
    blk_prefix[(size_t)b + 1] = blk_prefix[(size_t)b] + ((n + ((int64_t)1 << bs) - 1) >> bs);
  }
  const int64_t n_blocks = blk_prefix[(size_t)B];
  const int64_t total = B ? win_offsets[B] : 0;
  if (total > 0 && (!x || !y || !t || !p || !word || !tbase || (fmt == 6 && !dt16))) { set_error("null event / output array"); return EVREP_EINVAL; }
  WordJob j;
  j.x = x; j.y = y; j.t = t; j.t_bytes = t_bytes; j.p = p; j.offs = win_offsets; j.blk_prefix = blk_prefix.data();
  j.B = B; j.H = H; j.W = W; j.xb = xb; j.yb = yb; j.fmt = fmt; j.bs = bs;
  j.word = word; j.dt16 = dt16; j.tbase = tbase;
  if (n_threads < 1) n_threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  auto run = [&j](int64_t b0, int64_t b1) {
    if (j.t_bytes == 4) word_blocks_t<int32_t>(j, b0, b1);
    else word_blocks_t<int64_t>(j, b0, b1);
  };
  if (n_threads <= 1 || n_blocks < 4096) {
    run(0, n_blocks);
  } else {
    std::vector<std::thread> th;
    const int64_t per = (n_blocks + n_threads - 1) / n_threads;
    for (int k = 0; k < n_threads; ++k) {
      const int64_t b0 = k * per, b1 = std::min(n_blocks, b0 + per);
      if (b0 >= b1) break;
      th.emplace_back([&run, b0, b1] { run(b0, b1); });
    }
    for (auto& t_ : th) t_.join();
  }
  const int bad = j.status.load();
  if (bad & 4) { set_error("event outside the sensor"); return EVREP_EINVAL; }
  if (bad & 2) { set_error("polarities must be in {-1, 0, 1}"); return EVREP_EINVAL; }
  if (bad & 1) { set_error("a block spans more time than wire format %d holds", fmt); return EVREP_EUNSUPPORTED; }
  if (bad & 8) { set_error("a block starts 2^31 us or more from its window's first event"); return EVREP_EUNSUPPORTED; }
  return EVREP_OK;
}
