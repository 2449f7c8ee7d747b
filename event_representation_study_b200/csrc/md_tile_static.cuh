// Mixed-density tile kernels with the accumulator plan as a compile-time constant: every group test, word index, limb
// count and channel formula folds, and the loops disappear.  Instantiated ahead of time for the ERGO-12 tuples
// (mixed_density.cu) and at run time, through NVRTC, for any other tuple a caller asks to specialise (md_jit.cu) - so the
// header must compile with nothing but md_device.cuh and md_plan.cuh in front of it (no system headers).
//
// Reference semantics: representations/representation_search/operations.py:15-89, mixed_density_event_stack.py:25-151.
#pragma once
#include "md_device.cuh"
#include "md_plan.cuh"

namespace evrep {

// compile-time integer sequences (std::integer_sequence without <utility>, which NVRTC does not have)
template <int... I>
struct iseq {};
template <int N, int... I>
struct make_iseq_t : make_iseq_t<N - 1, N - 1, I...> {};
template <int... I>
struct make_iseq_t<0, I...> { using type = iseq<I...>; };
template <int N>
using make_iseq = typename make_iseq_t<N>::type;

__device__ __forceinline__ uint32_t md_cnt(const MdPlan& P, const MdGroup& G, const uint32_t* a) {
  return P.packed ? ((a[G.w_cnt] >> G.cnt_shift) & 0xffffu) : a[G.w_cnt];
}
// number of events of the channel's class at this pixel (or just 0 / 1 when only "touched" is tracked)
__device__ __forceinline__ uint32_t md_count(const MdPlan& P, const MdChan& ch, const uint32_t* a) {
  if (ch.g_pos >= 0 && ch.g_neg >= 0 && ch.g_oth >= 0)  // "all events": every event bumped exactly one class counter
    return md_cnt(P, P.grp[ch.g_pos], a) + md_cnt(P, P.grp[ch.g_neg], a) + md_cnt(P, P.grp[ch.g_oth], a);
  const MdGroup& G = P.grp[ch.g_main];
  if (G.flags & G_CNT) return md_cnt(P, G, a);
  if (G.flags & G_MAX) return a[G.w_max] != 0u;
  if (G.flags & G_MIN) return a[G.w_min] != 0u;
  if (G.flags & G_PRES) return (a[P.w_pres] >> G.pres_bit) & 1u;
  return 0u;
}

// value of a multi-limb sum as a double (limbs are base-2^lw digits with 32-bit headroom): Horner in fp64, exact while the
// sum stays below 2^53 and correctly rounded to ~1e-16 relative beyond
__device__ __forceinline__ double md_limb_sum(const uint32_t* a, int w0, int nl, int lw) {
  const double radix = (double)(1u << lw);
  double s = (double)a[w0 + nl - 1];
  for (int l = nl - 2; l >= 0; --l) s = fma(s, radix, (double)a[w0 + l]);
  return s;
}

// The same channel formulas for the compile-time specialised kernels, trimmed for instruction count (the finalise phase
// is ~40 % of the headline kernel's instructions): reciprocals are single MUFU.RCP approximations (<= 1 ulp; every
// divisor is a non-negative integer, 0 gives inf and 0 * inf = NaN reproduces the reference's 0/0), packed plans
// (buckets below 65536 events) do the polarity variance in 32-bit integers, and sums of t are converted limb by limb in
// fp32 (limb sums stay below 2^32).  Every result stays within 4e-7 relative of the reference's fp64 value.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float md_value_fast(const MdPlan& P, const MdChan& ch, const uint32_t* a, float inv_delta, double delta, uint32_t delta_u,
                                               uint32_t has_m1) {
  if (!ch.valid) return 0.f;
  if (ch.func == EVREP_FUNC_POLARITY) {
    const uint32_t c1 = md_cnt(P, P.grp[ch.g_pos], a);
    const uint32_t cn = md_cnt(P, P.grp[ch.g_neg], a);
    const uint32_t cm = ((has_m1 >> ch.win) & 1u) ? cn : 0u;  // the "negative" class holds the p == 0 events when the window has no -1
    if (ch.agg == EVREP_AGG_SUM) return (float)((int)c1 - (int)cm);
    const uint32_t call = c1 + cn + md_cnt(P, P.grp[ch.g_oth], a);
    if (ch.agg == EVREP_AGG_MEAN) return call ? (float)((int)c1 - (int)cm) * rcp_approx((float)call) : 0.f;
    if (ch.agg == EVREP_AGG_VARIANCE) {  // mean(p^2) - mean(p)^2 = ((c1+cm) call - (c1-cm)^2) / call^2, exact in integers
      if (P.packed) {                    // call < 65536: everything fits 32 bits
        const int d = (int)c1 - (int)cm;
        const uint32_t num = (c1 + cm) * call - (uint32_t)(d * d);
        return call ? (float)num * rcp_approx((float)(call * call)) : 0.f;
      }
      const long long dd = (long long)((int)c1 - (int)cm);
      const unsigned long long num = (unsigned long long)(c1 + cm) * call - (unsigned long long)(dd * dd);
      return call ? __ull2float_rn(num) * rcp_approx(__ull2float_rn((unsigned long long)call * call)) : 0.f;
    }
    if (call == 0u) return 0.f;
    if (ch.agg == EVREP_AGG_MIN) return cm > 0 ? -1.f : (call - c1 - cm > 0 ? 0.f : 1.f);  // min of the raw polarities
    return c1 > 0 ? 1.f : (call - c1 - cm > 0 ? 0.f : -1.f);                               // max of the raw polarities
  }
  const bool is_count = (ch.func == EVREP_FUNC_COUNT || ch.func == EVREP_FUNC_COUNT_POS || ch.func == EVREP_FUNC_COUNT_NEG);
  if (ch.g_main < 0 && ch.g_pos < 0) return 0.f;  // variance of a constant
  if (is_count) {
    const uint32_t c = md_count(P, ch, a);
    return ch.agg == EVREP_AGG_SUM ? (float)c : (c ? 1.f : 0.f);  // torch_scatter leaves untouched pixels at 0
  }
  // timestamps: t_s = (t - t_min) / (t_max - t_min); delta == 0 gives NaN exactly like the reference
  const MdGroup& G = P.grp[ch.g_main];
  if (ch.agg == EVREP_AGG_MAX || ch.agg == EVREP_AGG_MIN) {
    const uint32_t w = ch.agg == EVREP_AGG_MAX ? a[G.w_max] : a[G.w_min];
    const uint32_t v = ch.agg == EVREP_AGG_MAX ? w - 1u : ~w;
    float r = (float)v * inv_delta;
    if (v == delta_u && delta_u) r = 1.f;  // the window's last event maps to exactly 1
    return w ? r : 0.f;
  }
  const uint32_t c = md_count(P, ch, a);
  if (c == 0u) return 0.f;
  if (ch.agg == EVREP_AGG_SUM || ch.agg == EVREP_AGG_MEAN) {
    float st = (float)a[G.w_st + P.nl1 - 1];
    for (int l = P.nl1 - 2; l >= 0; --l) st = fmaf(st, (float)(1u << P.lw), (float)a[G.w_st + l]);
    st *= inv_delta;
    return ch.agg == EVREP_AGG_SUM ? st : st * rcp_approx((float)c);
  }
  if (c == 1u && delta_u) return 0.f;  // a single event: t_s^2 - t_s^2, exactly 0 in the reference too (NaN when delta == 0)
  unsigned long long sti = a[G.w_st + P.nl1 - 1];
  for (int l = P.nl1 - 2; l >= 0; --l) sti = (sti << P.lw) + a[G.w_st + l];
  const double st = (double)sti;
  const double cd = (double)c * delta;
  const double st2 = md_limb_sum(a, G.w_st2, P.nl2, P.lw);
  // mean(t_s^2) - mean(t_s)^2 = (c sum(t^2) - sum(t)^2) / (c delta)^2
  return (float)fma((double)c, st2, -st * st) * rcp_approx((float)(cd * cd));
}

// ---------------------------------------------------------------------------------------------
// compile-time specialised variant (ERGO-12): same algorithm, the plan is a constant expression, so every
// group test, word index, limb count and channel formula is folded and the loops disappear.
// ---------------------------------------------------------------------------------------------
// CLS selects the groups an event can belong to: 0 = any (buckets not split by polarity), 1 = the event has p > 0
// (groups of class "all" and "positive"), 2 = it has not (classes "all", "negative", "neither").
// counter word WI of a packed plan: both 16-bit counters with ONE atomic (md_plan.cuh pairs counters that the same
// event tends to bump); wide plans have one counter per word
template <typename PS>
constexpr int md_cnt_group_at(int word, int shift) {  // the group whose counter lives in that half-word, or -1
  for (int g = 0; g < PS::value.G; ++g)
    if ((PS::value.grp[g].flags & G_CNT) && PS::value.grp[g].w_cnt == word && PS::value.grp[g].cnt_shift == shift) return g;
  return -1;
}
template <typename PS, int CLS>
constexpr bool md_group_live(int g) {  // can an event of class CLS belong to group g at all?
  if (g < 0) return false;
  const int gcls = PS::value.grp[g].bit >> 3;
  return !((CLS == 1 && gcls >= 2) || (CLS == 2 && gcls == 1));
}
template <typename PS, int WI, int CLS>
__device__ __forceinline__ void md_acc_cnt_word(uint32_t* a, uint32_t M) {
  constexpr int g_lo = md_cnt_group_at<PS>(WI, 0), g_hi = md_cnt_group_at<PS>(WI, 16);
  constexpr bool l_lo = md_group_live<PS, CLS>(g_lo), l_hi = md_group_live<PS, CLS>(g_hi);
  uint32_t inc = 0;
  if constexpr (l_lo) inc |= (M >> PS::value.grp[g_lo].bit) & 1u;
  if constexpr (l_hi) inc |= ((M >> PS::value.grp[g_hi].bit) & 1u) << 16;
  if constexpr (l_lo || l_hi) {
    if (inc) atomicAdd(a + WI, inc);
  }
}
template <typename PS, int CLS, int... WI>
__device__ __forceinline__ void md_acc_cnt_all(uint32_t* a, uint32_t M, iseq<WI...>) {
  (md_acc_cnt_word<PS, WI, CLS>(a, M), ...);
}
// highest counter word index + 1 (counter words are contiguous, after the presence word)
template <typename PS>
constexpr int md_cnt_words_end() {
  int e = 0;
  for (int g = 0; g < PS::value.G; ++g)
    if ((PS::value.grp[g].flags & G_CNT) && PS::value.grp[g].w_cnt + 1 > e) e = PS::value.grp[g].w_cnt + 1;
  return e;
}

// CLS selects the groups an event can belong to: 0 = any (buckets not split by polarity), 1 = the event has p > 0
// (groups of class "all" and "positive"), 2 = it has not (classes "all", "negative", "neither").
template <typename PS, int GI, int CLS>
__device__ __forceinline__ void md_acc_group(uint32_t* a, uint32_t M, uint32_t tt, unsigned long long tt2, uint32_t& pres) {
  constexpr MdGroup G = PS::value.grp[GI];
  constexpr int LW = PS::value.lw, NL1 = PS::value.nl1, NL2 = PS::value.nl2;
  constexpr uint32_t MASK = (1u << LW) - 1u;
  constexpr int gcls = G.bit >> 3;
  if constexpr ((CLS == 1 && gcls >= 2) || (CLS == 2 && gcls == 1)) return;
  if constexpr (G.flags == G_PRES) {  // presence only: no branch
    pres |= ((M >> G.bit) & 1u) << G.pres_bit;
    return;
  }
  if constexpr ((G.flags & ~G_CNT) == 0) return;  // counters are handled word by word (md_acc_cnt_word)
  if (!((M >> G.bit) & 1u)) return;
  if constexpr (G.flags & G_PRES) pres |= 1u << G.pres_bit;
  if constexpr (G.flags & G_MAX) atomicMax(a + G.w_max, tt + 1u);
  if constexpr (G.flags & G_MIN) atomicMax(a + G.w_min, ~tt);  // earliest timestamp as the maximum of ~t (tt < 2^31: never 0)
  if constexpr (G.flags & G_ST) {
#pragma unroll
    for (int l = 0; l < NL1; ++l) atomicAdd(a + G.w_st + l, (tt >> (l * LW)) & MASK);  // adding a zero limb is harmless
  }
  if constexpr (G.flags & G_ST2) {
#pragma unroll
    for (int l = 0; l < NL2; ++l) {
      const uint32_t limb = (uint32_t)(tt2 >> (l * LW)) & MASK;
      if (l < 2 || limb) atomicAdd(a + G.w_st2 + l, limb);  // the high limbs of t^2 are zero for most events
    }
  }
}
template <typename PS, int CLS, int... GI>
__device__ __forceinline__ void md_acc_all(uint32_t* a, uint32_t M, uint32_t tt, uint32_t& pres, iseq<GI...>) {
  const unsigned long long tt2 = (unsigned long long)tt * (unsigned long long)tt;
  md_acc_cnt_all<PS, CLS>(a, M, make_iseq<md_cnt_words_end<PS>()>{});
  (md_acc_group<PS, GI, CLS>(a, M, tt, tt2, pres), ...);
}

template <typename PS, int CI>
__device__ __forceinline__ float md_value_static(const uint32_t* a, float inv_delta, double delta, uint32_t delta_u, uint32_t has_m1) {
  constexpr MdPlan P = PS::value;
  constexpr MdChan ch = PS::value.ch[CI];
  return md_value_fast(P, ch, a, inv_delta, delta, delta_u, has_m1);
}
template <typename PS, int... CI>
__device__ __forceinline__ void md_finalise_static(const uint32_t* a, float inv_delta, double delta, uint32_t delta_u, uint32_t has_m1,
                                                   float (&o)[sizeof...(CI)], iseq<CI...>) {
  ((o[CI] = md_value_static<PS, CI>(a, inv_delta, delta, delta_u, has_m1)), ...);
}

// SBT windows (mixed_density_event_stack.py:76-107): the reference's float64 comparisons on t_s = (t - t_min) / (t_max - t_min),
// bit for bit (a zero interval gives NaN / inf: member of window 0 only).  Shared by the interpreted kernel, the
// specialised kernels and k_sbt_negsel.
__device__ __forceinline__ uint32_t md_sbt_wmask(uint32_t tt, double delta) {
  const double ts = (double)tt / delta;
  const double f = 1.0 / 3.0;
  uint32_t wmask = 1u;
  if (ts <= 1.0 * f && ts >= 0.0 * f) wmask |= 2u;
  if (ts <= 2.0 * f && ts >= 1.0 * f) wmask |= 4u;
  if (ts <= 3.0 * f && ts >= 2.0 * f) wmask |= 8u;
  if (ts <= 0.5) wmask |= 16u;
  if (ts <= 0.25) wmask |= 32u;
  if (ts <= 0.125) wmask |= 64u;
  if (ts <= 0.0625) wmask |= 128u;
  return wmask;
}
// prefetch slots past the end of a bucket: a record that touches nothing (SBN: window mask 0; SBT derives the mask from the
// timestamp, so there it must be the null record proper)
template <typename PS>
__device__ __forceinline__ uint2 md_padding_record() {
  return make_uint2(0u, PS::value.stacking == EVREP_STACK_SBT ? REC_NULL_META : 0u);
}
template <typename PS>
__device__ __forceinline__ bool md_record_live(uint32_t meta) {
  return PS::value.stacking == EVREP_STACK_SBT ? !rec_is_null(meta) : meta != 0u;
}

template <typename PS>
__device__ __forceinline__ void md_acc_presence(uint32_t* a, uint32_t pres) {
  if (pres) {  // only plans that still keep presence bits
    uint32_t* pw = a + PS::value.w_pres;
    if ((*(volatile uint32_t*)pw & pres) != pres) atomicOr(pw, pres);
  }
}
// An event whose SBN window mask is the compile-time constant WM: the membership word M is a constant too, so every group
// test of md_acc_all folds and what remains is the straight-line list of atomics this (zone, polarity class) owes - no
// bit tests, and none of the BSSY / BRA / BSYNC triples ptxas puts around a conditional shared-memory atomic.
template <typename PS, int CLS, uint32_t WM>
__device__ __forceinline__ void md_acc_fixed(uint32_t* a, uint32_t tt) {
  constexpr uint32_t M = CLS == 1 ? (WM | (WM << 8)) : (WM | (WM << 16));
  uint32_t pres = 0;
  md_acc_all<PS, CLS>(a, M, tt, pres, make_iseq<PS::value.G>{});
  md_acc_presence<PS>(a, pres);
}

template <typename PS, int CLS>
__device__ __forceinline__ void md_accumulate_static(uint32_t* acc, const uint2 r, int32_t tmin, uint32_t not_m1, double delta) {
  constexpr int STRIDE = PS::value.stride, G = PS::value.G;
  uint32_t* a = acc + (r.y & 0xffffu) * STRIDE;
  const uint32_t pc = (r.y >> 24) & 3u;
  const uint32_t tt = (uint32_t)((int32_t)r.x - tmin);
  uint32_t wmask = (r.y >> 16) & 0xffu;  // 0 for null and padding records: member of no window
  if constexpr (PS::value.stacking == EVREP_STACK_SBT) wmask = rec_is_null(r.y) ? 0u : md_sbt_wmask(tt, delta);
  if constexpr (PS::value.stacking == EVREP_STACK_SBN && CLS != 0) {
    // SBN windows are index ranges (mixed_density_event_stack.py:55-74): an event lies in one of seven zones, each with its own
    // constant mask, and the records of a bucket are in stream order, so the 32 events of a warp nearly always share a zone -
    // dispatch on the mask and run that zone's specialised list.  Class 2 qualifies when its events count as "negative" in every
    // window they belong to (p == -1, or p == 0 with no -1 in those windows: operations.py:59-61); anything else - and masks of
    // degenerate windows, where zone boundaries coincide - takes the generic walk below.
    if (CLS == 1 || pc == 3u || (wmask & ~not_m1) == 0u) {
      switch (wmask) {
        case 0u: return;  // null and padding records
        case 3u: md_acc_fixed<PS, CLS, 3u>(a, tt); return;      // first third
        case 5u: md_acc_fixed<PS, CLS, 5u>(a, tt); return;      // second third, first half
        case 21u: md_acc_fixed<PS, CLS, 21u>(a, tt); return;    // second third, second half
        case 25u: md_acc_fixed<PS, CLS, 25u>(a, tt); return;    // last third, third quarter
        case 57u: md_acc_fixed<PS, CLS, 57u>(a, tt); return;    // last third, seventh eighth
        case 121u: md_acc_fixed<PS, CLS, 121u>(a, tt); return;  // last third, last eighth
        case 113u: md_acc_fixed<PS, CLS, 113u>(a, tt); return;  // the n % 3 events after the last third
        default: break;
      }
    }
  }
  uint32_t M;
  if constexpr (CLS == 1) {
    M = wmask | (wmask << 8);
  } else {
    // "negative" events of a window: p == -1, or p == 0 when the window holds no -1 (operations.py:59-61,78-80)
    const uint32_t posm = (CLS == 0 && pc == 1u) ? wmask : 0u;
    const uint32_t negm = (pc == 3u) ? wmask : (pc == 0u ? (wmask & not_m1) : 0u);
    M = wmask | (posm << 8) | (negm << 16) | ((wmask & ~(posm | negm)) << 24);
  }
  uint32_t pres = 0;
  md_acc_all<PS, CLS>(a, M, tt, pres, make_iseq<G>{});
  md_acc_presence<PS>(a, pres);
}
// record i of a bucket pair whose first n_pos records are the p > 0 events (n_pos = 0xffffffff: not split, any class)
template <typename PS, bool SPLIT>
__device__ __forceinline__ void md_accumulate_at(uint32_t* acc, const uint2 r, uint32_t i, uint32_t n_pos, int32_t tmin, uint32_t not_m1, double delta) {
  if constexpr (SPLIT) {
    if (i < n_pos) md_accumulate_static<PS, 1>(acc, r, tmin, not_m1, delta);
    else md_accumulate_static<PS, 2>(acc, r, tmin, not_m1, delta);
  } else {
    md_accumulate_static<PS, 0>(acc, r, tmin, not_m1, delta);
  }
}

// TP = pixels per tile (compile time here), PPT = pixels per thread.
// Persistent CTAs (three per SM) pull (window, tile) buckets from a ticket counter.  A warp owns PPT * 32
// consecutive pixels of the tile, i.e. one contiguous slab of the accumulator array.  Timeline of one bucket:
//   atomics over the bucket's records (already in registers: they were fetched while the previous bucket was
//   being finalised)                                                     -> __syncthreads (A)
//   per warp, no CTA-wide barrier: finalise its pixels into registers, repack the 12 floats of each pixel
//   contiguously at the head of its own slab (over accumulators it has already consumed), hand the slab's
//   output (PPT * 32 * C * 4 bytes, contiguous in the output tensor) to the TMA engine with ONE
//   cp.async.bulk shared -> global, prefetch the next bucket's records, wait until the engine has read the
//   slab, zero the slab                                                  -> __syncthreads (B)
// Two CTA barriers per bucket; header and record loads of the next bucket are issued early, so no global-load
// latency sits on the critical path after the first bucket.
struct TileHdr {
  int b, pix0;
  uint32_t count, n_pos, has_m1, delta_u;
  int32_t tmin;
  const uint2* rec;
};

template <bool SPLIT>
__device__ __forceinline__ TileHdr md_load_hdr(int id, const Geom& g, int TP, const uint2* records, const uint32_t* base,
                                               const uint32_t* hist, const WinParams* wp) {
  TileHdr h;
  h.b = (int)(((unsigned long long)(uint32_t)id * g.t_magic) >> 44);  // id / T, exact for id < 2^32 and T <= 4096
  h.pix0 = (id - h.b * g.T) * TP;
  const WinParams* w = wp + h.b;
  if constexpr (SPLIT) {  // two buckets per tile, p > 0 first, contiguous
    const uint2 c2 = __ldg(reinterpret_cast<const uint2*>(hist) + id);
    h.n_pos = c2.x;
    h.count = c2.x + c2.y;
    h.rec = records + w->start + __ldg(base + 2 * id);
  } else {
    h.n_pos = 0xffffffffu;
    h.count = __ldg(hist + id);
    h.rec = records + w->start + __ldg(base + id);
  }
  h.tmin = w->tmin_rel;
  const int32_t tmax = w->tmax_rel;
  h.delta_u = tmax >= h.tmin ? (uint32_t)(tmax - h.tmin) : 0u;
  h.has_m1 = w->has_m1;
  return h;
}

// finalise + repack + TMA store of one warp's slab; leaves the bulk store in flight (lane 0 owns the bulk group)
template <typename PS, int TP>
__device__ __forceinline__ void md_finalise_store_warp(uint32_t* slab, const TileHdr& h, const Geom& g, float* __restrict__ out) {
  constexpr int STRIDE = PS::value.stride, C = PS::value.C, PPT = TP / TILE_THREADS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double delta = (double)h.delta_u;
  const float inv_delta = rcp_approx((float)h.delta_u);
  float4* stage = reinterpret_cast<float4*>(slab);  // [PPT * 32][C] floats, contiguous = the global layout of the slab's pixels
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    float o[C];
    md_finalise_static<PS>(slab + (k * 32 + lane) * STRIDE, inv_delta, delta, h.delta_u, h.has_m1, o, make_iseq<C>{});
    __syncwarp();  // every lane has read its accumulators: rows k*32 .. may now be overwritten (C <= STRIDE keeps row k+1 intact)
    if constexpr ((C & 3) == 0) {
#pragma unroll
      for (int q = 0; q < C / 4; ++q) stage[(k * 32 + lane) * (C / 4) + q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    } else {  // run-time specialised tuples of any channel count (callers check that H * W * C is a multiple of 4: 16-byte bulk stores)
      float* st = reinterpret_cast<float*>(slab) + (k * 32 + lane) * C;
#pragma unroll
      for (int c = 0; c < C; ++c) st[c] = o[c];
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    const int p0 = warp * (PPT * 32);
    const int np = min(PPT * 32, min(TP, g.HW - h.pix0) - p0);
    if (np > 0) {
      float* dst = out + ((size_t)h.b * g.HW + h.pix0 + p0) * C;
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(slab);
      const uint32_t bytes = (uint32_t)np * C * 4u;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// the warp's slab may be written again once the TMA engine has read it
__device__ __forceinline__ void md_wait_store_warp() {
  if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}
template <int WORDS>
__device__ __forceinline__ void md_zero_slab(uint32_t* slab) {
  static_assert(WORDS % 4 == 0, "slab size");
  uint4* a4 = reinterpret_cast<uint4*>(slab);
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < (WORDS / 4 + 31) / 32; ++i)
    if (i * 32 + lane < WORDS / 4) a4[i * 32 + lane] = make_uint4(0, 0, 0, 0);
}

// LIGHT_ONLY: the plan is a packed one (16-bit counters / limbs); buckets with >= 65536 events are left to
// k_md_tile_heavy, which runs the wide plan on them afterwards.
template <typename PS, int TP, bool LIGHT_ONLY, bool SPLIT>
__global__ void __launch_bounds__(TILE_THREADS, (PS::value.stride * TP * 4 <= 74 * 1024) ? 3 : 2)
    k_md_tile_static(const uint2* __restrict__ records, const uint32_t* __restrict__ base, const uint32_t* __restrict__ hist,
                     const WinParams* __restrict__ wp, const Geom g, uint32_t* __restrict__ ticket, float* __restrict__ out) {
  extern __shared__ __align__(128) uint32_t acc[];
  __shared__ int s_next;
  constexpr int STRIDE = PS::value.stride, C = PS::value.C, PPT = TP / TILE_THREADS;
  constexpr int SLAB = PPT * 32 * STRIDE;  // accumulator words of one warp's pixels
  constexpr int PRE = 3;
  static_assert(TP % TILE_THREADS == 0, "whole pixels per thread");
  static_assert(C <= STRIDE, "outputs must fit the accumulator footprint");
  static_assert((SLAB * 4) % 16 == 0, "slabs must start on 16-byte boundaries");
  static_assert(!LIGHT_ONLY || PS::value.packed, "only packed plans have an event limit");
  const int tid = threadIdx.x;
  const int n_tiles = g.B * g.T;
  // bucket order: blockIdx.x, blockIdx.x + gridDim.x, then tickets 2*gridDim.x + k; the ticket for the bucket
  // after next is requested a whole iteration early so that its round trip never stalls the CTA
  int cur = blockIdx.x, nxt = blockIdx.x + (int)gridDim.x;
  if (cur >= n_tiles) return;
  uint32_t* slab = acc + (tid >> 5) * SLAB;
  md_zero_slab<SLAB>(slab);  // before the wait: runs while k_bin drains
  pdl_wait();                // k_bin's records and bucket tables
  pdl_trigger();

  // buckets are taken last window first: k_bin wrote the last windows' records most recently, so they are the ones
  // still in L2 when this kernel starts
  TileHdr h = md_load_hdr<SPLIT>(n_tiles - 1 - cur, g, TP, records, base, hist, wp);
  uint2 pre[PRE];
#pragma unroll
  for (int j = 0; j < PRE; ++j) {
    const uint32_t i = tid + j * TILE_THREADS;
    pre[j] = i < h.count ? __ldg(h.rec + i) : md_padding_record<PS>();
  }
  __syncthreads();

  while (true) {
    const bool skip = LIGHT_ONLY && h.count >= MD_PACKED_LIMIT;  // CTA-uniform
    int my_ticket = 0;
    if (tid == 0) my_ticket = (int)atomicAdd(ticket, 1u);  // consumed after the atomics phase
    const bool more = nxt < n_tiles;
    TileHdr hn = h;
    if (more) hn = md_load_hdr<SPLIT>(n_tiles - 1 - nxt, g, TP, records, base, hist, wp);  // in flight during the atomics below

    if (!skip) {
      const uint32_t not_m1 = ~h.has_m1;
      const double delta = (double)h.delta_u;  // (only the SBT window test reads it)
#pragma unroll
      for (int j = 0; j < PRE; ++j)
        if (md_record_live<PS>(pre[j].y)) md_accumulate_at<PS, SPLIT>(acc, pre[j], tid + j * TILE_THREADS, h.n_pos, h.tmin, not_m1, delta);
      for (uint32_t i = tid + PRE * TILE_THREADS; i < h.count; i += TILE_THREADS)
        md_accumulate_at<PS, SPLIT>(acc, __ldg(h.rec + i), i, h.n_pos, h.tmin, not_m1, delta);
    }
#pragma unroll
    for (int j = 0; j < PRE; ++j) {  // next bucket's records: in flight during finalise + store
      const uint32_t i = tid + j * TILE_THREADS;
      pre[j] = (more && i < hn.count) ? __ldg(hn.rec + i) : md_padding_record<PS>();
    }
    if (tid == 0) s_next = my_ticket + 2 * (int)gridDim.x;
    __syncthreads();  // (A) every record of the bucket is accumulated
    const int nn = s_next;

    if (!skip) {
      md_finalise_store_warp<PS, TP>(slab, h, g, out);
      md_wait_store_warp();  // shared memory must outlive the engine's read (also at the end of the kernel)
      md_zero_slab<SLAB>(slab);
    }
    if (!more) break;
    h = hn;
    cur = nxt;
    nxt = nn;
    __syncthreads();  // (B) every slab is zero again; s_next may be rewritten
  }
}

// The buckets a packed plan must not touch (>= 65536 events: one hot tile), with the wide plan, one at a time.
template <typename PS, int TP, bool SPLIT>
__global__ void __launch_bounds__(TILE_THREADS, 2) k_md_tile_heavy(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                                   const uint32_t* __restrict__ hist, const WinParams* __restrict__ wp,
                                                                   const Geom g, float* __restrict__ out) {
  extern __shared__ __align__(128) uint32_t acc[];
  __shared__ uint32_t s_heavy[TILE_THREADS / 32];
  constexpr int STRIDE = PS::value.stride;
  const int tid = threadIdx.x;
  const int n_tiles = g.B * g.T;
  pdl_wait();  // the light kernel: its CTAs skip the heavy buckets, and every kernel of the chain waits for its predecessor
  pdl_trigger();
  for (int base_id = blockIdx.x * TILE_THREADS; base_id < n_tiles; base_id += gridDim.x * TILE_THREADS) {
    const int id = base_id + tid;
    bool heavy = false;
    if (id < n_tiles) heavy = (SPLIT ? __ldg(hist + 2 * id) + __ldg(hist + 2 * id + 1) : __ldg(hist + id)) >= MD_PACKED_LIMIT;
    const uint32_t m = __ballot_sync(0xffffffffu, heavy);
    if ((tid & 31) == 0) s_heavy[tid >> 5] = m;
    __syncthreads();
    for (int wd = 0; wd < TILE_THREADS / 32; ++wd) {
      uint32_t bits = s_heavy[wd];  // same value in every thread: uniform control flow
      while (bits) {
        const int tile_id = base_id + wd * 32 + (__ffs(bits) - 1);
        bits &= bits - 1;
        const TileHdr h = md_load_hdr<SPLIT>(tile_id, g, TP, records, base, hist, wp);
        uint4* a4 = reinterpret_cast<uint4*>(acc);
        for (int i = tid; i < (STRIDE * TP + 3) / 4; i += TILE_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const uint32_t not_m1 = ~h.has_m1;
        const double delta = (double)h.delta_u;
        for (uint32_t i = tid; i < h.count; i += TILE_THREADS) md_accumulate_at<PS, SPLIT>(acc, __ldg(h.rec + i), i, h.n_pos, h.tmin, not_m1, delta);
        __syncthreads();
        md_finalise_store_warp<PS, TP>(acc + (tid >> 5) * (TP / TILE_THREADS * 32 * STRIDE), h, g, out);
        md_wait_store_warp();
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

}  // namespace evrep
