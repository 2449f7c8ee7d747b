// MixedDensityEventStack / ERGO-12: per-tile shared-memory reduction + fused finalise.
//
// One CTA owns one (window, tile) bucket produced by binning.cu.  It keeps, for every pixel of the tile,
// the integer accumulators the requested channels need (event counts, presence bits, latest timestamp,
// and exact multi-limb sums of t and t^2), updates them with native 32-bit shared-memory atomics, and
// then turns them into the C float32 channels of its output slice, written once, coalesced.
//
// Reference semantics: representations/representation_search/operations.py:15-89 (what each
// (function, aggregation) pair computes), mixed_density_event_stack.py:25-151 (windows, t normalisation,
// swallow-and-zero), optimized_representation.py:86-134 (the ERGO-12 tuple).
#include <string.h>

#include <utility>

#include "evrep_common.cuh"
#include "md_plan.cuh"

namespace evrep {

size_t md_tile_smem_bytes(const MdPlan& plan, int tile_px) { return align_up((size_t)plan.stride * (size_t)tile_px * sizeof(uint32_t), 16); }

// ---------------------------------------------------------------------------------------------
// host: (window, function, aggregation) tuple -> accumulator plan (md_plan.cuh does the work)
// ---------------------------------------------------------------------------------------------
int build_md_plan(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int64_t n_max, MdPlan* out) {
  if (C < 1 || C > EVREP_MAX_CHANNELS) {
    set_error("C = %d outside 1..%d", C, EVREP_MAX_CHANNELS);
    return EVREP_EINVAL;
  }
  if (stacking != EVREP_STACK_SBN && stacking != EVREP_STACK_SBT) {
    set_error("stacking must be EVREP_STACK_SBN or EVREP_STACK_SBT");
    return EVREP_EINVAL;
  }
  int lw = md_limb_width(n_max);
  // ERGO-12 has kernels specialised at compile time for a menu of limb widths: round down onto the menu
  out->static_id = 0;
  if (C == 12 && stacking == EVREP_STACK_SBN) {
    int ver = 0;
    if (!memcmp(win, kErgoWin2, 12) && !memcmp(func, kErgoFunc2, 12) && !memcmp(agg, kErgoAgg2, 12)) ver = 2;
    if (!memcmp(win, kErgoWin1, 12) && !memcmp(func, kErgoFunc1, 12) && !memcmp(agg, kErgoAgg1, 12)) ver = 1;
    if (ver) {
      int pick = 0;
      for (int m : kErgoLimbMenu)
        if (m <= lw) { pick = m; break; }
      if (pick) {
        lw = pick;
        if (md_plan_build(win, func, agg, C, stacking, lw, *out)) return EVREP_EUNSUPPORTED;
        out->static_id = ver * 100 + lw;
        return EVREP_OK;
      }
    }
  }
  if (md_plan_build(win, func, agg, C, stacking, lw, *out)) {
    set_error("mixed-density plan needs more than 250 accumulator words per pixel");
    return EVREP_EUNSUPPORTED;
  }
  return EVREP_OK;
}

// ---------------------------------------------------------------------------------------------
// device
//
// Shared-memory layout: array of structures, pixel p owns acc[p*stride .. p*stride+stride) with an ODD
// stride >= max(words, C), so that (a) atomics of different pixels spread over all banks, (b) a thread can
// finalise its pixel in place (outputs overwrite the pixel's own dead accumulators) and (c) the CTA then
// streams the tile's output slice to global memory with fully coalesced 16-byte stores.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t md_touched(const MdPlan& P, const MdGroup& G, const uint32_t* a) {
  if (G.flags & G_CNT) return a[G.w_cnt];
  if (G.flags & G_MAX) return a[G.w_max] != 0u;
  if (G.flags & G_PRES) return (a[P.w_pres] >> G.pres_bit) & 1u;
  return 0u;
}

// exact value of a multi-limb sum as a double (limbs are base-2^lw digits with 32-bit headroom)
__device__ __forceinline__ double md_limb_sum(const uint32_t* a, int w0, int nl, int lw) {
  // low three limbs and the rest are each combined exactly in 64-bit integers, then joined in fp64
  unsigned long long lo = 0, hi = 0;
  for (int l = 0; l < nl && l < 3; ++l) lo += (unsigned long long)a[w0 + l] << (l * lw);
  for (int l = 3; l < nl; ++l) hi += (unsigned long long)a[w0 + l] << ((l - 3) * lw);
  double s = (double)lo;
  if (nl > 3) s = fma((double)hi, (double)(1ull << (3 * lw)), s);
  return s;
}

// One channel of one pixel.  Numerators and denominators are computed exactly (integers / fp64), then divided
// once in fp32 with the hardware reciprocal (<= 2 ulp): no fp64 division and no cancellation after rounding,
// within 4e-7 relative of the reference's fp64 result.  0 * rcp(0) = NaN reproduces the reference's 0/0.
__device__ __forceinline__ float md_value(const MdPlan& P, const MdChan& ch, const uint32_t* a, float inv_delta, double delta, uint32_t delta_u,
                                          uint32_t has_m1) {
  if (!ch.valid) return 0.f;
  if (ch.func == EVREP_FUNC_POLARITY) {
    const int c1 = (int)a[P.grp[ch.g_pos].w_cnt];
    const int cm = ((has_m1 >> ch.win) & 1u) ? (int)a[P.grp[ch.g_neg].w_cnt] : 0;
    if (ch.agg == EVREP_AGG_SUM) return (float)(c1 - cm);
    const int call = (int)a[P.grp[ch.g_all].w_cnt];
    if (call == 0) return 0.f;
    if (ch.agg == EVREP_AGG_MEAN) return __fdividef((float)(c1 - cm), (float)call);
    if (ch.agg == EVREP_AGG_VARIANCE) {  // mean(p^2) - mean(p)^2 = ((c1+cm) call - (c1-cm)^2) / call^2, numerator exact in fp64
      const double dc = (double)call, dd = (double)(c1 - cm);
      return __fdividef((float)fma((double)(c1 + cm), dc, -dd * dd), (float)(dc * dc));
    }
    return c1 > 0 ? 1.f : (call - c1 - cm > 0 ? 0.f : -1.f);  // max of the raw polarities
  }
  if (ch.g_main < 0) return 0.f;
  const MdGroup& G = P.grp[ch.g_main];
  const uint32_t c = md_touched(P, G, a);
  if (c == 0u) return 0.f;  // torch_scatter leaves untouched pixels at 0
  const bool is_count = (ch.func == EVREP_FUNC_COUNT || ch.func == EVREP_FUNC_COUNT_POS || ch.func == EVREP_FUNC_COUNT_NEG);
  if (is_count) return ch.agg == EVREP_AGG_SUM ? (float)c : 1.f;
  // timestamps: t_s = (t - t_min) / (t_max - t_min); delta == 0 gives NaN exactly like the reference
  if (ch.agg == EVREP_AGG_MAX) {
    const uint32_t v = a[G.w_max] - 1u;
    return (v == delta_u && delta_u) ? 1.f : (float)v * inv_delta;  // the window's last event maps to exactly 1
  }
  const double st = md_limb_sum(a, G.w_st, P.nl1, P.lw);
  if (ch.agg == EVREP_AGG_SUM) return (float)st * inv_delta;
  const double cd = (double)c * delta;
  if (ch.agg == EVREP_AGG_MEAN) return __fdividef((float)st, (float)cd);
  const double st2 = md_limb_sum(a, G.w_st2, P.nl2, P.lw);
  // mean(t_s^2) - mean(t_s)^2 = (c sum(t^2) - sum(t)^2) / (c delta)^2
  return __fdividef((float)fma((double)c, st2, -st * st), (float)(cd * cd));
}

template <int CMAX>
__global__ void __launch_bounds__(TILE_THREADS) k_md_tile(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                          const uint32_t* __restrict__ cursor, const WinParams* __restrict__ wp,
                                                          const __grid_constant__ MdPlan P, const Geom g, float* __restrict__ out) {
  extern __shared__ __align__(128) uint32_t acc[];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int TP = g.tile_px;
  const int pix0 = tile << g.tile_shift;
  const int npix = min(TP, g.HW - pix0);
  const int stride = P.stride;

  {  // zero the accumulators
    uint4* a4 = reinterpret_cast<uint4*>(acc);
    const int n4 = (stride * TP + 3) / 4;
    for (int i = tid; i < n4; i += TILE_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
  }
  const WinParams w = wp[b];
  const uint32_t count = cursor[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];
  const int32_t tmin = w.tmin_rel;
  const uint32_t delta_u = (w.tmax_rel >= w.tmin_rel) ? (uint32_t)(w.tmax_rel - w.tmin_rel) : 0u;
  const double delta = (double)delta_u;
  const float inv_delta = 1.f / (float)delta_u;
  const uint32_t limb_mask = (1u << P.lw) - 1u;  // lw <= 31
  __syncthreads();

  for (uint32_t i = tid; i < count; i += TILE_THREADS) {
    const uint2 r = __ldg(rec + i);
    uint32_t* a = acc + (r.y & 0xffffu) * stride;
    const uint32_t pc = (r.y >> 24) & 3u;
    const uint32_t tt = (uint32_t)((int32_t)r.x - tmin);  // t - t_min, < 2^31
    uint32_t wmask;
    if (P.stacking == EVREP_STACK_SBN) {
      wmask = (r.y >> 16) & 0xffu;
    } else {
      // SBT windows (mixed_density_event_stack.py:76-107): float64 comparisons on t_s
      const double ts = (double)tt / delta;
      const double f = 1.0 / 3.0;
      wmask = 1u;
      if (ts <= 1.0 * f && ts >= 0.0 * f) wmask |= 2u;
      if (ts <= 2.0 * f && ts >= 1.0 * f) wmask |= 4u;
      if (ts <= 3.0 * f && ts >= 2.0 * f) wmask |= 8u;
      if (ts <= 0.5) wmask |= 16u;
      if (ts <= 0.25) wmask |= 32u;
      if (ts <= 0.125) wmask |= 64u;
      if (ts <= 0.0625) wmask |= 128u;
    }
    // "negative" events of a window: p == -1, or p == 0 when the window holds no -1 (operations.py:59-61,78-80)
    const uint32_t posm = (pc == 1u) ? wmask : 0u;
    const uint32_t negm = (pc == 3u) ? wmask : (pc == 0u ? (wmask & ~w.has_m1) : 0u);
    const uint32_t M = wmask | (posm << 8) | (negm << 16);
    uint32_t pres = 0;
    for (int gi = 0; gi < P.G; ++gi) {
      const MdGroup G = P.grp[gi];
      if (!((M >> G.bit) & 1u)) continue;
      if (G.flags & G_CNT) atomicAdd(a + G.w_cnt, 1u);
      if (G.flags & G_PRES) pres |= 1u << G.pres_bit;
      if (G.flags & G_MAX) atomicMax(a + G.w_max, tt + 1u);
      if (G.flags & G_ST) {
        uint32_t v = tt;
        for (int l = 0; l < P.nl1 && v; ++l, v >>= P.lw) {
          const uint32_t limb = v & limb_mask;
          if (limb) atomicAdd(a + G.w_st + l, limb);
        }
      }
      if (G.flags & G_ST2) {
        unsigned long long v = (unsigned long long)tt * (unsigned long long)tt;
        for (int l = 0; l < P.nl2 && v; ++l, v >>= P.lw) {
          const uint32_t limb = (uint32_t)v & limb_mask;
          if (limb) atomicAdd(a + G.w_st2 + l, limb);
        }
      }
    }
    if (pres) {
      uint32_t* pw = a + P.w_pres;
      if ((*(volatile uint32_t*)pw & pres) != pres) atomicOr(pw, pres);
    }
  }
  __syncthreads();

  // finalise in place: thread -> pixel, channels in lock-step across the warp (no divergence between kinds)
  const int C = P.C;
  for (int p = tid; p < npix; p += TILE_THREADS) {
    uint32_t* a = acc + p * stride;
    float o[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) o[c] = md_value(P, P.ch[c], a, inv_delta, delta, delta_u, w.has_m1);
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) a[c] = __float_as_uint(o[c]);
  }
  __syncthreads();

  // stream the tile's slice of the (B, H, W, C) output: consecutive threads write consecutive 16 bytes
  float* dst = out + ((size_t)b * g.HW + pix0) * C;
  if ((C & 3) == 0) {
    const int q_per_px = C >> 2, n_q = npix * q_per_px;
    float4* dst4 = reinterpret_cast<float4*>(dst);  // (b*HW + pix0)*C*4 bytes: multiple of 16 because C % 4 == 0
    for (int e = tid; e < n_q; e += TILE_THREADS) {
      const int p = e / q_per_px, q = e - p * q_per_px;
      const uint32_t* a = acc + p * stride + 4 * q;
      __stcs(dst4 + e, make_float4(__uint_as_float(a[0]), __uint_as_float(a[1]), __uint_as_float(a[2]), __uint_as_float(a[3])));
    }
  } else {
    const int n_el = npix * C;
    for (int e = tid; e < n_el; e += TILE_THREADS) {
      const int p = e / C, c = e - p * C;
      __stcs(dst + e, __uint_as_float(acc[p * stride + c]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// compile-time specialised variant (ERGO-12): same algorithm, the plan is a constant expression, so every
// group test, word index, limb count and channel formula is folded and the loops disappear.
// ---------------------------------------------------------------------------------------------
template <typename PS, int GI>
__device__ __forceinline__ void md_acc_group(uint32_t* a, uint32_t M, uint32_t tt, uint32_t& pres) {
  constexpr MdGroup G = PS::value.grp[GI];
  constexpr int LW = PS::value.lw, NL1 = PS::value.nl1, NL2 = PS::value.nl2;
  constexpr uint32_t MASK = (1u << LW) - 1u;
  if (!((M >> G.bit) & 1u)) return;
  if constexpr (G.flags & G_CNT) atomicAdd(a + G.w_cnt, 1u);
  if constexpr (G.flags & G_PRES) pres |= 1u << G.pres_bit;
  if constexpr (G.flags & G_MAX) atomicMax(a + G.w_max, tt + 1u);
  if constexpr (G.flags & G_ST) {
#pragma unroll
    for (int l = 0; l < NL1; ++l) {
      const uint32_t limb = (tt >> (l * LW)) & MASK;
      if (limb) atomicAdd(a + G.w_st + l, limb);
    }
  }
  if constexpr (G.flags & G_ST2) {
    const unsigned long long v = (unsigned long long)tt * (unsigned long long)tt;
#pragma unroll
    for (int l = 0; l < NL2; ++l) {
      const uint32_t limb = (uint32_t)(v >> (l * LW)) & MASK;
      if (limb) atomicAdd(a + G.w_st2 + l, limb);
    }
  }
}
template <typename PS, int... GI>
__device__ __forceinline__ void md_acc_all(uint32_t* a, uint32_t M, uint32_t tt, uint32_t& pres, std::integer_sequence<int, GI...>) {
  (md_acc_group<PS, GI>(a, M, tt, pres), ...);
}

template <typename PS, int CI>
__device__ __forceinline__ float md_value_static(const uint32_t* a, float inv_delta, double delta, uint32_t delta_u, uint32_t has_m1) {
  constexpr MdPlan P = PS::value;
  constexpr MdChan ch = PS::value.ch[CI];
  return md_value(P, ch, a, inv_delta, delta, delta_u, has_m1);
}
template <typename PS, int... CI>
__device__ __forceinline__ void md_finalise_static(const uint32_t* a, float inv_delta, double delta, uint32_t delta_u, uint32_t has_m1,
                                                   float (&o)[sizeof...(CI)], std::integer_sequence<int, CI...>) {
  ((o[CI] = md_value_static<PS, CI>(a, inv_delta, delta, delta_u, has_m1)), ...);
}

template <typename PS>
__device__ __forceinline__ void md_accumulate_static(uint32_t* acc, const uint2 r, int32_t tmin, uint32_t not_m1) {
  constexpr int STRIDE = PS::value.stride, G = PS::value.G;
  uint32_t* a = acc + (r.y & 0xffffu) * STRIDE;
  const uint32_t pc = (r.y >> 24) & 3u;
  const uint32_t tt = (uint32_t)((int32_t)r.x - tmin);
  const uint32_t wmask = (r.y >> 16) & 0xffu;
  const uint32_t posm = (pc == 1u) ? wmask : 0u;
  const uint32_t negm = (pc == 3u) ? wmask : (pc == 0u ? (wmask & not_m1) : 0u);
  const uint32_t M = wmask | (posm << 8) | (negm << 16);
  uint32_t pres = 0;
  md_acc_all<PS>(a, M, tt, pres, std::make_integer_sequence<int, G>{});
  if (pres) {
    uint32_t* pw = a + PS::value.w_pres;
    if ((*(volatile uint32_t*)pw & pres) != pres) atomicOr(pw, pres);
  }
}

// TP = pixels per tile (compile time here), PPT = pixels per thread.
// Timeline of one CTA: issue the first record loads -> zero the accumulators while they fly -> atomics ->
// finalise every pixel into registers -> barrier -> repack the tile's output slice contiguously in shared
// memory (it overwrites the dead accumulators) -> one thread per 12 KB hands it to the TMA engine
// (cp.async.bulk shared -> global), which streams it out while the SM's other CTA computes.
template <typename PS, int TP>
__global__ void __launch_bounds__(TILE_THREADS, 2) k_md_tile_static(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                                    const uint32_t* __restrict__ cursor, const WinParams* __restrict__ wp,
                                                                    const Geom g, float* __restrict__ out) {
  extern __shared__ __align__(128) uint32_t acc[];
  constexpr int STRIDE = PS::value.stride, C = PS::value.C;
  constexpr int PPT = TP / TILE_THREADS, PRE = 3;
  static_assert(PS::value.stacking == EVREP_STACK_SBN && (C & 3) == 0 && TP % TILE_THREADS == 0, "static path: SBN, C % 4 == 0");
  static_assert(C * 4 <= STRIDE * 4, "outputs must fit the accumulator footprint");
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int pix0 = tile * TP;
  const int npix = min(TP, g.HW - pix0);
  const WinParams w = wp[b];
  const uint32_t count = cursor[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];

  uint2 pre[PRE];
#pragma unroll
  for (int j = 0; j < PRE; ++j) {
    const uint32_t i = tid + j * TILE_THREADS;
    pre[j] = i < count ? __ldg(rec + i) : make_uint2(0u, 0u);  // meta 0: member of no window, touches nothing
  }
  {
    uint4* a4 = reinterpret_cast<uint4*>(acc);
    constexpr int N4 = (STRIDE * TP + 3) / 4;
#pragma unroll 4
    for (int i = tid; i < N4; i += TILE_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
  }
  const int32_t tmin = w.tmin_rel;
  const uint32_t delta_u = (w.tmax_rel >= w.tmin_rel) ? (uint32_t)(w.tmax_rel - w.tmin_rel) : 0u;
  const double delta = (double)delta_u;
  const float inv_delta = 1.f / (float)delta_u;
  const uint32_t not_m1 = ~w.has_m1;
  __syncthreads();

#pragma unroll
  for (int j = 0; j < PRE; ++j)
    if (pre[j].y) md_accumulate_static<PS>(acc, pre[j], tmin, not_m1);
  for (uint32_t i = tid + PRE * TILE_THREADS; i < count; i += TILE_THREADS) md_accumulate_static<PS>(acc, __ldg(rec + i), tmin, not_m1);
  __syncthreads();

  float o[PPT][C];
#pragma unroll
  for (int k = 0; k < PPT; ++k)
    md_finalise_static<PS>(acc + (tid + k * TILE_THREADS) * STRIDE, inv_delta, delta, delta_u, w.has_m1, o[k], std::make_integer_sequence<int, C>{});
  __syncthreads();
  float4* stage = reinterpret_cast<float4*>(acc);  // [TP][C] floats, contiguous = the global layout of the slice
#pragma unroll
  for (int k = 0; k < PPT; ++k)
#pragma unroll
    for (int q = 0; q < C / 4; ++q)
      stage[(tid + k * TILE_THREADS) * (C / 4) + q] = make_float4(o[k][4 * q], o[k][4 * q + 1], o[k][4 * q + 2], o[k][4 * q + 3]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  constexpr int SLICE_PX = TP / 4;  // four issuing threads, one per 1/4 of the tile
  if ((tid & 31) == 0 && tid < 128) {
    const int p0 = (tid >> 5) * SLICE_PX;
    const int np = min(SLICE_PX, npix - p0);
    if (np > 0) {
      float* dst = out + ((size_t)b * g.HW + pix0 + p0) * C;
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(acc + p0 * C);
      const uint32_t bytes = (uint32_t)np * C * 4u;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the read
    }
  }
}

template <typename PS>
static int launch_static(const Geom& g, const Workspace& ws, size_t smem, float* out, cudaStream_t stream) {
  constexpr int TP = 1024;  // what choose_tile picks for the ERGO-12 footprint on every sensor below 4 Mpx
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_md_tile_static<PS, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(EVREP_K_TILE, stream);
  k_md_tile_static<PS, TP><<<g.B * g.T, TILE_THREADS, smem, stream>>>(ws.records, ws.base, ws.cursor, ws.wp, g, out);
  prof_end(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

// SBT only: which time windows hold a p == -1 event (needed before the tile pass can decide what
// "negative" means in each window).  Runs after the binning pass, which produced t_min / t_max.
template <typename TT>
__global__ void k_sbt_negsel(const TT* __restrict__ t, const int8_t* __restrict__ p, WinParams* __restrict__ wp) {
  const int b = blockIdx.y;
  const WinParams w = wp[b];
  const double delta = (w.tmax_rel >= w.tmin_rel) ? (double)(uint32_t)(w.tmax_rel - w.tmin_rel) : 0.0;
  uint32_t m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < w.n; i += (int64_t)gridDim.x * blockDim.x) {
    if (p[w.start + i] != -1) continue;
    const int64_t d = (int64_t)t[w.start + i] - w.t_base;
    if (d >= T_REL_LIMIT || d <= -T_REL_LIMIT) continue;
    const double ts = (double)(uint32_t)((int32_t)d - w.tmin_rel) / delta;
    const double f = 1.0 / 3.0;
    m |= 1u;
    if (ts <= 1.0 * f && ts >= 0.0 * f) m |= 2u;
    if (ts <= 2.0 * f && ts >= 1.0 * f) m |= 4u;
    if (ts <= 3.0 * f && ts >= 2.0 * f) m |= 8u;
    if (ts <= 0.5) m |= 16u;
    if (ts <= 0.25) m |= 32u;
    if (ts <= 0.125) m |= 64u;
    if (ts <= 0.0625) m |= 128u;
  }
  m = __reduce_or_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicOr(&wp[b].has_m1, m);
}

int launch_md_tile(const Geom& g, const Workspace& ws, const MdPlan& plan, const Events& ev, float* out, cudaStream_t stream) {
  if (plan.stacking == EVREP_STACK_SBT) {
    dim3 grid(64, g.B);
    if (ev.t_bytes == 4)
      k_sbt_negsel<int32_t><<<grid, 256, 0, stream>>>((const int32_t*)ev.t, ev.p, ws.wp);
    else
      k_sbt_negsel<int64_t><<<grid, 256, 0, stream>>>((const int64_t*)ev.t, ev.p, ws.wp);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  const size_t smem = md_tile_smem_bytes(plan, g.tile_px);
  switch (g.tile_px == 1024 ? plan.static_id : 0) {
#define EVREP_STATIC_CASE(VER, LW) \
  case VER * 100 + LW:             \
    return launch_static<ErgoPlan<VER, LW>>(g, ws, smem, out, stream);
    EVREP_STATIC_CASE(2, 16) EVREP_STATIC_CASE(2, 14) EVREP_STATIC_CASE(2, 12) EVREP_STATIC_CASE(2, 10) EVREP_STATIC_CASE(2, 8)
    EVREP_STATIC_CASE(1, 16) EVREP_STATIC_CASE(1, 14) EVREP_STATIC_CASE(1, 12) EVREP_STATIC_CASE(1, 10) EVREP_STATIC_CASE(1, 8)
#undef EVREP_STATIC_CASE
    default: break;
  }
  auto kern = plan.C <= 12 ? k_md_tile<12> : k_md_tile<EVREP_MAX_CHANNELS>;
  EVREP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(EVREP_K_TILE, stream);
  kern<<<g.B * g.T, TILE_THREADS, smem, stream>>>(ws.records, ws.base, ws.cursor, ws.wp, plan, g, out);
  prof_end(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

}  // namespace evrep
