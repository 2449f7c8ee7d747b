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

#include "evrep_common.cuh"
#include "md_plan.cuh"
#include "md_tile_static.cuh"

#define EVREP_TRY_RC(expr)       \
  do {                          \
    int _rc = (expr);           \
    if (_rc != EVREP_OK) return _rc; \
  } while (0)

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
        if (md_plan_build(win, func, agg, C, stacking, lw, false, *out)) return EVREP_EUNSUPPORTED;
        out->static_id = ver * 100 + lw;
        return EVREP_OK;
      }
    }
  }
  if (md_plan_build(win, func, agg, C, stacking, lw, false, *out)) {
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
// One channel of one pixel.  Numerators and denominators are computed exactly (integers / fp64), then divided
// once in fp32 with the hardware reciprocal (<= 2 ulp): no fp64 division and no cancellation after rounding,
// within 4e-7 relative of the reference's fp64 result.  0 * rcp(0) = NaN reproduces the reference's 0/0.
__device__ __forceinline__ float md_value(const MdPlan& P, const MdChan& ch, const uint32_t* a, float inv_delta, double delta, uint32_t delta_u,
                                          uint32_t has_m1) {
  if (!ch.valid) return 0.f;
  if (ch.func == EVREP_FUNC_POLARITY) {
    const int c1 = (int)md_cnt(P, P.grp[ch.g_pos], a);
    const int cn = (int)md_cnt(P, P.grp[ch.g_neg], a);
    const int cm = ((has_m1 >> ch.win) & 1u) ? cn : 0;  // the "negative" class holds the p == 0 events when the window has no -1
    if (ch.agg == EVREP_AGG_SUM) return (float)(c1 - cm);
    const int call = c1 + cn + (int)md_cnt(P, P.grp[ch.g_oth], a);
    if (call == 0) return 0.f;
    if (ch.agg == EVREP_AGG_MEAN) return __fdividef((float)(c1 - cm), (float)call);
    if (ch.agg == EVREP_AGG_VARIANCE) {  // mean(p^2) - mean(p)^2 = ((c1+cm) call - (c1-cm)^2) / call^2, exact in 64-bit integers
      const long long dd = (long long)(c1 - cm);
      const unsigned long long num = (unsigned long long)(uint32_t)(c1 + cm) * (uint32_t)call - (unsigned long long)(dd * dd);
      return __fdividef(__ull2float_rn(num), __ull2float_rn((unsigned long long)(uint32_t)call * (uint32_t)call));
    }
    if (ch.agg == EVREP_AGG_MIN) return cm > 0 ? -1.f : (call - c1 - cm > 0 ? 0.f : 1.f);  // min of the raw polarities
    return c1 > 0 ? 1.f : (call - c1 - cm > 0 ? 0.f : -1.f);  // max of the raw polarities
  }
  const bool is_count = (ch.func == EVREP_FUNC_COUNT || ch.func == EVREP_FUNC_COUNT_POS || ch.func == EVREP_FUNC_COUNT_NEG);
  if (ch.g_main < 0 && ch.g_pos < 0) return 0.f;  // variance of a constant
  const uint32_t c = md_count(P, ch, a);
  if (c == 0u) return 0.f;  // torch_scatter leaves untouched pixels at 0
  if (is_count) return ch.agg == EVREP_AGG_SUM ? (float)c : 1.f;
  // timestamps: t_s = (t - t_min) / (t_max - t_min); delta == 0 gives NaN exactly like the reference
  const MdGroup& G = P.grp[ch.g_main];
  if (ch.agg == EVREP_AGG_MAX) {
    const uint32_t v = a[G.w_max] - 1u;
    return (v == delta_u && delta_u) ? 1.f : (float)v * inv_delta;  // the window's last event maps to exactly 1
  }
  if (ch.agg == EVREP_AGG_MIN) {
    const uint32_t v = ~a[G.w_min];
    return (v == delta_u && delta_u) ? 1.f : (float)v * inv_delta;
  }
  // sum of t: below 2^63 (fewer than 2^32 events of t < 2^31), exact in 64-bit integers
  unsigned long long sti = a[G.w_st + P.nl1 - 1];
  for (int l = P.nl1 - 2; l >= 0; --l) sti = (sti << P.lw) + a[G.w_st + l];
  if (ch.agg == EVREP_AGG_SUM) return __ull2float_rn(sti) * inv_delta;
  if (ch.agg == EVREP_AGG_MEAN) return __fdividef(__ull2float_rn(sti), __ull2float_rn((unsigned long long)c * delta_u));
  if (c == 1u && delta_u) return 0.f;  // a single event: t_s^2 - t_s^2, exactly 0 in the reference too (NaN when delta == 0)
  const double st = (double)sti;
  const double cd = (double)c * delta;
  const double st2 = md_limb_sum(a, G.w_st2, P.nl2, P.lw);
  // mean(t_s^2) - mean(t_s)^2 = (c sum(t^2) - sum(t)^2) / (c delta)^2
  return __fdividef((float)fma((double)c, st2, -st * st), (float)(cd * cd));
}

template <int CMAX>
__global__ void __launch_bounds__(TILE_THREADS) k_md_tile(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                          const uint32_t* __restrict__ hist, const WinParams* __restrict__ wp,
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
  const uint32_t count = hist[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];
  const int32_t tmin = w.tmin_rel;
  const uint32_t delta_u = (w.tmax_rel >= w.tmin_rel) ? (uint32_t)(w.tmax_rel - w.tmin_rel) : 0u;
  const double delta = (double)delta_u;
  const float inv_delta = 1.f / (float)delta_u;
  const uint32_t limb_mask = (1u << P.lw) - 1u;  // lw <= 31
  __syncthreads();

  for (uint32_t i = tid; i < count; i += TILE_THREADS) {
    const uint2 r = __ldg(rec + i);
    if (rec_is_null(r.y)) continue;
    uint32_t* a = acc + (r.y & 0xffffu) * stride;
    const uint32_t pc = (r.y >> 24) & 3u;
    const uint32_t tt = (uint32_t)((int32_t)r.x - tmin);  // t - t_min, < 2^31
    uint32_t wmask;
    if (P.stacking == EVREP_STACK_SBN) {
      wmask = (r.y >> 16) & 0xffu;
    } else {
      wmask = md_sbt_wmask(tt, delta);
    }
    // "negative" events of a window: p == -1, or p == 0 when the window holds no -1 (operations.py:59-61,78-80)
    const uint32_t posm = (pc == 1u) ? wmask : 0u;
    const uint32_t negm = (pc == 3u) ? wmask : (pc == 0u ? (wmask & ~w.has_m1) : 0u);
    const uint32_t M = wmask | (posm << 8) | (negm << 16) | ((wmask & ~(posm | negm)) << 24);
    uint32_t pres = 0;
    for (int gi = 0; gi < P.G; ++gi) {
      const MdGroup G = P.grp[gi];
      if (!((M >> G.bit) & 1u)) continue;
      if (G.flags & G_CNT) atomicAdd(a + G.w_cnt, 1u << G.cnt_shift);
      if (G.flags & G_PRES) pres |= 1u << G.pres_bit;
      if (G.flags & G_MAX) atomicMax(a + G.w_max, tt + 1u);
      if (G.flags & G_MIN) atomicMax(a + G.w_min, ~tt);  // tt < 2^31: never 0
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

static int sm_count(int* n_sm) {  // of the calling thread's current device (cached per device ordinal)
  constexpr int MAX_DEV = 64;
  static int cached[MAX_DEV] = {};
  int dev = 0;
  EVREP_CUDA_OK(cudaGetDevice(&dev));
  int n = (dev >= 0 && dev < MAX_DEV) ? cached[dev] : 0;
  if (!n) {
    EVREP_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < MAX_DEV) cached[dev] = n;  // benign race: every writer stores the same value
  }
  *n_sm = n;
  return EVREP_OK;
}

static_assert(ErgoPlan<2, 16, true>::value.words == 17 && ErgoPlan<2, 16, true>::value.stride == 17, "packed ERGO-12 v2 plan: 17 words per pixel");
static_assert(ErgoPlan<2, 12, false>::value.words == 24, "wide ERGO-12 v2 plan at 2^20 events: 24 words per pixel");

// ERGO-12 at the standard 1024-pixel tile: packed plan on every bucket below 65536 events, wide plan on the rest
template <int VER, int LW, bool SPLIT>
static int launch_static(const Geom& g, const Workspace& ws, float* out, cudaStream_t stream) {
  constexpr int TP = 1024;
  using Packed = ErgoPlan<VER, 16, true>;
  using Wide = ErgoPlan<VER, LW, false>;
  int n_sm = 0;
  EVREP_TRY_RC(sm_count(&n_sm));
  const int n_tiles = g.B * g.T;
  auto light = k_md_tile_static<Packed, TP, true, SPLIT>;
  auto heavy = k_md_tile_heavy<Wide, TP, SPLIT>;
  const size_t smem_l = align_up((size_t)Packed::value.stride * TP * 4, 128), smem_h = align_up((size_t)Wide::value.stride * TP * 4, 128);
  EVREP_CUDA_OK(cudaFuncSetAttribute(light, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
  EVREP_CUDA_OK(cudaFuncSetAttribute(heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h));
  int per_sm = 1;
  EVREP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, light, TILE_THREADS, smem_l));
  if (per_sm < 1) per_sm = 1;
  const int grid = n_tiles < per_sm * n_sm ? n_tiles : per_sm * n_sm;
  prof_begin(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(launch_pdl(light, grid, TILE_THREADS, smem_l, stream, ws.records, ws.base, ws.hist, ws.wp, g, ws.ticket, out));
  if (!(g.n_max > 0 && g.n_max < (int64_t)MD_PACKED_LIMIT))  // a bucket cannot hold more events than its window: nothing for the wide plan
    EVREP_CUDA_OK(launch_pdl(heavy, (n_tiles + TILE_THREADS - 1) / TILE_THREADS < n_sm ? (n_tiles + TILE_THREADS - 1) / TILE_THREADS : n_sm, TILE_THREADS, smem_h,
                             stream, ws.records, ws.base, ws.hist, ws.wp, g, out));
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
    m |= md_sbt_wmask((uint32_t)((int32_t)d - w.tmin_rel), delta);
  }
  m = __reduce_or_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicOr(&wp[b].has_m1, m);
}

int launch_sbt_negsel(const Geom& g, const Workspace& ws, const Events& ev, cudaStream_t stream) {
  dim3 grid(64, g.B);
  if (ev.t_bytes == 4)
    k_sbt_negsel<int32_t><<<grid, 256, 0, stream>>>((const int32_t*)ev.t, ev.p, ws.wp);
  else
    k_sbt_negsel<int64_t><<<grid, 256, 0, stream>>>((const int64_t*)ev.t, ev.p, ws.wp);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

int launch_md_tile(const Geom& g, const Workspace& ws, const MdPlan& plan, const Events& ev, float* out, cudaStream_t stream) {
  if (plan.stacking == EVREP_STACK_SBT) EVREP_TRY_RC(launch_sbt_negsel(g, ws, ev, stream));
  const size_t smem = md_tile_smem_bytes(plan, g.tile_px);
  switch (g.tile_px == 1024 ? plan.static_id : 0) {
#define EVREP_STATIC_CASE(VER, LW) \
  case VER * 100 + LW:             \
    return g.split ? launch_static<VER, LW, true>(g, ws, out, stream) : launch_static<VER, LW, false>(g, ws, out, stream);
    EVREP_STATIC_CASE(2, 16) EVREP_STATIC_CASE(2, 14) EVREP_STATIC_CASE(2, 12) EVREP_STATIC_CASE(2, 10) EVREP_STATIC_CASE(2, 8)
    EVREP_STATIC_CASE(1, 16) EVREP_STATIC_CASE(1, 14) EVREP_STATIC_CASE(1, 12) EVREP_STATIC_CASE(1, 10) EVREP_STATIC_CASE(1, 8)
#undef EVREP_STATIC_CASE
    default: break;
  }
  auto kern = plan.C <= 12 ? k_md_tile<12> : k_md_tile<EVREP_MAX_CHANNELS>;
  EVREP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(EVREP_K_TILE, stream);
  kern<<<g.B * g.T, TILE_THREADS, smem, stream>>>(ws.records, ws.base, ws.hist, ws.wp, plan, g, out);
  prof_end(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

}  // namespace evrep
