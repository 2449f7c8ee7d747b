// The three "latest event wins" representations on top of the tile buckets of binning.cu:
//   EventStack   representations/event_stack.py:15-131 as called at gen1_transforms.py:33-42
//   TimeSurface  representations/time_surface.py:25-74 as called at gen1_transforms.py:69-87
//   TORE         representations/tore.py:6-83 as called at gen1_transforms.py:51-67
// The order-dependent stores of the reference (np.put, timestamp_memory[...] = t, the per-pixel FIFO)
// become order-independent integer atomicMax operations on keys that grow with stream order, so the
// result does not depend on which thread gets there first.
#include <math.h>

#include "evrep_common.cuh"

namespace evrep {

__device__ __forceinline__ void zero_smem(uint32_t* acc, int n_words) {
  uint4* a4 = reinterpret_cast<uint4*>(acc);
  for (int i = threadIdx.x; i < n_words / 4; i += TILE_THREADS) a4[i] = make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------------------------
// EventStack: out[y, x, k] = polarity sign of the latest event at the pixel if its index >= s_k else 0,
// s_k = start of the k-th nested suffix window (event_stack.py:70-82: c //= 2; x = x[c:]).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS) k_event_stack_tile(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                                   const uint32_t* __restrict__ hist, const WinParams* __restrict__ wp,
                                                                   const Geom g, int K, float* __restrict__ out) {
  extern __shared__ __align__(16) uint32_t acc[];
  __shared__ uint32_t s_start[EVREP_MAX_CHANNELS];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int TP = g.tile_px, pix0 = tile << g.tile_shift, npix = min(TP, g.HW - pix0);
  zero_smem(acc, TP);
  const WinParams w = wp[b];
  if (tid == 0) {
    int64_t c = w.n, s = 0;
    for (int k = 0; k < K; ++k) {
      s_start[k] = (uint32_t)min(s, w.n);
      c /= 2;
      s += c;
    }
  }
  const uint32_t count = hist[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];
  __syncthreads();
  for (uint32_t i = tid; i < count; i += TILE_THREADS) {
    const uint2 r = __ldg(rec + i);
    if (rec_is_null(r.y)) continue;
    const uint32_t pol = (((r.y >> 24) & 3u) == 1u) ? 1u : 0u;  // p > 0
    atomicMax(&acc[r.y & 0xffffu], ((r.x + 1u) << 1) | pol);
  }
  __syncthreads();
  float* dst = out + ((size_t)b * g.HW + pix0) * K;
  const int n_el = npix * K;
  for (int e = tid; e < n_el; e += TILE_THREADS) {
    const int pix = e / K, k = e - pix * K;
    const uint32_t v = acc[pix];
    float o = 0.f;
    if (v && ((v >> 1) - 1u) >= s_start[k]) o = (v & 1u) ? 1.f : -1.f;
    dst[e] = o;
  }
}

// Compile-time K (the reference's stack_size = 12): one thread per pixel computes all K channels, the K floats of 512
// pixels are staged in shared memory ([pixel][K], 16-byte stores at a 48-byte stride are conflict free) and leave as
// consecutive float4: no integer division, no per-element address arithmetic (ncu r01: the generic kernel spent 38
// instructions per output element and was issue bound at 42 % of the DRAM peak).
template <int K, bool FUSED>
__global__ void __launch_bounds__(TILE_THREADS, 4) k_event_stack_tile_k(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                                     const uint32_t* __restrict__ hist, const WinParams* __restrict__ wp,
                                                                     const Geom g, float* __restrict__ out) {
  static_assert(K % 4 == 0, "float4 staging");
  extern __shared__ __align__(16) uint32_t acc[];  // TP accumulators, then TILE_THREADS * K staged floats
  __shared__ uint32_t s_start[K];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int TP = g.tile_px, pix0 = tile << g.tile_shift, npix = min(TP, g.HW - pix0);
  float4* stage = reinterpret_cast<float4*>(acc + TP);
  const WinParams w = wp[b];
  const uint32_t count = hist[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];
  // the thread's first records are requested before the accumulators are cleared: the header -> record dependent chain of
  // global loads (ncu: 11 warps per issue slot waiting on it) then runs under the zeroing and the barrier
  constexpr int PRE = 2;
  uint2 pre[PRE];
#pragma unroll
  for (int j = 0; j < PRE; ++j) pre[j] = (uint32_t)(tid + j * TILE_THREADS) < count ? __ldg(rec + tid + j * TILE_THREADS) : make_uint2(0u, 0u);
  zero_smem(acc, TP);
  if (tid == 0) {
    int64_t c = w.n, st = 0;
    for (int k = 0; k < K; ++k) {
      s_start[k] = (uint32_t)min(st, w.n);
      c /= 2;
      st += c;
    }
  }
  auto feed = [&](const uint2 r) {
    if (FUSED) {  // REC_T_IDX records: the stream index travels in the meta word
      if (fused_pc(r.y) == 2u) return;
      atomicMax(&acc[fused_pix(r.y)], ((fused_idx(r.y) + 1u) << 1) | (fused_pc(r.y) == 1u ? 1u : 0u));
      return;
    }
    if (rec_is_null(r.y)) return;
    const uint32_t pol = (((r.y >> 24) & 3u) == 1u) ? 1u : 0u;  // p > 0
    atomicMax(&acc[r.y & 0xffffu], ((r.x + 1u) << 1) | pol);
  };
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PRE; ++j)
    if ((uint32_t)(tid + j * TILE_THREADS) < count) feed(pre[j]);
  for (uint32_t i = tid + PRE * TILE_THREADS; i < count; i += TILE_THREADS) feed(__ldg(rec + i));
  __syncthreads();
  uint32_t start[K];
#pragma unroll
  for (int k = 0; k < K; ++k) start[k] = s_start[k];
  float4* dst4 = reinterpret_cast<float4*>(out + ((size_t)b * g.HW + pix0) * K);
  for (int p0 = 0; p0 < npix; p0 += TILE_THREADS) {
    const uint32_t v = (p0 + tid < npix) ? acc[p0 + tid] : 0u;
    const float sgn = (v & 1u) ? 1.f : -1.f;
    const uint32_t idx = (v >> 1) - 1u;  // v == 0: 0xffffffff, masked by the v test
    float o[K];
#pragma unroll
    for (int k = 0; k < K; ++k) o[k] = (v && idx >= start[k]) ? sgn : 0.f;
    // a warp repacks its own 32 pixels ([pixel][K] -> consecutive float4) in its own slice of the staging area: no CTA barrier
    float4* wst = stage + (tid & ~31) * (K / 4);
    const int lane = tid & 31;
#pragma unroll
    for (int q = 0; q < K / 4; ++q) wst[lane * (K / 4) + q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    __syncwarp();
    const int n4 = min(32, npix - p0 - (tid & ~31)) * (K / 4);  // may be <= 0 in the last tile of the image
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
      const int e = q * 32 + lane;
      if (e < n4) __stcs(dst4 + (size_t)(p0 + (tid & ~31)) * (K / 4) + e, wst[e]);
    }
    __syncwarp();
  }
}

int launch_event_stack_tile(const Geom& g, const Workspace& ws, int stack_size, float* out, cudaStream_t stream) {
  if (stack_size == 12 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    const size_t smem12 = sizeof(uint32_t) * (size_t)g.tile_px + sizeof(float) * 12 * TILE_THREADS;
    EVREP_CUDA_OK(cudaFuncSetAttribute(k_event_stack_tile_k<12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem12));
    prof_begin(EVREP_K_TILE, stream);
    k_event_stack_tile_k<12, false><<<g.B * g.T, TILE_THREADS, smem12, stream>>>(ws.records, ws.base, ws.hist, ws.wp, g, out);
    prof_end(EVREP_K_TILE, stream);
    EVREP_CUDA_OK(cudaGetLastError());
    return EVREP_OK;
  }
  const size_t smem = sizeof(uint32_t) * (size_t)g.tile_px;
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_event_stack_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(EVREP_K_TILE, stream);
  k_event_stack_tile<<<g.B * g.T, TILE_THREADS, smem, stream>>>(ws.records, ws.base, ws.hist, ws.wp, g, stack_size, out);
  prof_end(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

// ---------------------------------------------------------------------------------------------
// TimeSurface: surface s = exp((mem - t_s) / tau), mem = timestamp of the latest event at
// (polarity, y, x) with index <= indices[s], or -(3 tau + 1) where there is none.
// Shared memory holds, per (first snapshot fed, polarity, pixel), the latest timestamp; a running
// maximum over the snapshot axis reproduces the sequential memory of time_surface.py:66-74.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS) k_time_surface_tile(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                                    const uint32_t* __restrict__ hist, const WinParams* __restrict__ wp,
                                                                    const SnapParams* __restrict__ snap, const Geom g, int S, double tau,
                                                                    float* __restrict__ out) {
  extern __shared__ __align__(16) uint32_t acc[];  // [S][2][TP]
  __shared__ int32_t s_trel[MAX_SNAP];
  __shared__ float s_empty[MAX_SNAP];
  __shared__ int s_nvalid;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int TP = g.tile_px, pix0 = tile << g.tile_shift, npix = min(TP, g.HW - pix0);
  zero_smem(acc, S * 2 * TP);
  const WinParams w = wp[b];
  if (tid < MAX_SNAP) {
    const int32_t tr = snap[b].t_rel[tid];
    s_trel[tid] = tr;
    // untouched pixels: exp((-(3 tau + 1) - t_snapshot) / tau) with the ABSOLUTE snapshot timestamp
    s_empty[tid] = (float)exp((-(tau * 3.0 + 1.0) - (double)(w.t_base + (int64_t)tr)) / tau);
  }
  if (tid == 0) s_nvalid = snap[b].n_valid;
  const uint32_t count = hist[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];
  const int32_t tmin = w.tmin_rel;
  __syncthreads();
  for (uint32_t i = tid; i < count; i += TILE_THREADS) {
    const uint2 r = __ldg(rec + i);
    if (rec_is_null(r.y)) continue;
    const uint32_t s = (r.y >> 16) & 0xffu;
    const uint32_t plane = (((r.y >> 24) & 3u) == 1u) ? 1u : 0u;
    atomicMax(&acc[(s * 2u + plane) * TP + (r.y & 0xffffu)], (uint32_t)((int32_t)r.x - tmin) + 1u);
  }
  __syncthreads();
  const double inv_tau = 1.0 / tau;
  const int nvalid = s_nvalid;
  for (int e = tid; e < 2 * npix; e += TILE_THREADS) {
    const int plane = e / npix, pix = e - plane * npix;
    uint32_t m = 0;
    for (int s = 0; s < S; ++s) {
      m = max(m, acc[(s * 2 + plane) * TP + pix]);
      float o = 0.f;
      if (s < nvalid) {
        if (m) {
          const int64_t mem_rel = (int64_t)(m - 1u) + (int64_t)tmin;
          o = expf((float)((double)(mem_rel - (int64_t)s_trel[s]) * inv_tau));
        } else {
          o = s_empty[s];
        }
      }
      out[(((size_t)b * S + s) * 2 + plane) * g.HW + pix0 + pix] = o;
    }
  }
}

// Compile-time S (the reference's 6 snapshots): the snapshot loop is unrolled, the running maximum lives in a register
// and the S stores of a thread go to one base pointer plus constant strides (the generic kernel: 62 instructions per
// output element, issue bound at 23 % of the DRAM peak).
template <int S, bool FUSED>
__global__ void __launch_bounds__(TILE_THREADS, 4) k_time_surface_tile_s(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                                      const uint32_t* __restrict__ hist, const WinParams* __restrict__ wp,
                                                                      const SnapParams* __restrict__ snap, const Geom g, double tau,
                                                                      float* __restrict__ out) {
  extern __shared__ __align__(16) uint32_t acc[];  // [S][2][TP]
  __shared__ int32_t s_trel[S];
  __shared__ int32_t s_sidx[S];
  __shared__ float s_empty[S];
  __shared__ int s_nvalid;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int TP = g.tile_px, pix0 = tile << g.tile_shift, npix = min(TP, g.HW - pix0);
  const WinParams w = wp[b];
  const uint32_t count = hist[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];
  constexpr int PRE = 2;  // first records requested before the accumulators are cleared (see k_event_stack_tile_k)
  uint2 pre[PRE];
#pragma unroll
  for (int j = 0; j < PRE; ++j) pre[j] = (uint32_t)(tid + j * TILE_THREADS) < count ? __ldg(rec + tid + j * TILE_THREADS) : make_uint2(0u, 0u);
  zero_smem(acc, S * 2 * TP);
  if (tid < S) {
    const int32_t tr = snap[b].t_rel[tid];
    s_trel[tid] = tr;
    s_sidx[tid] = snap[b].idx[tid];
  }
  if (tid == 0) s_nvalid = snap[b].n_valid;
  const int32_t tmin = w.tmin_rel;
  __syncthreads();
  // untouched pixels: exp((-(3 tau + 1) - t_snapshot) / tau) with the ABSOLUTE snapshot timestamp.  A double-precision exp
  // (~100 dependent instructions) that only the finalise needs: evaluated by S threads while the CTA accumulates instead of
  // ahead of the barrier every warp waits at
  if (tid < S) s_empty[tid] = (float)exp((-(tau * 3.0 + 1.0) - (double)(w.t_base + (int64_t)s_trel[tid])) / tau);
  auto feed = [&](const uint2 r) {
    if (FUSED) {  // REC_T_IDX records: the first snapshot an event feeds is found from its stream index here
      if (fused_pc(r.y) == 2u) return;
      const int idx = (int)fused_idx(r.y), ns = s_nvalid;
      int sn = 0;
      while (sn < ns && idx > s_sidx[sn]) ++sn;
      if (sn >= ns) return;  // after the last emitted surface
      atomicMax(&acc[((uint32_t)sn * 2u + (fused_pc(r.y) == 1u ? 1u : 0u)) * TP + fused_pix(r.y)], (uint32_t)((int32_t)r.x - tmin) + 1u);
      return;
    }
    if (rec_is_null(r.y)) return;
    const uint32_t sn = (r.y >> 16) & 0xffu;
    const uint32_t plane = (((r.y >> 24) & 3u) == 1u) ? 1u : 0u;
    atomicMax(&acc[(sn * 2u + plane) * TP + (r.y & 0xffffu)], (uint32_t)((int32_t)r.x - tmin) + 1u);
  };
#pragma unroll
  for (int j = 0; j < PRE; ++j)
    if ((uint32_t)(tid + j * TILE_THREADS) < count) feed(pre[j]);
  for (uint32_t i = tid + PRE * TILE_THREADS; i < count; i += TILE_THREADS) feed(__ldg(rec + i));
  __syncthreads();
  const double inv_tau_log2e = 1.4426950408889634 / tau;
  const int nvalid = s_nvalid;
  int32_t trel[S];
  float empty[S];
#pragma unroll
  for (int k = 0; k < S; ++k) {
    trel[k] = s_trel[k] - tmin + 1;  // compare against the stored key (t - tmin + 1)
    empty[k] = k < nvalid ? s_empty[k] : 0.f;
  }
  const size_t plane_stride = (size_t)g.HW, snap_stride = 2 * (size_t)g.HW;
  // (mem - t_snapshot) / tau with mem = key - 1 + tmin, t_snapshot = trel - 1 + tmin: the difference d of the keys, and
  // exp(d / tau) = 2^(d log2(e) / tau).  Keys lie in [1, t_max - t_min + 1]: for windows shorter than 2^24 us (16.7 s) d is
  // exact in float32 and the product with the constant, split into a float32 head and tail, carries one rounding
  // (<= 6e-8 |argument|, i.e. 3e-6 relative in the result at |argument| = 80 where the result is 1e-24); ex2.approx is good
  // to 2^-22.  No conversion or double-precision instruction per output: the int -> double -> float path kept the XU pipe
  // 51 % busy and the kernel at 23 instructions per output (ncu, profiles/README.md).  Longer windows take that path.
  const uint32_t delta_u = (w.tmax_rel >= w.tmin_rel) ? (uint32_t)(w.tmax_rel - w.tmin_rel) : 0u;
  const bool small = delta_u < (1u << 24) - 1u;  // CTA-uniform
  const float c_hi = (float)inv_tau_log2e, c_lo = (float)(inv_tau_log2e - (double)c_hi);
  if (small && (npix & 3) == 0 && (g.HW & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    // four consecutive pixels of one polarity plane per thread: one 16-byte shared load and one 16-byte streaming store
    // per snapshot (the scalar version spent ~18 instructions per output, most of them addressing and loop control)
    const int nq = npix >> 2;
    for (int item = tid; item < 2 * nq; item += TILE_THREADS) {
      const int plane = item >= nq ? 1 : 0, q = item - plane * nq;
      float* dst = out + ((size_t)b * S * 2 + plane) * plane_stride + pix0 + 4 * q;
      uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const uint4 a = *reinterpret_cast<const uint4*>(&acc[(k * 2 + plane) * TP + 4 * q]);
        m.x = max(m.x, a.x), m.y = max(m.y, a.y), m.z = max(m.z, a.z), m.w = max(m.w, a.w);
        const float em = empty[k];
        auto val = [&](uint32_t mm) {
          const float df = (float)((int32_t)mm - trel[k]);
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(df, c_lo, df * c_hi)));
          return mm ? e : em;
        };
        float4 o = make_float4(em, em, em, em);  // k >= nvalid: zeros
        if (k < nvalid) o = make_float4(val(m.x), val(m.y), val(m.z), val(m.w));
        __stcs(reinterpret_cast<float4*>(dst + k * snap_stride), o);
      }
    }
    return;
  }
  for (int plane = 0; plane < 2; ++plane) {
    float* dst = out + ((size_t)b * S * 2 + plane) * plane_stride + pix0;
    if (small) {
      for (int pix = tid; pix < npix; pix += TILE_THREADS) {
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < S; ++k) {
          m = max(m, acc[(k * 2 + plane) * TP + pix]);
          const float df = (float)((int32_t)m - trel[k]);
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(df, c_lo, df * c_hi)));
          __stcs(dst + k * snap_stride + pix, (m && k < nvalid) ? e : empty[k]);
        }
      }
    } else {
      for (int pix = tid; pix < npix; pix += TILE_THREADS) {
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < S; ++k) {
          m = max(m, acc[(k * 2 + plane) * TP + pix]);
          float o = empty[k];
          if (m && k < nvalid) o = exp2f((float)((double)((int32_t)m - trel[k]) * inv_tau_log2e));
          __stcs(dst + k * snap_stride + pix, o);
        }
      }
    }
  }
}

int launch_time_surface_tile(const Geom& g, const Workspace& ws, int S, double tau, float* out, cudaStream_t stream) {
  if (S == 6) {
    const size_t smem6 = sizeof(uint32_t) * (size_t)g.tile_px * 2 * 6;
    EVREP_CUDA_OK(cudaFuncSetAttribute(k_time_surface_tile_s<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6));
    prof_begin(EVREP_K_TILE, stream);
    k_time_surface_tile_s<6, false><<<g.B * g.T, TILE_THREADS, smem6, stream>>>(ws.records, ws.base, ws.hist, ws.wp, ws.snap, g, tau, out);
    prof_end(EVREP_K_TILE, stream);
    EVREP_CUDA_OK(cudaGetLastError());
    return EVREP_OK;
  }
  const size_t smem = sizeof(uint32_t) * (size_t)g.tile_px * 2 * S;
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_time_surface_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(EVREP_K_TILE, stream);
  k_time_surface_tile<<<g.B * g.T, TILE_THREADS, smem, stream>>>(ws.records, ws.base, ws.hist, ws.wp, ws.snap, g, S, tau, out);
  prof_end(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

// float32 constants of tore.py:69-79 as literals (nvcc does not fold log() of a constant: every thread spent ~60 instructions,
// most of them double precision, on them - ncu r02): log(minTime + 1) and the value of a slot that never saw an event,
// log(float32(maxTime) + 1) - log(151) in float32 as numpy evaluates it
#define TORE_LOG151 __uint_as_float(0x40a08d8eu)  /* 5.0172796 */
#define TORE_EMPTY __uint_as_float(0x41703497u)   /* 15.012839 */

// ---------------------------------------------------------------------------------------------
// TORE: per pixel and polarity class the k most recent ages T - t among events with t < T
// (T = timestamp of the window's last event), ascending, then the float32 log compression of
// tore.py:69-79.  The FIFO of the reference becomes a cascade of atomicMax: slot j receives what
// slot j-1 displaced, which leaves the k largest timestamps sorted whatever the arrival order,
// duplicates included.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS) k_tore_tile(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                            const uint32_t* __restrict__ hist, const WinParams* __restrict__ wp,
                                                            const Geom g, int K, float* __restrict__ out) {
  extern __shared__ __align__(16) uint32_t acc[];  // [2][K][TP]
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int TP = g.tile_px, pix0 = tile << g.tile_shift, npix = min(TP, g.HW - pix0);
  zero_smem(acc, 2 * K * TP);
  const WinParams w = wp[b];
  const uint32_t count = hist[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];
  const int32_t tmin = w.tmin_rel;
  __syncthreads();
  for (uint32_t i = tid; i < count; i += TILE_THREADS) {
    const uint2 r = __ldg(rec + i);
    if (rec_is_null(r.y)) continue;
    const uint32_t plane = (((r.y >> 24) & 3u) == 1u) ? 0u : 1u;  // positive first (tore.py:63-65)
    uint32_t v = (uint32_t)((int32_t)r.x - tmin) + 1u;
    uint32_t* slot = &acc[plane * K * TP + (r.y & 0xffffu)];
    for (int j = 0; j < K && v; ++j) {
      const uint32_t old = atomicMax(slot + j * TP, v);
      v = min(old, v);
    }
  }
  __syncthreads();
  const float max_time = 500e6f;
  const float log151 = TORE_LOG151;
  const float empty = TORE_EMPTY;
  float* dst = out + ((size_t)b * g.HW + pix0) * (2 * K);
  const int n_el = npix * 2 * K;
  for (int e = tid; e < n_el; e += TILE_THREADS) {
    const int pix = e / (2 * K), c = e - pix * 2 * K;
    const uint32_t v = acc[c * TP + pix];
    float o = empty;
    if (v) {
      const int64_t t_rel = (int64_t)(v - 1u) + (int64_t)tmin;
      float age = (float)((int64_t)w.tlast_rel - t_rel);
      age = fminf(age, max_time);
      o = fmaxf(logf(age + 1.f) - log151, 0.f);
    }
    dst[e] = o;
  }
}

// Compile-time K (the reference's k = 6): one thread per pixel, 2K outputs staged as [pixel][2K] and copied out as
// consecutive float4; logf only runs for warps that hold a filled slot (slots beyond the first are nearly always empty).
template <int K, bool FUSED>
__global__ void __launch_bounds__(TILE_THREADS) k_tore_tile_k(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                              const uint32_t* __restrict__ hist, const WinParams* __restrict__ wp,
                                                              const Geom g, float* __restrict__ out) {
  constexpr int C = 2 * K;
  static_assert(C % 4 == 0, "float4 staging");
  extern __shared__ __align__(16) uint32_t acc[];  // [2][K][TP], then TILE_THREADS * C staged floats
  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int TP = g.tile_px, pix0 = tile << g.tile_shift, npix = min(TP, g.HW - pix0);
  float4* stage = reinterpret_cast<float4*>(acc + C * TP);
  const WinParams w = wp[b];
  const uint32_t count = hist[blockIdx.x];
  const uint2* rec = records + w.start + base[blockIdx.x];
  constexpr int PRE = 2;  // first records requested before the accumulators are cleared (see k_event_stack_tile_k)
  uint2 pre[PRE];
#pragma unroll
  for (int j = 0; j < PRE; ++j) pre[j] = (uint32_t)(tid + j * TILE_THREADS) < count ? __ldg(rec + tid + j * TILE_THREADS) : make_uint2(0u, 0u);
  zero_smem(acc, C * TP);
  const int32_t tmin = w.tmin_rel;
  auto feed = [&](const uint2 r) {
    uint32_t plane, pix;
    if (FUSED) {  // REC_T_IDX records: the strict `t < sample time` cut of tore.py:17 is applied here
      if (fused_pc(r.y) == 2u || (int32_t)r.x >= w.tlast_rel) return;
      plane = fused_pc(r.y) == 1u ? 0u : 1u;
      pix = fused_pix(r.y);
    } else {
      if (rec_is_null(r.y)) return;
      plane = (((r.y >> 24) & 3u) == 1u) ? 0u : 1u;  // positive first (tore.py:63-65)
      pix = r.y & 0xffffu;
    }
    uint32_t v = (uint32_t)((int32_t)r.x - tmin) + 1u;
    uint32_t* slot = &acc[plane * K * TP + pix];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (!v) break;
      const uint32_t old = atomicMax(slot + j * TP, v);
      v = min(old, v);
    }
  };
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PRE; ++j)
    if ((uint32_t)(tid + j * TILE_THREADS) < count) feed(pre[j]);
  for (uint32_t i = tid + PRE * TILE_THREADS; i < count; i += TILE_THREADS) feed(__ldg(rec + i));
  __syncthreads();
  const float max_time = 500e6f;
  const float log151 = TORE_LOG151;
  const float empty = TORE_EMPTY;
  const int32_t age0 = w.tlast_rel - tmin + 1;  // age = tlast_rel - (key - 1 + tmin)
  float4* dst4 = reinterpret_cast<float4*>(out + ((size_t)b * g.HW + pix0) * C);
  for (int p0 = 0; p0 < npix; p0 += TILE_THREADS) {
    const bool live = p0 + tid < npix;
    float o[C];
#pragma unroll
    for (int plane = 0; plane < 2; ++plane) {
      // the cascade keeps the k most recent timestamps sorted: slot j is filled only if slot j - 1 is, so a warp stops
      // loading a polarity's slots at the first one that is empty in all of its 32 pixels (slots past the second nearly
      // always are)
      bool alive = true;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int c = plane * K + j;
        o[c] = empty;
        if (alive) {
          const uint32_t v = live ? acc[c * TP + p0 + tid] : 0u;
          alive = __any_sync(0xffffffffu, v != 0u);
          // log(age + 1) through lg2.approx (|error| <= 2^-22 in log2, 1.7e-7 here) and one fused multiply-add: inside the
          // 1e-6 absolute floor the float32 reference itself needs around age = 150; logf cost ~25 instructions per slot
          const float age = fminf((float)(age0 - (int32_t)v), max_time);
          float l2;
          asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(age + 1.f));  // argument >= 1: no denormal handling needed
          const float lg = fmaxf(fmaf(l2, 0.693147180559945f, -log151), 0.f);
          o[c] = v ? lg : empty;
        }
      }
    }
    // a warp repacks its own 32 pixels in its own slice of the staging area: no CTA barrier
    float4* wst = stage + (tid & ~31) * (C / 4);
    const int lane = tid & 31;
#pragma unroll
    for (int q = 0; q < C / 4; ++q) wst[lane * (C / 4) + q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    __syncwarp();
    const int n4 = min(32, npix - p0 - (tid & ~31)) * (C / 4);  // may be <= 0 in the last tile of the image
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
      const int e = q * 32 + lane;
      if (e < n4) __stcs(dst4 + (size_t)(p0 + (tid & ~31)) * (C / 4) + e, wst[e]);
    }
    __syncwarp();
  }
}

int launch_tore_tile(const Geom& g, const Workspace& ws, int k, float* out, cudaStream_t stream) {
  if (k == 6 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    const size_t smem6 = sizeof(uint32_t) * (size_t)g.tile_px * 12 + sizeof(float) * 12 * TILE_THREADS;
    EVREP_CUDA_OK(cudaFuncSetAttribute(k_tore_tile_k<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6));
    prof_begin(EVREP_K_TILE, stream);
    k_tore_tile_k<6, false><<<g.B * g.T, TILE_THREADS, smem6, stream>>>(ws.records, ws.base, ws.hist, ws.wp, g, out);
    prof_end(EVREP_K_TILE, stream);
    EVREP_CUDA_OK(cudaGetLastError());
    return EVREP_OK;
  }
  const size_t smem = sizeof(uint32_t) * (size_t)g.tile_px * 2 * k;
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_tore_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(EVREP_K_TILE, stream);
  k_tore_tile<<<g.B * g.T, TILE_THREADS, smem, stream>>>(ws.records, ws.base, ws.hist, ws.wp, g, k, out);
  prof_end(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

// EventStack(12) + TimeSurface(6 snapshots) + TORE(k = 6) from ONE binning pass (REC_T_IDX records, 1024-pixel tiles):
// BASELINE configs[2] asks for the three together, and two of the three binning passes were pure repetition.
int launch_order_ops_fused(const Geom& g, const Workspace& ws, double tau, float* out_es, float* out_ts, float* out_tore, cudaStream_t stream) {
  const size_t sm_es = sizeof(uint32_t) * (size_t)g.tile_px + sizeof(float) * 12 * TILE_THREADS;
  const size_t sm_ts = sizeof(uint32_t) * (size_t)g.tile_px * 12;
  const size_t sm_tore = sizeof(uint32_t) * (size_t)g.tile_px * 12 + sizeof(float) * 12 * TILE_THREADS;
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_event_stack_tile_k<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_es));
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_time_surface_tile_s<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_ts));
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_tore_tile_k<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_tore));
  prof_begin(EVREP_K_TILE, stream);
  k_event_stack_tile_k<12, true><<<g.B * g.T, TILE_THREADS, sm_es, stream>>>(ws.records, ws.base, ws.hist, ws.wp, g, out_es);
  k_time_surface_tile_s<6, true><<<g.B * g.T, TILE_THREADS, sm_ts, stream>>>(ws.records, ws.base, ws.hist, ws.wp, ws.snap, g, tau, out_ts);
  k_tore_tile_k<6, true><<<g.B * g.T, TILE_THREADS, sm_tore, stream>>>(ws.records, ws.base, ws.hist, ws.wp, g, out_tore);
  prof_end(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

}  // namespace evrep
