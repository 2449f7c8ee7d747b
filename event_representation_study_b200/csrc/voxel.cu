// Voxel grids (three flavours) and the 2-channel event histogram: additive representations, scattered
// straight into the (zeroed) output with L2 float reductions - one window's grid (<= 44 MB at 1 Mpx x 12
// bins) stays L2 resident while its events stream through.
//   tonic      tonic.transforms.ToVoxelGrid as called at representations/gen1_transforms.py:21-25
//   evlicious  ev-licious/src/evlicious/tools/utils.py:51-85 (+ :93-108), including its weight quirk
//   gwd        representations/representation_search/gromov_wasserstein.py:72-82 (compute_repr)
//   histogram  tonic.transforms.ToImage as called at gen1_transforms.py:44-49
#include <algorithm>

#include "evrep_common.cuh"

namespace evrep {

struct VoxelArgs {
  int flavour, n_bins, has_t0t1;
  int64_t t0, t1;
  int divider;  // ev-licious only: > 1 = x, y are sub-pixel integers, the event sits at (x / divider, y / divider)
};

template <typename TT>
__global__ void __launch_bounds__(256) k_voxel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, const TT* __restrict__ t,
                                               const int8_t* __restrict__ p, WinParams* __restrict__ wp, const Geom g, const VoxelArgs a,
                                               float* __restrict__ out) {
  const int b = blockIdx.y;
  const WinParams w = wp[b];
  if (w.n <= 0) return;
  if (a.flavour == EVREP_VOXEL_EVLICIOUS && w.n < 2) return;  // utils.py:52-53: fewer than 2 events -> zeros
  const int nb = a.n_bins;
  const int64_t t_first = w.t_base, t_last = w.t_base + w.tlast_rel;
  float* grid = out + (size_t)b * nb * g.HW;
  uint32_t flags = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < w.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ai = w.start + i;
    const uint32_t xv = x[ai], yv = y[ai];
    const int64_t tv = (int64_t)t[ai];
    int pv = p[ai];
    if (pv > 1 || pv < -1) { flags |= EVREP_WF_BAD_POLARITY; pv = pv > 0 ? 1 : -1; }
    if (a.flavour == EVREP_VOXEL_EVLICIOUS && a.divider > 1) {
      // Events.x = _x.astype(float32) / divider (events.py:37-47); _draw_xy_to_voxel_grid (utils.py:93-103): the four pixels
      // around (x, y) with weights (1 - |xlim - x|)(1 - |ylim - y|) in float64 (int32 - float32 promotes), taps outside the
      // grid dropped (:105-108); in time the floor bin gets p and the next bin weight 0 (the quirk of :74)
      const float xf = __fdiv_rn((float)xv, (float)a.divider), yf = __fdiv_rn((float)yv, (float)a.divider);
      if (xf > (float)(g.W - 1) || yf > (float)(g.H - 1)) { flags |= EVREP_WF_OUT_OF_RANGE; continue; }  // Events asserts max(x) <= width - 1
      const int64_t t0 = a.has_t0t1 ? a.t0 : t_first, t1 = a.has_t0t1 ? a.t1 : t_last;
      const double dT = (t1 - t0) == 0 ? 1.0 : (double)(t1 - t0);
      const double tn = (double)((int64_t)(nb - 1) * (tv - t0)) / dT;
      const int ti = (int)fmax(fmin(tn, 2.0e9), -2.0e9);
      if (ti < 0 || ti >= nb) continue;
      const double pol = pv == 0 ? -1.0 : (double)pv;
      const int xi = (int)xf, yi = (int)yf;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int xl = xi + dx, yl = yi + dy;
          if (xl >= g.W || yl >= g.H) continue;
          const double wgt = (1.0 - fabs((double)xl - (double)xf)) * (1.0 - fabs((double)yl - (double)yf));
          atomicAdd(grid + (size_t)ti * g.HW + (size_t)yl * g.W + xl, (float)(wgt * pol));
        }
      continue;
    }
    if (xv >= (uint32_t)g.W || yv >= (uint32_t)g.H) { flags |= EVREP_WF_OUT_OF_RANGE; continue; }
    const uint32_t lin = yv * (uint32_t)g.W + xv;
    if (a.flavour == EVREP_VOXEL_TONIC) {
      // ts = n_bins * (t - t[0]) / (t[-1] - t[0]); value p*(1-dt) into bin int(ts), p*dt into the next
      const double ts = (double)nb * (double)(tv - t_first) / (double)(t_last - t_first);
      if (!(ts >= 0.0)) { if (ts < 0.0) flags |= EVREP_WF_UNSORTED; continue; }  // NaN (t[-1] == t[0]): dropped
      const int ti = (int)fmin(ts, 2.0e9);
      const double dt = ts - (double)ti;
      const double pol = pv == 0 ? -1.0 : (double)pv;
      if (ti < nb) atomicAdd(grid + (size_t)ti * g.HW + lin, (float)(pol * (1.0 - dt)));
      if (ti + 1 < nb) atomicAdd(grid + (size_t)(ti + 1) * g.HW + lin, (float)(pol * dt));
    } else if (a.flavour == EVREP_VOXEL_EVLICIOUS) {
      // t_norm = (B-1)(t-t0)/dT, floor bin gets p, the "next" bin gets weight 0 (utils.py:74 passes t_norm_int)
      const int64_t t0 = a.has_t0t1 ? a.t0 : t_first, t1 = a.has_t0t1 ? a.t1 : t_last;
      const double dT = (t1 - t0) == 0 ? 1.0 : (double)(t1 - t0);
      const double tn = (double)((int64_t)(nb - 1) * (tv - t0)) / dT;
      const int ti = (int)fmax(fmin(tn, 2.0e9), -2.0e9);
      const float pol = pv == 0 ? -1.f : (float)pv;  // Events.__init__ maps p == 0 to -1 (events.py:20)
      if (ti >= 0 && ti < nb) atomicAdd(grid + (size_t)ti * g.HW + lin, pol);
    } else {
      // compute_repr: t01 in [0,1]; b = (bins-1) t; both neighbouring bins with weight 1 - |bin - b|
      const double t01 = (double)(tv - t_first) / (double)(t_last - t_first);
      const double bb = (double)(nb - 1) * t01;
      if (!(bb >= 0.0)) { if (bb < 0.0) flags |= EVREP_WF_UNSORTED; continue; }
      const int bi = (int)fmin(bb, 2.0e9);
      float* cell = out + ((size_t)b * g.HW + lin) * nb;  // (B, H, W, bins)
      if (bi < nb) atomicAdd(cell + bi, (float)((1.0 - fabs((double)bi - bb)) * (double)pv));
      if (bi + 1 < nb) atomicAdd(cell + bi + 1, (float)((1.0 - fabs((double)(bi + 1) - bb)) * (double)pv));
    }
  }
  flags = __reduce_or_sync(0xffffffffu, flags);
  if ((threadIdx.x & 31) == 0 && flags) atomicOr(&wp[b].flags, flags);
}

// ev-licious normalisation (utils.py:77-83): over the non-zero voxels of one window, (v - mean) / (std + 1e-5) if std > 0.
__global__ void __launch_bounds__(256) k_voxel_stats(const float* __restrict__ out, size_t per_window, double* __restrict__ stats) {
  const int b = blockIdx.y;
  const float* grid = out + (size_t)b * per_window;
  double s = 0.0, s2 = 0.0, c = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_window; i += (size_t)gridDim.x * blockDim.x) {
    const float v = grid[i];
    if (v != 0.f) { s += v; s2 += (double)v * v; c += 1.0; }
  }
  for (int d = 16; d > 0; d >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    c += __shfl_xor_sync(0xffffffffu, c, d);
  }
  if ((threadIdx.x & 31) == 0 && c > 0.0) {  // voxel values are integers: these double sums are exact, hence order independent
    atomicAdd(stats + 4 * b + 0, s);
    atomicAdd(stats + 4 * b + 1, s2);
    atomicAdd(stats + 4 * b + 2, c);
  }
}

__global__ void __launch_bounds__(256) k_voxel_norm(float* __restrict__ out, size_t per_window, const double* __restrict__ stats) {
  const int b = blockIdx.y;
  const double c = stats[4 * b + 2];
  if (c <= 0.0) return;
  const double mean = stats[4 * b + 0] / c;
  const double var = fmax(stats[4 * b + 1] / c - mean * mean, 0.0);
  const double sd = sqrt(var);
  if (!(sd > 0.0)) return;
  const float fm = (float)mean, fs = (float)(1e-5 + (double)(float)sd);
  float* grid = out + (size_t)b * per_window;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_window; i += (size_t)gridDim.x * blockDim.x) {
    const float v = grid[i];
    if (v != 0.f) grid[i] = (v - fm) / fs;
  }
}

int launch_voxel(const Events& ev, const int64_t* win_offsets_host, const Geom& g, const Workspace& ws, int flavour, int n_bins,
                 int normalize, const int64_t* t0_t1_host, int divider, float* out, cudaStream_t stream) {
  int n_chunks = 0;
  int rc = prepare_windows(ev, win_offsets_host, g, ws, &n_chunks, stream);
  if (rc) return rc;
  const size_t per_window = (size_t)n_bins * g.HW;
  EVREP_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * per_window * g.B, stream));
  VoxelArgs a;
  a.flavour = flavour;
  a.n_bins = n_bins;
  a.has_t0t1 = t0_t1_host != nullptr;
  a.t0 = t0_t1_host ? t0_t1_host[0] : 0;
  a.t1 = t0_t1_host ? t0_t1_host[1] : 0;
  a.divider = divider;
  int64_t n_max = 0;
  for (int b = 0; b < g.B; ++b) n_max = std::max<int64_t>(n_max, win_offsets_host[b + 1] - win_offsets_host[b]);
  if (n_max > 0) {
    dim3 grid((unsigned)std::min<int64_t>((n_max + 255) / 256, 148 * 8), g.B);
    if (ev.t_bytes == 4)
      k_voxel<int32_t><<<grid, 256, 0, stream>>>(ev.x, ev.y, (const int32_t*)ev.t, ev.p, ws.wp, g, a, out);
    else
      k_voxel<int64_t><<<grid, 256, 0, stream>>>(ev.x, ev.y, (const int64_t*)ev.t, ev.p, ws.wp, g, a, out);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  if (flavour == EVREP_VOXEL_EVLICIOUS && normalize) {
    EVREP_CUDA_OK(cudaMemsetAsync(ws.stats, 0, sizeof(double) * 4 * g.B, stream));
    dim3 grid((unsigned)std::min<size_t>((per_window + 255) / 256, 148 * 4), g.B);
    k_voxel_stats<<<grid, 256, 0, stream>>>(out, per_window, ws.stats);
    EVREP_CUDA_OK(cudaGetLastError());
    k_voxel_norm<<<grid, 256, 0, stream>>>(out, per_window, ws.stats);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  return EVREP_OK;
}

__global__ void __launch_bounds__(256) k_histogram(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                                                   const int8_t* __restrict__ p, WinParams* __restrict__ wp, const Geom g,
                                                   float* __restrict__ out) {
  const int b = blockIdx.y;
  const WinParams w = wp[b];
  float* grid = out + (size_t)b * 2 * g.HW;
  uint32_t flags = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < w.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ai = w.start + i;
    const uint32_t xv = x[ai], yv = y[ai];
    if (xv >= (uint32_t)g.W || yv >= (uint32_t)g.H) { flags |= EVREP_WF_OUT_OF_RANGE; continue; }
    atomicAdd(grid + (size_t)(p[ai] > 0 ? 1 : 0) * g.HW + yv * (uint32_t)g.W + xv, 1.f);
  }
  flags = __reduce_or_sync(0xffffffffu, flags);
  if ((threadIdx.x & 31) == 0 && flags) atomicOr(&wp[b].flags, flags);
}

int launch_histogram(const Events& ev, const int64_t* win_offsets_host, const Geom& g, const Workspace& ws, float* out,
                     cudaStream_t stream) {
  int n_chunks = 0;
  int rc = prepare_windows(ev, win_offsets_host, g, ws, &n_chunks, stream);
  if (rc) return rc;
  EVREP_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * 2 * (size_t)g.HW * g.B, stream));
  int64_t n_max = 0;
  for (int b = 0; b < g.B; ++b) n_max = std::max<int64_t>(n_max, win_offsets_host[b + 1] - win_offsets_host[b]);
  if (n_max > 0) {
    dim3 grid((unsigned)std::min<int64_t>((n_max + 255) / 256, 148 * 8), g.B);
    k_histogram<<<grid, 256, 0, stream>>>(ev.x, ev.y, ev.p, ws.wp, g, out);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  return EVREP_OK;
}

}  // namespace evrep
