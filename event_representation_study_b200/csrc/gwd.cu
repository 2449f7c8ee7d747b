// GWD-A: the ranking distance the paper pipeline actually evaluates
// (representations/representation_search/compute_otmi.py:50-93 with max_iter=0 and a loss that ignores its
// arguments): cost = mean over the L x L zero-padded grid of |Ks - Kt|, L = max(n, m), with Gaussian
// kernels K = exp(-(D / (h std))^2 / 2), std = sqrt(mean(D^2) / 2) (compute_otmi.py:6-32) on the
// Euclidean distance matrices of the event samples Xs (n x ds) and the representation pixels Xt (m x dt).
// Nothing n x n is ever materialised: each CTA evaluates one 64 x 64 tile of both kernels from the
// coordinates and reduces |Ks - Kt| on the fly; the matrix is symmetric, so only tiles on or above the
// diagonal are visited.  fp32 per cell (coordinates centred first), fp64 reductions, fixed summation
// order => bit-reproducible.
#include <string.h>

#include <vector>

#include "evrep_common.cuh"

namespace evrep {

constexpr int GW_TILE = 64;
constexpr int GW_THREADS = 256;
constexpr int GW_MAX_D = 64;

struct GwPair {      // device, one per pair
  int64_t s0, t0;    // first row of Xs / Xt
  int32_t n, m;      // rows
  int32_t tile0;     // first tile index of this pair
  int32_t nt;        // tiles per side: ceil(max(n, m) / 64)
  double coef_s, coef_t;  // 1 / (h^2 mean(D^2)):  K = exp(-D^2 * coef)
  double mu_s[GW_MAX_D], mu_t[GW_MAX_D];
};

// mean vector and mean squared pairwise distance: mean_ij |xi - xj|^2 = 2/n * sum_i |xi - mu|^2
__global__ void __launch_bounds__(GW_THREADS) k_gw_moments(const double* __restrict__ Xs, const double* __restrict__ Xt, int ds, int dt,
                                                           double h, GwPair* __restrict__ pairs) {
  GwPair& P = pairs[blockIdx.x];
  const int side = blockIdx.y;
  const double* X = side == 0 ? Xs + P.s0 * ds : Xt + P.t0 * dt;
  const int n = side == 0 ? P.n : P.m, d = side == 0 ? ds : dt;
  __shared__ double mu[GW_MAX_D];
  __shared__ double red[GW_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = 0; k < d; ++k) {
    double s = 0.0;
    for (int i = tid; i < n; i += GW_THREADS) s += X[(size_t)i * d + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double a = 0.0;
      for (int q = 0; q < GW_THREADS / 32; ++q) a += red[q];
      mu[k] = n > 0 ? a / n : 0.0;
    }
    __syncthreads();
  }
  double s = 0.0;
  for (int i = tid; i < n; i += GW_THREADS)
    for (int k = 0; k < d; ++k) {
      const double v = X[(size_t)i * d + k] - mu[k];
      s += v * v;
    }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    for (int q = 0; q < GW_THREADS / 32; ++q) a += red[q];
    const double mean_d2 = n > 0 ? 2.0 * a / n : 0.0;
    const double coef = 1.0 / (h * h * mean_d2);  // inf when all points coincide -> NaN kernel, like the reference
    if (side == 0) P.coef_s = coef; else P.coef_t = coef;
  }
  if (tid < d) (side == 0 ? P.mu_s : P.mu_t)[tid] = mu[tid];
}

__device__ __forceinline__ void gw_load_rows(float* dst, int stride, const double* __restrict__ X, const double* mu, int row0, int n, int d) {
  for (int e = threadIdx.x; e < GW_TILE * d; e += GW_THREADS) {
    const int r = e / d, k = e - r * d;
    dst[r * stride + k] = (row0 + r < n) ? (float)(X[(size_t)(row0 + r) * d + k] - mu[k]) : 0.f;
  }
}

__global__ void __launch_bounds__(GW_THREADS) k_gw_tiles(const double* __restrict__ Xs, const double* __restrict__ Xt, int ds, int dt,
                                                         const GwPair* __restrict__ pairs, int n_pairs, double* __restrict__ partial) {
  extern __shared__ float sh[];
  const int sds = ds | 1, sdt = dt | 1;
  float* si = sh;
  float* sj = si + GW_TILE * sds;
  float* ti = sj + GW_TILE * sds;
  float* tj = ti + GW_TILE * sdt;
  __shared__ double red[GW_THREADS / 32];

  // which pair / which tile
  int lo = 0, hi = n_pairs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pairs[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const GwPair& P = pairs[lo];
  int rem = blockIdx.x - P.tile0, bi = 0;
  while (rem >= P.nt - bi) { rem -= P.nt - bi; ++bi; }  // row bi holds tiles bj = bi .. nt-1
  const int bj = bi + rem;

  gw_load_rows(si, sds, Xs + P.s0 * ds, P.mu_s, bi * GW_TILE, P.n, ds);
  gw_load_rows(sj, sds, Xs + P.s0 * ds, P.mu_s, bj * GW_TILE, P.n, ds);
  gw_load_rows(ti, sdt, Xt + P.t0 * dt, P.mu_t, bi * GW_TILE, P.m, dt);
  gw_load_rows(tj, sdt, Xt + P.t0 * dt, P.mu_t, bj * GW_TILE, P.m, dt);
  __syncthreads();

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float cs = (float)P.coef_s, ct = (float)P.coef_t;
  float d2s[4][4], d2t[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) d2s[a][b] = d2t[a][b] = 0.f;
  for (int k = 0; k < ds; ++k) {
    float vi[4], vj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { vi[a] = si[(ty + 16 * a) * sds + k]; vj[a] = sj[(tx + 16 * a) * sds + k]; }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { const float df = vi[a] - vj[b]; d2s[a][b] = fmaf(df, df, d2s[a][b]); }
  }
  for (int k = 0; k < dt; ++k) {
    float vi[4], vj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { vi[a] = ti[(ty + 16 * a) * sdt + k]; vj[a] = tj[(tx + 16 * a) * sdt + k]; }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { const float df = vi[a] - vj[b]; d2t[a][b] = fmaf(df, df, d2t[a][b]); }
  }
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = bi * GW_TILE + ty + 16 * a, j = bj * GW_TILE + tx + 16 * b;
      const float ks = (i < P.n && j < P.n) ? expf(-d2s[a][b] * cs) : 0.f;
      const float kt = (i < P.m && j < P.m) ? expf(-d2t[a][b] * ct) : 0.f;
      acc += fabsf(ks - kt);
    }
  double s = (double)acc;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int q = 0; q < GW_THREADS / 32; ++q) a += red[q];
    partial[blockIdx.x] = (bi == bj) ? a : 2.0 * a;  // symmetric: the mirrored tile contributes the same
  }
}

__global__ void __launch_bounds__(GW_THREADS) k_gw_finish(const GwPair* __restrict__ pairs, const double* __restrict__ partial,
                                                          double* __restrict__ out) {
  const GwPair& P = pairs[blockIdx.x];
  const int n_tiles = P.nt * (P.nt + 1) / 2;
  __shared__ double red[GW_THREADS];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_tiles; i += GW_THREADS) s += partial[P.tile0 + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = GW_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double L = (double)max(P.n, P.m);
    out[blockIdx.x] = L > 0 ? red[0] / (L * L) : 0.0;
  }
}

static int gw_tiles(const int64_t* so, const int64_t* to, int n_pairs, std::vector<GwPair>* pairs, int64_t* total_tiles) {
  int64_t acc = 0;
  for (int i = 0; i < n_pairs; ++i) {
    const int64_t n = so[i + 1] - so[i], m = to[i + 1] - to[i];
    if (n < 0 || m < 0 || n > (1 << 24) || m > (1 << 24)) {
      set_error("pair %d: row counts %lld / %lld outside 0..2^24", i, (long long)n, (long long)m);
      return EVREP_EINVAL;
    }
    const int64_t L = n > m ? n : m, nt = (L + GW_TILE - 1) / GW_TILE;
    if (pairs) {
      GwPair P;
      memset(&P, 0, sizeof(P));
      P.s0 = so[i]; P.t0 = to[i]; P.n = (int32_t)n; P.m = (int32_t)m; P.tile0 = (int32_t)acc; P.nt = (int32_t)nt;
      pairs->push_back(P);
    }
    acc += nt * (nt + 1) / 2;
    if (acc > (int64_t)0x7fffffff) {
      set_error("too many tiles in one call; split the batch of pairs");
      return EVREP_EUNSUPPORTED;
    }
  }
  *total_tiles = acc;
  return EVREP_OK;
}

size_t gwd_workspace_bytes(const int64_t* so, const int64_t* to, int n_pairs) {
  int64_t tiles = 0;
  if (n_pairs <= 0 || gw_tiles(so, to, n_pairs, nullptr, &tiles)) return 0;
  return align_up(sizeof(GwPair) * (size_t)n_pairs, 256) + align_up(sizeof(double) * (size_t)(tiles + 1), 256);
}

int launch_gwd(const double* Xs, const int64_t* so, int ds, const double* Xt, const int64_t* to, int dt, int n_pairs, double h,
               double* out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (ds < 1 || ds > GW_MAX_D || dt < 1 || dt > GW_MAX_D) {
    set_error("feature widths ds=%d dt=%d outside 1..%d", ds, dt, GW_MAX_D);
    return EVREP_EUNSUPPORTED;
  }
  std::vector<GwPair> pairs;
  int64_t tiles = 0;
  int rc = gw_tiles(so, to, n_pairs, &pairs, &tiles);
  if (rc) return rc;
  const size_t need = gwd_workspace_bytes(so, to, n_pairs);
  if (workspace_bytes < need || !workspace) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return EVREP_EWORKSPACE;
  }
  GwPair* d_pairs = (GwPair*)workspace;
  double* d_partial = (double*)((char*)workspace + align_up(sizeof(GwPair) * (size_t)n_pairs, 256));
  EVREP_CUDA_OK(cudaMemcpyAsync(d_pairs, pairs.data(), sizeof(GwPair) * (size_t)n_pairs, cudaMemcpyHostToDevice, stream));
  k_gw_moments<<<dim3(n_pairs, 2), GW_THREADS, 0, stream>>>(Xs, Xt, ds, dt, h, d_pairs);
  EVREP_CUDA_OK(cudaGetLastError());
  if (tiles > 0) {
    const size_t smem = sizeof(float) * GW_TILE * 2 * (size_t)((ds | 1) + (dt | 1));
    EVREP_CUDA_OK(cudaFuncSetAttribute(k_gw_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gw_tiles<<<(unsigned)tiles, GW_THREADS, smem, stream>>>(Xs, Xt, ds, dt, d_pairs, n_pairs, d_partial);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  k_gw_finish<<<n_pairs, GW_THREADS, 0, stream>>>(d_pairs, d_partial, out);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

}  // namespace evrep
