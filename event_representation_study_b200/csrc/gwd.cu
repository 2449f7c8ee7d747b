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
constexpr int GW_CHUNK = 16;  // tiles of one row band per CTA: a band longer than this is cut, so that one huge pair still fills the machine

struct GwPair {      // device, one per pair
  int64_t s0, t0;    // first row of Xs / Xt
  int32_t n, m;      // rows
  int32_t tile0;     // first CTA (= partial sum) of this pair
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

// 64 rows of X, centred and converted to float, into dst[r * stride + k]; four threads per row, no integer division
__device__ __forceinline__ void gw_load_rows(float* dst, int stride, const double* __restrict__ X, const double* mu, int row0, int n, int d) {
  const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
  const bool live = row0 + r < n;
  const double* src = X + (size_t)(row0 + r) * d;
  for (int k = q; k < d; k += 4) dst[r * stride + k] = live ? (float)(src[k] - mu[k]) : 0.f;
}
// squared norms of the 64 rows of a loaded tile, with the same fmaf chain over k as the dot products
__device__ __forceinline__ void gw_row_norms(float* nrm, const float* src, int stride, int d) {
  if (threadIdx.x < GW_TILE) {
    float a = 0.f;
    for (int k = 0; k < d; ++k) a = fmaf(src[threadIdx.x * stride + k], src[threadIdx.x * stride + k], a);
    nrm[threadIdx.x] = a;
  }
}

// |xi - xj|^2 = |xi|^2 + |xj|^2 - 2 xi.xj on the CENTRED coordinates: one FMA per dimension and cell instead of a subtraction
// and an FMA (the kernel is issue bound: ncu r02, 75 % issue-slot utilisation at 145 instructions per cell).  The norms are
// accumulated with the same FMA chain as the dot product, so the diagonal (i == j) cancels to exactly 0; elsewhere the
// cancellation costs <= 2^-23 (|xi|^2 + |xj|^2) in d^2, i.e. a few 1e-7 absolute in K = exp(-d^2 / (h^2 mean d^2)).
// exp through ex2.approx on the pre-scaled argument (2 ulp).
__device__ __forceinline__ float gw_kernel_value(float ni, float nj, float dot, float coef_log2e) {
  const float d2 = fmaxf(fmaf(-2.f, dot, ni + nj), 0.f);
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-d2 * coef_log2e));
  return r;
}

// One CTA per piece of a ROW BAND of one pair: tile row bi against up to GW_CHUNK of the tile columns bj = bi .. nt - 1 (the
// matrix is symmetric).  The band's own rows (si, ti and their norms) are loaded once per CTA; per tile only the column rows
// are.  band0[p] = first CTA of pair p; inside a pair the CTAs go band by band, piece by piece.
__global__ void __launch_bounds__(GW_THREADS) k_gw_tiles(const double* __restrict__ Xs, const double* __restrict__ Xt, int ds, int dt,
                                                         const GwPair* __restrict__ pairs, const int32_t* __restrict__ band0, int n_pairs,
                                                         double* __restrict__ partial) {
  extern __shared__ float sh[];
  const int sds = ds | 1, sdt = dt | 1;
  float* si = sh;
  float* sj = si + GW_TILE * sds;
  float* ti = sj + GW_TILE * sds;
  float* tj = ti + GW_TILE * sdt;
  __shared__ float nrm[4][GW_TILE];  // squared norms of the rows of si, sj, ti, tj
  __shared__ double red[GW_THREADS / 32];
  __shared__ int s_pair, s_bi, s_bj0;

  if (threadIdx.x == 0) {  // which pair: binary search over the compact first-CTA table, once per CTA; then which band and piece
    int lo = 0, hi = n_pairs;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(band0 + mid) <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    s_pair = lo;
    const int nt = pairs[lo].nt;
    int c = (int)blockIdx.x - __ldg(band0 + lo), b = 0;
    while (c >= (nt - b + GW_CHUNK - 1) / GW_CHUNK) { c -= (nt - b + GW_CHUNK - 1) / GW_CHUNK; ++b; }
    s_bi = b;
    s_bj0 = b + c * GW_CHUNK;
  }
  __syncthreads();
  const GwPair& P = pairs[s_pair];
  const int bi = s_bi, bj0 = s_bj0, bj1 = min(P.nt, bj0 + GW_CHUNK);
  const double* Xsp = Xs + P.s0 * ds;
  const double* Xtp = Xt + P.t0 * dt;
  gw_load_rows(si, sds, Xsp, P.mu_s, bi * GW_TILE, P.n, ds);
  gw_load_rows(ti, sdt, Xtp, P.mu_t, bi * GW_TILE, P.m, dt);
  __syncthreads();
  gw_row_norms(nrm[0], si, sds, ds);
  if (threadIdx.x >= GW_THREADS - GW_TILE) {  // another 64 threads take the second array
    const int r = threadIdx.x - (GW_THREADS - GW_TILE);
    float a = 0.f;
    for (int k = 0; k < dt; ++k) a = fmaf(ti[r * sdt + k], ti[r * sdt + k], a);
    nrm[2][r] = a;
  }

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float cs = (float)(P.coef_s * 1.4426950408889634), ct = (float)(P.coef_t * 1.4426950408889634);
  const int inner = min(P.n, P.m);
  double band = 0.0;  // this thread's share of the band, tiles added in column order
  for (int bj = bj0; bj < bj1; ++bj) {
    __syncthreads();  // the previous tile's readers are done with sj / tj (and the norms above are written)
    gw_load_rows(sj, sds, Xsp, P.mu_s, bj * GW_TILE, P.n, ds);
    gw_load_rows(tj, sdt, Xtp, P.mu_t, bj * GW_TILE, P.m, dt);
    __syncthreads();
    gw_row_norms(nrm[1], sj, sds, ds);
    if (threadIdx.x >= GW_THREADS - GW_TILE) {
      const int r = threadIdx.x - (GW_THREADS - GW_TILE);
      float a = 0.f;
      for (int k = 0; k < dt; ++k) a = fmaf(tj[r * sdt + k], tj[r * sdt + k], a);
      nrm[3][r] = a;
    }
    float dts[4][4], dtt[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) dts[a][b] = dtt[a][b] = 0.f;
    for (int k = 0; k < ds; ++k) {
      float vi[4], vj[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { vi[a] = si[(ty + 16 * a) * sds + k]; vj[a] = sj[(tx + 16 * a) * sds + k]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dts[a][b] = fmaf(vi[a], vj[b], dts[a][b]);
    }
    for (int k = 0; k < dt; ++k) {
      float vi[4], vj[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { vi[a] = ti[(ty + 16 * a) * sdt + k]; vj[a] = tj[(tx + 16 * a) * sdt + k]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dtt[a][b] = fmaf(vi[a], vj[b], dtt[a][b]);
    }
    __syncthreads();  // norms of this tile
    float nsi[4], nsj[4], nti[4], ntj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      nsi[a] = nrm[0][ty + 16 * a], nsj[a] = nrm[1][tx + 16 * a];
      nti[a] = nrm[2][ty + 16 * a], ntj[a] = nrm[3][tx + 16 * a];
    }
    float acc = 0.f;
    if ((bj + 1) * GW_TILE <= inner && (bi + 1) * GW_TILE <= inner) {  // CTA-uniform: the tile lies inside both matrices (all but the rim)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc += fabsf(gw_kernel_value(nsi[a], nsj[b], dts[a][b], cs) - gw_kernel_value(nti[a], ntj[b], dtt[a][b], ct));
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = bi * GW_TILE + ty + 16 * a, j = bj * GW_TILE + tx + 16 * b;
          const float ks = (i < P.n && j < P.n) ? gw_kernel_value(nsi[a], nsj[b], dts[a][b], cs) : 0.f;
          const float kt = (i < P.m && j < P.m) ? gw_kernel_value(nti[a], ntj[b], dtt[a][b], ct) : 0.f;
          acc += fabsf(ks - kt);
        }
    }
    band += (bi == bj) ? (double)acc : 2.0 * (double)acc;  // symmetric: the mirrored tile contributes the same
  }
  double s = band;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int q = 0; q < GW_THREADS / 32; ++q) a += red[q];
    partial[blockIdx.x] = a;
  }
}

__global__ void __launch_bounds__(GW_THREADS) k_gw_finish(const GwPair* __restrict__ pairs, const double* __restrict__ partial,
                                                          double* __restrict__ out) {
  const GwPair& P = pairs[blockIdx.x];
  int n_tiles = 0;  // one partial per CTA: pieces of the row bands
  for (int b = 0; b < P.nt; ++b) n_tiles += (P.nt - b + GW_CHUNK - 1) / GW_CHUNK;
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
    for (int64_t b = 0; b < nt; ++b) acc += (nt - b + GW_CHUNK - 1) / GW_CHUNK;  // one CTA and one partial sum per piece of a row band
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
  return align_up(sizeof(GwPair) * (size_t)n_pairs, 256) + align_up(sizeof(double) * (size_t)(tiles + 1), 256) + align_up(sizeof(int32_t) * (size_t)n_pairs, 256);
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
  int32_t* d_tile0 = (int32_t*)((char*)d_partial + align_up(sizeof(double) * (size_t)(tiles + 1), 256));
  std::vector<int32_t> tile0((size_t)n_pairs);
  for (int i = 0; i < n_pairs; ++i) tile0[(size_t)i] = pairs[(size_t)i].tile0;
  EVREP_CUDA_OK(cudaMemcpyAsync(d_pairs, pairs.data(), sizeof(GwPair) * (size_t)n_pairs, cudaMemcpyHostToDevice, stream));
  EVREP_CUDA_OK(cudaMemcpyAsync(d_tile0, tile0.data(), sizeof(int32_t) * (size_t)n_pairs, cudaMemcpyHostToDevice, stream));
  k_gw_moments<<<dim3(n_pairs, 2), GW_THREADS, 0, stream>>>(Xs, Xt, ds, dt, h, d_pairs);
  EVREP_CUDA_OK(cudaGetLastError());
  if (tiles > 0) {
    const size_t smem = sizeof(float) * GW_TILE * 2 * (size_t)((ds | 1) + (dt | 1));
    EVREP_CUDA_OK(cudaFuncSetAttribute(k_gw_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gw_tiles<<<(unsigned)tiles, GW_THREADS, smem, stream>>>(Xs, Xt, ds, dt, d_pairs, d_tile0, n_pairs, d_partial);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  k_gw_finish<<<n_pairs, GW_THREADS, 0, stream>>>(d_pairs, d_partial, out);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

}  // namespace evrep
