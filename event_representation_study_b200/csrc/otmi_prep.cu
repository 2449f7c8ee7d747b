// The data preparation of otmi() (representations/representation_search/compute_otmi.py:96-203) on the device:
// events (N x 4: x, y, t, p) are split into the four sensor quadrants with the reference's inclusive / exclusive bounds,
// the densest quadrant is dropped, the other three are rebased, normalised and filtered into the source point sets Xs; the
// representation (rep_size x rep_size x C) is cropped per quadrant, gets two positional channels, loses its all-zero pixels
// and becomes the target point sets Xt.  Both compactions keep the input order (stream order / row-major pixel order), as
// the boolean-mask indexing of the reference does, so the point sets are bit-identical to it and deterministic.
//
// One pass for the per-quadrant statistics (count, min x, min y, first / last event, min / max polarity), one tiny kernel
// that picks the quadrant to drop and the three output slots, then count -> scan -> write over fixed blocks of 1024 items
// (events and crop pixels alike).  The arithmetic runs in the dtype of the event array, as the reference's expressions do on
// the array they are given: its own caller passes a torch int32 tensor (gen1_compute.py:57-59), so the differences are exact
// integers and the quotients torch's int / int true division, i.e. float32; float32 / float64 arrays divide in their own
// precision.  The point sets are stored as float64 for evrep_gwd_kernel_l1.
#include <math.h>

#include "evrep_common.cuh"

namespace evrep {
namespace {

constexpr int OT_BLOCK = 1024;  // items per compaction block (= threads per CTA)

template <typename T> struct QuotientOf { using type = T; };
template <> struct QuotientOf<int32_t> { using type = float; };  // torch: int tensor / int -> float32

struct OtmiStats {  // per quadrant
  unsigned long long count;
  unsigned long long min_x, min_y, min_p, max_p;  // order-preserving keys of the values
  long long first, last;                           // event indices
};
struct OtmiPlan {
  int slot[4];            // output slot of the quadrant (0..2), -1 for the dropped one
  int dropped;
  int empty_quadrant;     // 1 + index of an empty quadrant among 1..3 (the reference's min() raises), 0 if none
  double min_x[4], min_y[4], t0[4], t1[4], p_min[4], p_max[4];
  long long n_in[4];      // events of the quadrant before the final mask
};

template <typename T>
__device__ __forceinline__ unsigned long long ord_key(T v) {  // monotone map of a float value onto unsigned integers
  const double d = (double)v;
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_value(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// compute_otmi.py:97-132 - note the mixed >= / > on the inner edges
template <typename T>
__device__ __forceinline__ int quadrant_of(T x, T y, double w2, double h2, double w1, double h1) {
  const double xd = (double)x, yd = (double)y;
  const bool left = xd >= 0.0 && xd <= w2, right = xd > w2 && xd <= w1;
  const bool top = yd >= 0.0 && yd <= h2, bottom = yd > h2 && yd <= h1;
  if (left && top) return 0;
  if (right && top) return 1;
  if (left && bottom) return 2;
  if (right && bottom) return 3;
  return -1;
}

template <typename T>
__global__ void __launch_bounds__(256) k_otmi_stats(const T* __restrict__ ev, long long N, double w2, double h2, double w1, double h1,
                                                    OtmiStats* __restrict__ st) {
  __shared__ OtmiStats sh[4];
  if (threadIdx.x < 4) {
    sh[threadIdx.x].count = 0;
    sh[threadIdx.x].min_x = sh[threadIdx.x].min_y = sh[threadIdx.x].min_p = ~0ull;
    sh[threadIdx.x].max_p = 0ull;
    sh[threadIdx.x].first = 0x7fffffffffffffffll;
    sh[threadIdx.x].last = -1;
  }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const T x = ev[4 * i], y = ev[4 * i + 1], p = ev[4 * i + 3];
    const int q = quadrant_of(x, y, w2, h2, w1, h1);
    if (q < 0) continue;
    atomicAdd(&sh[q].count, 1ull);
    atomicMin(&sh[q].min_x, ord_key(x));
    atomicMin(&sh[q].min_y, ord_key(y));
    atomicMin(&sh[q].min_p, ord_key(p));
    atomicMax(&sh[q].max_p, ord_key(p));
    atomicMin(&sh[q].first, i);
    atomicMax(&sh[q].last, i);
  }
  __syncthreads();
  if (threadIdx.x < 4 && sh[threadIdx.x].count) {
    const int q = threadIdx.x;
    atomicAdd(&st[q].count, sh[q].count);
    atomicMin(&st[q].min_x, sh[q].min_x);
    atomicMin(&st[q].min_y, sh[q].min_y);
    atomicMin(&st[q].min_p, sh[q].min_p);
    atomicMax(&st[q].max_p, sh[q].max_p);
    atomicMin(&st[q].first, sh[q].first);
    atomicMax(&st[q].last, sh[q].last);
  }
}

__global__ void k_otmi_stats_init(OtmiStats* st) {
  const int q = threadIdx.x;
  if (q < 4) {
    st[q].count = 0;
    st[q].min_x = st[q].min_y = st[q].min_p = ~0ull;
    st[q].max_p = 0ull;
    st[q].first = 0x7fffffffffffffffll;
    st[q].last = -1;
  }
}

template <typename T>
__global__ void k_otmi_plan(const T* __restrict__ ev, const OtmiStats* __restrict__ st, OtmiPlan* __restrict__ plan) {
  if (threadIdx.x != 0) return;
  OtmiPlan P;
  int ind = 0;  // shape_sizes.index(max(shape_sizes)): the first maximum
  for (int q = 1; q < 4; ++q)
    if (st[q].count > st[ind].count) ind = q;
  P.dropped = ind;
  P.empty_quadrant = 0;
  int s = 0;
  for (int q = 0; q < 4; ++q) {
    P.slot[q] = (q == ind) ? -1 : s++;
    P.n_in[q] = (long long)st[q].count;
    const bool has = st[q].count > 0;
    if (!has && q > 0 && !P.empty_quadrant) P.empty_quadrant = q + 1;  // min() of an empty column (compute_otmi.py:140-147)
    // quadrant 0 "stays as expected": no rebase
    P.min_x[q] = (q > 0 && has) ? ord_value(st[q].min_x) : 0.0;
    P.min_y[q] = (q > 0 && has) ? ord_value(st[q].min_y) : 0.0;
    P.t0[q] = has ? (double)ev[4 * st[q].first + 2] : 0.0;
    P.t1[q] = has ? (double)ev[4 * st[q].last + 2] : 0.0;
    P.p_min[q] = has ? ord_value(st[q].min_p) : 0.0;
    P.p_max[q] = has ? ord_value(st[q].max_p) : 0.0;
  }
  *plan = P;
}

// an event's quadrant if it survives the final mask (rebased x < (W - 1) // 2 and y < (H - 1) // 2, compute_otmi.py:170-172)
template <typename T>
__device__ __forceinline__ int kept_quadrant(const T* ev, long long i, const OtmiPlan& P, double w2, double h2, double w1, double h1, T wd, T hd) {
  const T x = ev[4 * i], y = ev[4 * i + 1];
  const int q = quadrant_of(x, y, w2, h2, w1, h1);
  if (q < 0 || P.slot[q] < 0) return -1;
  const T xr = (T)(x - (T)P.min_x[q]), yr = (T)(y - (T)P.min_y[q]);
  return (xr < wd && yr < hd) ? q : -1;
}

// block-wide order-preserving ranks: returns this thread's rank among the flagged threads of the CTA and the CTA total
__device__ __forceinline__ uint32_t block_rank(bool flag, uint32_t* warp_cnt, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bal = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  uint32_t before = 0, tot = 0;
  for (uint32_t w = 0; w < OT_BLOCK / 32; ++w) {
    const uint32_t c = warp_cnt[w];
    if (w < warp) before += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return before + __popc(bal & ((1u << lane) - 1u));
}

// pass A (WRITE = false): per block and slot the number of survivors; pass B (WRITE = true): the rows
template <typename T, bool WRITE>
__global__ void __launch_bounds__(OT_BLOCK) k_otmi_events(const T* __restrict__ ev, long long N, const OtmiPlan* __restrict__ plan, double w2, double h2,
                                                          double w1, double h1, int wdiv, int hdiv, uint32_t* __restrict__ blk_cnt /* [3][nblk] */,
                                                          const long long* __restrict__ blk_off /* [3][nblk] */, double* __restrict__ Xs, long long cap) {
  __shared__ uint32_t warp_cnt[OT_BLOCK / 32];
  __shared__ OtmiPlan P;
  if (threadIdx.x == 0) P = *plan;
  __syncthreads();
  const long long i = (long long)blockIdx.x * OT_BLOCK + threadIdx.x;
  const int q = i < N ? kept_quadrant(ev, i, P, w2, h2, w1, h1, (T)wdiv, (T)hdiv) : -1;
  for (int qq = 0; qq < 4; ++qq) {
    const int s = P.slot[qq];
    if (s < 0) continue;
    uint32_t total;
    const uint32_t r = block_rank(q == qq, warp_cnt, &total);
    if (!WRITE) {
      if (threadIdx.x == 0) blk_cnt[(size_t)s * gridDim.x + blockIdx.x] = total;
    } else if (q == qq) {
      // compute_otmi.py:164-169: differences in the array's own dtype, quotients in its division dtype
      using F = typename QuotientOf<T>::type;
      const F x = (F)(T)(ev[4 * i] - (T)P.min_x[qq]) / (F)wdiv;
      const F y = (F)(T)(ev[4 * i + 1] - (T)P.min_y[qq]) / (F)hdiv;
      const F t = (F)(T)(ev[4 * i + 2] - (T)P.t0[qq]) / (F)(T)((T)P.t1[qq] - (T)P.t0[qq]);
      const F p = (F)(T)(ev[4 * i + 3] - (T)P.p_min[qq]) / (F)(T)((T)P.p_max[qq] - (T)P.p_min[qq]);
      double* row = Xs + ((size_t)s * cap + (size_t)(blk_off[(size_t)s * gridDim.x + blockIdx.x] + r)) * 4;
      row[0] = (double)x, row[1] = (double)y, row[2] = (double)t, row[3] = (double)p;
    }
  }
}

struct OtmiCrops {
  int y0[4], y1[4], x0[4], x1[4];  // inclusive crop bounds of the representation per quadrant (compute_otmi.py:150-158, 176-180)
};

// crop pixels: item = row-major pixel of quadrant q's crop; kept when some channel is non-zero (compute_otmi.py:197-198)
template <bool WRITE>
__global__ void __launch_bounds__(OT_BLOCK) k_otmi_pixels(const double* __restrict__ rep, int R, int C, const OtmiPlan* __restrict__ plan, const OtmiCrops crops,
                                                          int nblk_q /* blocks per quadrant */, uint32_t* __restrict__ blk_cnt /* [3][nblk_q] */,
                                                          const long long* __restrict__ blk_off, double* __restrict__ Xt, long long cap) {
  __shared__ uint32_t warp_cnt[OT_BLOCK / 32];
  const int qq = blockIdx.x / nblk_q, blk = blockIdx.x - qq * nblk_q;
  const int s = plan->slot[qq];
  if (s < 0) return;  // CTA-uniform
  const int a = crops.y1[qq] - crops.y0[qq] + 1, b = crops.x1[qq] - crops.x0[qq] + 1;
  const long long item = (long long)blk * OT_BLOCK + threadIdx.x;
  bool keep = false;
  int ii = 0, jj = 0;
  const double* px = nullptr;
  if (a > 0 && b > 0 && item < (long long)a * b) {
    ii = (int)(item / b), jj = (int)(item - (long long)ii * b);
    px = rep + ((size_t)(crops.y0[qq] + ii) * R + (size_t)(crops.x0[qq] + jj)) * C;
    double sum = 0.0;
    for (int c = 0; c < C; ++c) sum += fabs(px[c]);
    keep = sum > 0.0;
  }
  uint32_t total;
  const uint32_t r = block_rank(keep, warp_cnt, &total);
  if (!WRITE) {
    if (threadIdx.x == 0) blk_cnt[(size_t)s * nblk_q + blk] = total;
  } else if (keep) {
    double* row = Xt + ((size_t)s * cap + (size_t)(blk_off[(size_t)s * nblk_q + blk] + r)) * (C + 2);
    for (int c = 0; c < C; ++c) row[c] = px[c];
    row[C] = (double)ii / (double)(a - 1);      // np.arange(a) / (a - 1)
    row[C + 1] = (double)jj / (double)(b - 1);
  }
}

// exclusive scan of the per-block counts of the three slots (one CTA; a few hundred blocks at most per slot), totals out
__global__ void __launch_bounds__(256) k_otmi_scan(const uint32_t* __restrict__ cnt, int nblk, long long* __restrict__ off, long long* __restrict__ totals) {
  __shared__ long long part[256];
  for (int s = 0; s < 3; ++s) {
    const int per = (nblk + 255) / 256;
    const int b0 = threadIdx.x * per, b1 = min(nblk, b0 + per);
    long long sum = 0;
    for (int b = b0; b < b1; ++b) sum += cnt[(size_t)s * nblk + b];
    part[threadIdx.x] = sum;
    __syncthreads();
    long long run = 0;
    for (int k = 0; k < (int)threadIdx.x; ++k) run += part[k];
    for (int b = b0; b < b1; ++b) {
      off[(size_t)s * nblk + b] = run;
      run += cnt[(size_t)s * nblk + b];
    }
    if (threadIdx.x == 255) totals[s] = run;
    __syncthreads();
  }
}

struct OtmiWs {
  OtmiStats* st;
  OtmiPlan* plan;
  uint32_t *ecnt, *pcnt;
  long long *eoff, *poff, *totals;  // totals[0..2] events, [3..5] pixels
  size_t bytes;
};
OtmiWs otmi_carve(void* base, long long N, int R) {
  OtmiWs w;
  const size_t nblk_e = (size_t)((N + OT_BLOCK - 1) / OT_BLOCK) + 1;
  const size_t half = (size_t)R / 2 + 2;
  const size_t nblk_p = (half * half + OT_BLOCK - 1) / OT_BLOCK + 1;
  size_t o = 0;
  auto take = [&](size_t n) { const size_t at = o; o = align_up(o + n, 256); return at; };
  char* b = (char*)base;
  w.st = (OtmiStats*)(b + take(sizeof(OtmiStats) * 4));
  w.plan = (OtmiPlan*)(b + take(sizeof(OtmiPlan)));
  w.totals = (long long*)(b + take(sizeof(long long) * 8));
  w.ecnt = (uint32_t*)(b + take(sizeof(uint32_t) * 3 * nblk_e));
  w.eoff = (long long*)(b + take(sizeof(long long) * 3 * nblk_e));
  w.pcnt = (uint32_t*)(b + take(sizeof(uint32_t) * 3 * nblk_p));
  w.poff = (long long*)(b + take(sizeof(long long) * 3 * nblk_p));
  w.bytes = o;
  return w;
}

template <typename T>
int run_otmi_prepare(const T* ev, long long N, const double* rep, int R, int C, int height, int width, double* Xs, long long xs_cap, double* Xt,
                     long long xt_cap, long long* info_host, const OtmiWs& w, cudaStream_t stream) {
  // compute_otmi.py:97-107: the bounds are Python floats (true division)
  const double w2 = width / 2.0 - 1.0, h2 = height / 2.0 - 1.0, w1 = (double)width - 1.0, h1 = (double)height - 1.0;
  const int wdiv = (width - 1) / 2, hdiv = (height - 1) / 2;
  OtmiCrops cr;
  {  // compute_otmi.py:150-158 with int() of the float bounds (truncation)
    const double r = (double)R;
    const int lo0 = 0, lo1i = (int)(R / 2 - 1), hi0 = (int)(r / 2.0 - 1.0), hi1 = R - 1, lo1f = (int)(r / 2.0 - 1.0);
    // ([0, r // 2 - 1], [0, r / 2 - 1]), ([r / 2 - 1, r - 1], [0, r / 2 - 1]), ([0, r / 2 - 1], [r / 2 - 1, r - 1]), ([r / 2 - 1, r - 1], [r / 2 - 1, r - 1])
    cr.x0[0] = lo0, cr.x1[0] = lo1i, cr.y0[0] = lo0, cr.y1[0] = lo1f;
    cr.x0[1] = hi0, cr.x1[1] = hi1, cr.y0[1] = lo0, cr.y1[1] = lo1f;
    cr.x0[2] = lo0, cr.x1[2] = lo1f, cr.y0[2] = hi0, cr.y1[2] = hi1;
    cr.x0[3] = hi0, cr.x1[3] = hi1, cr.y0[3] = hi0, cr.y1[3] = hi1;
    for (int q = 0; q < 4; ++q) {  // numpy slicing clips at the array's edge
      cr.x0[q] = cr.x0[q] < 0 ? 0 : cr.x0[q], cr.y0[q] = cr.y0[q] < 0 ? 0 : cr.y0[q];
      cr.x1[q] = cr.x1[q] > R - 1 ? R - 1 : cr.x1[q], cr.y1[q] = cr.y1[q] > R - 1 ? R - 1 : cr.y1[q];
    }
  }
  const int nblk_e = (int)((N + OT_BLOCK - 1) / OT_BLOCK);
  long long max_px = 0;
  for (int q = 0; q < 4; ++q) {
    const long long a = cr.y1[q] - cr.y0[q] + 1, b = cr.x1[q] - cr.x0[q] + 1;
    if (a > 0 && b > 0 && a * b > max_px) max_px = a * b;
  }
  const int nblk_p = (int)((max_px + OT_BLOCK - 1) / OT_BLOCK);
  k_otmi_stats_init<<<1, 32, 0, stream>>>(w.st);
  if (N > 0) {
    const int grid = (int)std::min<long long>((N + 255) / 256, 1184);
    k_otmi_stats<T><<<grid, 256, 0, stream>>>(ev, N, w2, h2, w1, h1, w.st);
  }
  k_otmi_plan<T><<<1, 32, 0, stream>>>(ev, w.st, w.plan);
  EVREP_CUDA_OK(cudaMemsetAsync(w.totals, 0, sizeof(long long) * 8, stream));
  if (nblk_e > 0) {
    k_otmi_events<T, false><<<nblk_e, OT_BLOCK, 0, stream>>>(ev, N, w.plan, w2, h2, w1, h1, wdiv, hdiv, w.ecnt, w.eoff, Xs, xs_cap);
    k_otmi_scan<<<1, 256, 0, stream>>>(w.ecnt, nblk_e, w.eoff, w.totals);
    k_otmi_events<T, true><<<nblk_e, OT_BLOCK, 0, stream>>>(ev, N, w.plan, w2, h2, w1, h1, wdiv, hdiv, w.ecnt, w.eoff, Xs, xs_cap);
  }
  if (nblk_p > 0) {
    EVREP_CUDA_OK(cudaMemsetAsync(w.pcnt, 0, sizeof(uint32_t) * 3 * (size_t)nblk_p, stream));  // the dropped quadrant's CTAs write nothing
    k_otmi_pixels<false><<<4 * nblk_p, OT_BLOCK, 0, stream>>>(rep, R, C, w.plan, cr, nblk_p, w.pcnt, w.poff, Xt, xt_cap);
    k_otmi_scan<<<1, 256, 0, stream>>>(w.pcnt, nblk_p, w.poff, w.totals + 3);
    k_otmi_pixels<true><<<4 * nblk_p, OT_BLOCK, 0, stream>>>(rep, R, C, w.plan, cr, nblk_p, w.pcnt, w.poff, Xt, xt_cap);
  }
  EVREP_CUDA_OK(cudaGetLastError());
  // the caller slices the point sets by these counts: one small read-back, the call synchronises
  OtmiPlan hp;
  long long tot[8];
  EVREP_CUDA_OK(cudaMemcpyAsync(&hp, w.plan, sizeof(OtmiPlan), cudaMemcpyDeviceToHost, stream));
  EVREP_CUDA_OK(cudaMemcpyAsync(tot, w.totals, sizeof(tot), cudaMemcpyDeviceToHost, stream));
  EVREP_CUDA_OK(cudaStreamSynchronize(stream));
  for (int s = 0; s < 3; ++s) info_host[s] = tot[s], info_host[3 + s] = tot[3 + s];
  info_host[6] = hp.dropped;
  info_host[7] = hp.empty_quadrant;
  for (int q = 0; q < 4; ++q) info_host[8 + q] = hp.n_in[q];
  return EVREP_OK;
}

}  // namespace

size_t otmi_workspace_bytes(long long N, int R) { return otmi_carve(nullptr, N, R).bytes; }

int launch_otmi_prepare(const void* ev, int ev_type, long long N, const double* rep, int R, int C, int height, int width, double* Xs, long long xs_cap,
                        double* Xt, long long xt_cap, long long* info_host, void* workspace, cudaStream_t stream) {
  const OtmiWs w = otmi_carve(workspace, N, R);
  if (ev_type == 0)
    return run_otmi_prepare<int32_t>((const int32_t*)ev, N, rep, R, C, height, width, Xs, xs_cap, Xt, xt_cap, info_host, w, stream);
  if (ev_type == 1)
    return run_otmi_prepare<float>((const float*)ev, N, rep, R, C, height, width, Xs, xs_cap, Xt, xt_cap, info_host, w, stream);
  return run_otmi_prepare<double>((const double*)ev, N, rep, R, C, height, width, Xs, xs_cap, Xt, xt_cap, info_host, w, stream);
}

}  // namespace evrep
