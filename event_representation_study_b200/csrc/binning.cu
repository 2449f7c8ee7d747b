// Spatial pre-bucketing of event windows: count -> column scan -> scan -> scatter.
//
// Every tile-based representation starts here.  A window's pixels are cut into tiles of tile_px
// consecutive linear pixel indices (y*W + x); the events of each window are regrouped so that all
// events of one (window, tile) bucket are contiguous 8-byte records (two buckets per tile, p > 0 first,
// when the consumer asks for a polarity split).  The per-tile kernels then reduce a bucket entirely in
// shared memory and write their slice of the output exactly once.
//
// The placement is deterministic and needs no global atomics: a CTA owns one super-chunk of SUPER = 8192
// consecutive events.  k_hist writes that super-chunk's bucket counts as one row of `cc`; k_colscan writes, per
// column of `cc`, the exclusive prefix over the window's super-chunks into `cp` (the column total is the bucket
// size); k_scan turns bucket sizes into bucket starts; k_bin ranks its events inside its own counts with one
// shared-memory atomic each, stages the super-chunk sorted by bucket in shared memory and copies it out as
// runs (bucket start + prefix gives the run's place).  Records of different super-chunks therefore stay in
// stream order inside a bucket, and adjacent runs are written by CTAs that run at about the same time, so
// partially written sectors complete in L2 (scattering single records straight from registers was measured:
// 32 M sector writes, sectors evicted half full, 486 MB read + 387 MB written for 288 + 256 MB of payload).
//
// Replaces the per-channel boolean masking / np.concatenate / torch_scatter passes of the reference
// (representations/representation_search/mixed_density_event_stack.py:111-151, operations.py:39-89)
// and the np.put passes of event_stack.py:118-131.
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "evrep_common.cuh"

namespace evrep {

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8_u16(const uint16_t* __restrict__ p, int64_t g0, int64_t total, bool vec, uint32_t (&v)[EPT]) {
  if (vec && g0 + EPT <= total) {
    uint4 q = __ldg(reinterpret_cast<const uint4*>(p + g0));
    v[0] = q.x & 0xffffu; v[1] = q.x >> 16; v[2] = q.y & 0xffffu; v[3] = q.y >> 16;
    v[4] = q.z & 0xffffu; v[5] = q.z >> 16; v[6] = q.w & 0xffffu; v[7] = q.w >> 16;
  } else {
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = (g0 + e < total) ? (uint32_t)__ldg(p + g0 + e) : 0u;
  }
}

__device__ __forceinline__ void load8_i8(const int8_t* __restrict__ p, int64_t g0, int64_t total, bool vec, int (&v)[EPT]) {
  if (vec && g0 + EPT <= total) {
    uint2 q = __ldg(reinterpret_cast<const uint2*>(p + g0));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] = (int)(int8_t)((q.x >> (8 * e)) & 0xff);
      v[4 + e] = (int)(int8_t)((q.y >> (8 * e)) & 0xff);
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = (g0 + e < total) ? (int)__ldg(p + g0 + e) : 0;
  }
}

template <typename TT>
__device__ __forceinline__ void load8_t(const TT* __restrict__ p, int64_t g0, int64_t total, bool vec, int64_t (&v)[EPT]);

template <>
__device__ __forceinline__ void load8_t<int32_t>(const int32_t* __restrict__ p, int64_t g0, int64_t total, bool vec, int64_t (&v)[EPT]) {
  if (vec && g0 + EPT <= total) {
    int4 a = __ldg(reinterpret_cast<const int4*>(p + g0));
    int4 b = __ldg(reinterpret_cast<const int4*>(p + g0) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = (g0 + e < total) ? (int64_t)__ldg(p + g0 + e) : 0;
  }
}
template <>
__device__ __forceinline__ void load8_t<int64_t>(const int64_t* __restrict__ p, int64_t g0, int64_t total, bool vec, int64_t (&v)[EPT]) {
  if (vec && g0 + EPT <= total) {
    const longlong2* q = reinterpret_cast<const longlong2*>(p + g0);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      longlong2 a = __ldg(q + e);
      v[2 * e] = a.x; v[2 * e + 1] = a.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = (g0 + e < total) ? (int64_t)__ldg(p + g0 + e) : 0;
  }
}

// Packed forms of the same loads: the k_bin prefetch keeps a chunk's events in 18 (int32 t) / 26 (int64 t) registers
// while the bucket tables are being built, and unpacks them afterwards.
__device__ __forceinline__ uint4 load8_u16_raw(const uint16_t* __restrict__ p, int64_t g0, int64_t total, bool vec) {
  if (vec && g0 + EPT <= total) return __ldg(reinterpret_cast<const uint4*>(p + g0));
  uint32_t v[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) v[e] = (g0 + e < total) ? (uint32_t)__ldg(p + g0 + e) : 0u;
  return make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
}
__device__ __forceinline__ uint2 load8_i8_raw(const int8_t* __restrict__ p, int64_t g0, int64_t total, bool vec) {
  if (vec && g0 + EPT <= total) return __ldg(reinterpret_cast<const uint2*>(p + g0));
  uint32_t v[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) v[e] = (g0 + e < total) ? ((uint32_t)(int)__ldg(p + g0 + e) & 0xffu) : 0u;
  return make_uint2(v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24), v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24));
}
__device__ __forceinline__ uint32_t raw_u16(const uint4& q, int e) {
  const uint32_t w = (e >> 1) == 0 ? q.x : (e >> 1) == 1 ? q.y : (e >> 1) == 2 ? q.z : q.w;
  return (e & 1) ? (w >> 16) : (w & 0xffffu);
}
__device__ __forceinline__ int raw_i8(const uint2& q, int e) { return (int)(int8_t)(((e < 4 ? q.x : q.y) >> (8 * (e & 3))) & 0xffu); }

template <typename TT>
struct RawT;
template <>
struct RawT<int32_t> {
  int4 a, b;
  __device__ __forceinline__ void load(const int32_t* __restrict__ p, int64_t g0, int64_t total, bool vec) {
    if (vec && g0 + EPT <= total) {
      a = __ldg(reinterpret_cast<const int4*>(p + g0));
      b = __ldg(reinterpret_cast<const int4*>(p + g0) + 1);
    } else {
      int v[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) v[e] = (g0 + e < total) ? __ldg(p + g0 + e) : 0;
      a = make_int4(v[0], v[1], v[2], v[3]);
      b = make_int4(v[4], v[5], v[6], v[7]);
    }
  }
  __device__ __forceinline__ void load_vec(const int32_t* __restrict__ p, int64_t g0) {  // g0 % 8 == 0, all 8 in bounds
    a = __ldg(reinterpret_cast<const int4*>(p + g0));
    b = __ldg(reinterpret_cast<const int4*>(p + g0) + 1);
  }
  __device__ __forceinline__ int32_t get(int e) const {
    return e == 0 ? a.x : e == 1 ? a.y : e == 2 ? a.z : e == 3 ? a.w : e == 4 ? b.x : e == 5 ? b.y : e == 6 ? b.z : b.w;
  }
};
template <>
struct RawT<int64_t> {
  longlong2 q[4];
  __device__ __forceinline__ void load(const int64_t* __restrict__ p, int64_t g0, int64_t total, bool vec) {
    if (vec && g0 + EPT <= total) {
#pragma unroll
      for (int e = 0; e < 4; ++e) q[e] = __ldg(reinterpret_cast<const longlong2*>(p + g0) + e);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        q[e].x = (g0 + 2 * e < total) ? __ldg(p + g0 + 2 * e) : 0;
        q[e].y = (g0 + 2 * e + 1 < total) ? __ldg(p + g0 + 2 * e + 1) : 0;
      }
    }
  }
  __device__ __forceinline__ void load_vec(const int64_t* __restrict__ p, int64_t g0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) q[e] = __ldg(reinterpret_cast<const longlong2*>(p + g0) + e);
  }
  __device__ __forceinline__ int64_t get(int e) const { return (e & 1) ? q[e >> 1].y : q[e >> 1].x; }
};

// Shared-memory atomics on a 32-bit shared-window address (the generic-pointer form makes the compiler rebuild the
// window base around every atomic inside divergent code).
__device__ __forceinline__ void smem_inc(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory"); }
__device__ __forceinline__ uint32_t smem_fetch_inc(uint32_t addr) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
  return old;
}

// In-place exclusive scan of a[0..n) in shared memory by a BIN_THREADS-thread CTA (n <= MAX_TILES).
// `warp_tot` is BIN_THREADS/32 words of scratch.  Returns the total.  Ends with a __syncthreads().
template <int ITEMS>
__device__ __forceinline__ uint32_t block_exclusive_scan_n(uint32_t* a, int n, uint32_t* warp_tot) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + BIN_THREADS - 1) / BIN_THREADS;  // <= ITEMS
  const int i0 = tid * per;
  uint32_t loc[ITEMS];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    uint32_t v = (k < per && i0 + k < n) ? a[i0 + k] : 0u;
    loc[k] = sum;
    sum += v;
  }
  uint32_t inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < BIN_THREADS / 32) ? warp_tot[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    if (lane < BIN_THREADS / 32) warp_tot[lane] = winc - w;  // exclusive
    if (lane == BIN_THREADS / 32 - 1) warp_tot[BIN_THREADS / 32] = winc;
  }
  __syncthreads();
  const uint32_t basev = warp_tot[warp] + (inc - sum);
#pragma unroll
  for (int k = 0; k < ITEMS; ++k)
    if (k < per && i0 + k < n) a[i0 + k] = basev + loc[k];
  const uint32_t total = warp_tot[BIN_THREADS / 32];
  __syncthreads();
  return total;
}
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t* a, int n, uint32_t* warp_tot) {
  static_assert(MAX_TILES / BIN_THREADS == 8, "items per thread");
  if (n <= 2 * BIN_THREADS) return block_exclusive_scan_n<2>(a, n, warp_tot);  // CTA-uniform choice
  if (n <= 4 * BIN_THREADS) return block_exclusive_scan_n<4>(a, n, warp_tot);
  return block_exclusive_scan_n<8>(a, n, warp_tot);
}

// ---------------------------------------------------------------------------------------------
// per-window initialisation: one CTA per window; also labels the window's super-chunks
// ---------------------------------------------------------------------------------------------
// Offsets and super-chunk prefix of a small batch travel as a kernel argument (no host-to-device copy at all: nothing for
// the driver to stage, and the whole call can be captured into a CUDA graph); larger batches upload them.
constexpr int INIT_TABLE_MAX = 320;  // windows (+ 1): 12 bytes each, under the 4 KB kernel-parameter limit
struct InitTable {
  int64_t offsets[INIT_TABLE_MAX];
  int32_t sc_prefix[INIT_TABLE_MAX];
};

template <typename TT, bool BY_VALUE>
__global__ void k_init(const TT* __restrict__ t, int64_t* __restrict__ offsets, int32_t* __restrict__ sc_prefix, const __grid_constant__ InitTable tab,
                       int B, int32_t* __restrict__ sc_win, WinParams* __restrict__ wp, uint32_t* __restrict__ ticket) {
  const int b = blockIdx.x;
  pdl_wait();  // the previous call's kernels may still be reading the workspace
  pdl_trigger();
  if (b == 0 && threadIdx.x < 64) ticket[threadIdx.x] = 0;  // work counters of the persistent tile kernels
  int64_t o0, o1;
  int32_t s0, s1;
  if (BY_VALUE) {
    o0 = tab.offsets[b]; o1 = tab.offsets[b + 1];
    s0 = tab.sc_prefix[b]; s1 = tab.sc_prefix[b + 1];
    if (threadIdx.x == 0) {  // the tables also live in the workspace for the kernels that follow
      offsets[b] = o0; sc_prefix[b] = s0;
      if (b == B - 1) { offsets[B] = o1; sc_prefix[B] = s1; }
    }
  } else {
    o0 = offsets[b]; o1 = offsets[b + 1];
    s0 = sc_prefix[b]; s1 = sc_prefix[b + 1];
  }
  if (sc_win)
    for (int s = s0 + (int)threadIdx.x; s < s1; s += (int)blockDim.x) sc_win[s] = b;
  if (threadIdx.x != 0) return;
  WinParams w;
  w.start = o0;
  w.n = o1 - o0;
  w.flags = 0;
  w.has_m1 = 0;
  w.pad = 0;
  w.tmin_rel = INT_MAX;
  w.tmax_rel = INT_MIN;
  w.t_base = 0;
  w.tlast_rel = 0;
  if (w.n > 0) {
    w.t_base = (int64_t)t[w.start];
    int64_t d = (int64_t)t[w.start + w.n - 1] - w.t_base;
    if (d >= T_REL_LIMIT || d <= -T_REL_LIMIT) {
      w.flags |= EVREP_WF_T_RANGE;
      d = d > 0 ? T_REL_LIMIT - 1 : -(T_REL_LIMIT - 1);
    }
    w.tlast_rel = (int32_t)d;
  }
  wp[b] = w;
}

// Time-surface snapshot indices.  Either the caller's (B*S int64, already on the device) or the rule of
// representations/gen1_transforms.py:78-80: searchsorted((t - t0)/(tN - t0)*S, [1..S]) (left).
// Surface s is emitted only while the indices are strictly increasing and in range
// (time_surface.py:66-74: the `i == indices[s]` test can never fire again after a duplicate).
template <typename TT>
__global__ void k_snap_init(const TT* __restrict__ t, const WinParams* __restrict__ wp, const int64_t* __restrict__ user_idx,
                            int S, SnapParams* __restrict__ snap) {
  const int b = blockIdx.x;
  __shared__ int64_t idx[MAX_SNAP];
  const WinParams w = wp[b];
  const int s = threadIdx.x;
  if (s < S) {
    if (user_idx) {
      idx[s] = user_idx[(size_t)b * S + s];
    } else if (w.n > 0) {
      const double t0 = (double)((int64_t)t[w.start] - w.t_base);
      const double den = (double)((int64_t)t[w.start + w.n - 1] - w.t_base) - t0;
      const double target = (double)(s + 1);
      int64_t lo = 0, hi = w.n;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        double tn = ((double)((int64_t)t[w.start + mid] - w.t_base) - t0) / den * (double)S;
        if (tn < target) lo = mid + 1; else hi = mid;  // NaN compares false -> behaves as +inf, like numpy's sort order
      }
      idx[s] = lo;
    } else {
      idx[s] = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    SnapParams sp;
    for (int k = 0; k < MAX_SNAP; ++k) { sp.idx[k] = 0; sp.t_rel[k] = 0; }
    sp.pad[0] = sp.pad[1] = sp.pad[2] = 0;
    int nv = 0;
    int64_t prev = -1;
    for (int k = 0; k < S; ++k) {
      int64_t i = idx[k];
      if (i <= prev || i >= w.n) break;
      sp.idx[k] = (int32_t)i;
      int64_t d = (int64_t)t[w.start + i] - w.t_base;
      d = d >= T_REL_LIMIT ? T_REL_LIMIT - 1 : (d <= -T_REL_LIMIT ? -(T_REL_LIMIT - 1) : d);
      sp.t_rel[k] = (int32_t)d;
      prev = i;
      ++nv;
    }
    sp.n_valid = nv;
    snap[b] = sp;
  }
}

// ---------------------------------------------------------------------------------------------
// pass 1: bucket counts of one super-chunk -> one row of cc.  Counts every event with a valid pixel.
// HBM: reads x, y (+ p when the buckets are split by polarity).
// A thread's EPT events take the vector path (three 16-byte loads, no per-event bounds tests) whenever they all lie
// inside the window; only the first / last few threads of a window run the scalar path.
// ---------------------------------------------------------------------------------------------
template <bool SPLIT, bool DIV>
__global__ void __launch_bounds__(BIN_THREADS) k_hist(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                                                      const int8_t* __restrict__ p, const WinParams* __restrict__ wp,
                                                      const int32_t* __restrict__ sc_prefix, const int32_t* __restrict__ sc_win,
                                                      const Geom g, const bool vec, uint16_t* __restrict__ cc) {
  extern __shared__ uint32_t sh_hist[];
  pdl_wait();
  pdl_trigger();
  const int b = __ldg(sc_win + blockIdx.x);
  const int scl = blockIdx.x - __ldg(sc_prefix + b);
  const int64_t start = wp[b].start;
  const int n = (int)wp[b].n;
  const int64_t c0 = (start & ~(int64_t)(EPT - 1)) + (int64_t)scl * SUPER;  // first event of this CTA's super-chunk
  // the loads of both sub-chunks are issued before the histogram is cleared
  uint4 qx[SC_CHUNKS], qy[SC_CHUNKS];
  uint2 qp[SC_CHUNKS];
  bool full[SC_CHUNKS];
#pragma unroll
  for (int sub = 0; sub < SC_CHUNKS; ++sub) {
    const int64_t g0 = c0 + (int64_t)sub * CHUNK + (int64_t)threadIdx.x * EPT;
    const int idx0 = (int)(g0 - start);
    full[sub] = vec && idx0 >= 0 && idx0 + EPT <= n;
    if (full[sub]) {
      qx[sub] = __ldg(reinterpret_cast<const uint4*>(x + g0));
      qy[sub] = __ldg(reinterpret_cast<const uint4*>(y + g0));
      if (SPLIT) qp[sub] = __ldg(reinterpret_cast<const uint2*>(p + g0));
    }
  }
  for (int i = threadIdx.x; i < g.Tb; i += BIN_THREADS) sh_hist[i] = 0;
  __syncthreads();
  uint32_t hbase = (uint32_t)__cvta_generic_to_shared(sh_hist);
  asm volatile("" : "+r"(hbase));  // opaque: otherwise the window base is rebuilt around every atomic
  const uint32_t Wd = (uint32_t)g.W, Hd = (uint32_t)g.H;
  const int tile_shift = g.tile_shift;
  auto bucket = [&](uint32_t xe, uint32_t ye, int pe) {
    uint32_t bin = (ye * Wd + xe) >> tile_shift;
    if (SPLIT) bin = bin + bin + (pe > 0 ? 0u : 1u);
    return bin;
  };
#pragma unroll
  for (int sub = 0; sub < SC_CHUNKS; ++sub) {
    bool fast = full[sub];
    uint32_t xs[EPT], ys[EPT];
    if (fast) {  // all EPT pixels valid: the common case runs without a branch per event
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        xs[e] = raw_u16(qx[sub], e);
        ys[e] = raw_u16(qy[sub], e);
        if (DIV) {  // pixels are cells of div_x x div_y sensor pixels (the resize filter)
          xs[e] /= (uint32_t)g.div_x;
          ys[e] /= (uint32_t)g.div_y;
        }
        fast &= (xs[e] < Wd) & (ys[e] < Hd);
      }
    }
    if (fast) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) smem_inc(hbase + (bucket(xs[e], ys[e], SPLIT ? raw_i8(qp[sub], e) : 0) << 2));
    } else {
      const int64_t g0 = c0 + (int64_t)sub * CHUNK + (int64_t)threadIdx.x * EPT;
      const int idx0 = (int)(g0 - start);
      if (idx0 >= n || idx0 + EPT <= 0) continue;
#pragma unroll 1
      for (int e = 0; e < EPT; ++e) {
        if ((uint32_t)(idx0 + e) >= (uint32_t)n) continue;
        uint32_t xe = __ldg(x + g0 + e), ye = __ldg(y + g0 + e);
        if (DIV) {
          xe /= (uint32_t)g.div_x;
          ye /= (uint32_t)g.div_y;
        }
        if (xe < Wd && ye < Hd) smem_inc(hbase + (bucket(xe, ye, SPLIT ? (int)__ldg(p + g0 + e) : 0) << 2));
      }
    }
  }
  __syncthreads();
  // the row holds the exclusive scan of the counts (where each bucket's records start when the super-chunk is sorted
  // by bucket) followed by the total: k_bin needs exactly that, k_colscan recovers a count as a difference
  __shared__ uint32_t warp_tot[BIN_THREADS / 32 + 1];
  const uint32_t total = block_exclusive_scan(sh_hist, g.Tb, warp_tot);
  uint16_t* dst = cc + (size_t)blockIdx.x * cc_stride(g.Tb);
  for (int i = threadIdx.x; i < g.Tb; i += BIN_THREADS) dst[i] = (uint16_t)sh_hist[i];  // <= SUPER = 8192
  if (threadIdx.x == 0) dst[g.Tb] = (uint16_t)total;
}

// pass 2: per bucket (column of cc), exclusive prefix over the window's super-chunks; the total is the bucket size.
// A CTA owns 32 columns of one window; its 8 warps split the window's rows, so that a window of hundreds of
// super-chunks costs two short rounds of independent loads (row segment totals, then the prefix) instead of one long
// dependent walk.  A count is the difference of two neighbouring row entries: one load per lane plus a shuffle.
__global__ void __launch_bounds__(256) k_colscan(const uint16_t* __restrict__ cc, const int32_t* __restrict__ sc_prefix, int Tb,
                                                 uint32_t* __restrict__ cp, uint32_t* __restrict__ hist) {
  constexpr int U = 8;
  __shared__ uint32_t tot[8][32];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, c = c0 + lane;
  const int s0 = sc_prefix[b], s1 = sc_prefix[b + 1];
  const int per = (s1 - s0 + 7) / 8;
  const int r0 = min(s0 + seg * per, s1), r1 = min(r0 + per, s1);
  const bool ok = c < Tb;
  // counts of U rows at column c: every lane loads row[c] (the last lane / last column also row[c + 1]) for all U rows
  // first, so the loads are in flight together; the neighbour's value then comes from a shuffle
  const bool edge = ok && (lane == 31 || c + 1 == Tb);
  auto counts = [&](int r, uint32_t (&v)[U]) {
    uint32_t lo[U], ex[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint16_t* row = cc + (size_t)(r + k) * cc_stride(Tb);
      const bool live = ok && r + k < r1;
      lo[k] = live ? (uint32_t)row[c] : 0u;
      ex[k] = (live && edge) ? (uint32_t)row[c + 1] : 0u;
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t hi = __shfl_down_sync(0xffffffffu, lo[k], 1);
      v[k] = (edge ? ex[k] : hi) - lo[k];  // dead lanes / rows: 0 - 0
    }
  };
  uint32_t sum = 0;
  for (int r = r0; r < r1; r += U) {
    uint32_t v[U];
    counts(r, v);
#pragma unroll
    for (int k = 0; k < U; ++k) sum += v[k];
  }
  tot[seg][lane] = sum;
  __syncthreads();
  uint32_t run = 0, total = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t tk = tot[k][lane];
    if (k < seg) run += tk;
    total += tk;
  }
  for (int r = r0; r < r1; r += U) {
    uint32_t v[U];
    counts(r, v);
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (ok && r + k < r1) cp[(size_t)(r + k) * Tb + c] = run;
      run += v[k];
    }
  }
  if (ok && seg == 0) hist[(size_t)b * Tb + c] = total;
}

// bucket starts: one CTA per window, exclusive scan over its Tb buckets
__global__ void __launch_bounds__(BIN_THREADS) k_scan(const uint32_t* __restrict__ hist, uint32_t* __restrict__ base, int T) {
  __shared__ uint32_t a[MAX_TILES];
  __shared__ uint32_t warp_tot[BIN_THREADS / 32 + 1];
  pdl_wait();
  pdl_trigger();
  const size_t o = (size_t)blockIdx.x * T;
  for (int i = threadIdx.x; i < T; i += BIN_THREADS) a[i] = hist[o + i];
  __syncthreads();
  block_exclusive_scan(a, T, warp_tot);
  for (int i = threadIdx.x; i < T; i += BIN_THREADS) base[o + i] = a[i];
}

// ---------------------------------------------------------------------------------------------
// pass 3: scatter the events into their buckets as 8-byte records.  HBM: reads 9 B/event, writes 8 B/event.
//
// A thread owns EPT consecutive events per sub-chunk.  All its loads are issued before the bucket tables are
// read.  The EPT events take the FAST path when they lie inside the window, are time-sorted (also against the
// event before them), carry valid polarities and the first / last of them are inside the 31-bit time range
// (sorted => so are the ones between): then the per-event work is pixel -> bucket -> one returning shared-memory
// atomic -> one 8-byte + one 2-byte staged store.  Anything else (window edges, unsorted or out-of-range data)
// goes through the per-event checks of the slow path, which re-reads its events with scalar loads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool polarities_valid4(uint32_t w) {  // every byte in {0x00, 0x01, 0xff}
  const uint32_t a = w & 0xfefefefeu, nb = ~w;
  const uint32_t nza = (((a & 0x7f7f7f7fu) + 0x7f7f7f7fu) | a) & 0x80808080u;     // byte not in {0, 1}
  const uint32_t nzb = (((nb & 0x7f7f7f7fu) + 0x7f7f7f7fu) | nb) & 0x80808080u;  // byte != 0xff
  return (nza & nzb) == 0u;
}

#ifndef EVREP_BIN_CTAS
#define EVREP_BIN_CTAS 2
#endif

// what a CTA needs to know about one super-chunk
struct ChunkHdr {
  int b, n;        // window, its event count (< 2^31 - 8, checked on the host)
  int64_t start;   // absolute index of the window's first event
  int64_t t_base;  // timestamp of the window's first event
  int64_t c0;      // absolute index of the super-chunk's first event slot (a multiple of EPT)
  int32_t tlast_rel;
};
__device__ __forceinline__ ChunkHdr chunk_hdr(int sc, const WinParams* __restrict__ wp, const int32_t* __restrict__ sc_prefix,
                                              const int32_t* __restrict__ sc_win) {
  ChunkHdr h;
  h.b = __ldg(sc_win + sc);
  const int scl = sc - __ldg(sc_prefix + h.b);
  h.start = wp[h.b].start;
  h.n = (int)wp[h.b].n;
  h.t_base = wp[h.b].t_base;
  h.tlast_rel = wp[h.b].tlast_rel;
  h.c0 = (h.start & ~(int64_t)(EPT - 1)) + (int64_t)scl * SUPER;
  return h;
}

// one thread's events of one super-chunk, as loaded (packed)
template <typename TT>
struct ChunkRegs {
  uint4 qx[SC_CHUNKS], qy[SC_CHUNKS];
  uint2 qp[SC_CHUNKS];
  RawT<TT> qt[SC_CHUNKS];
  TT t_before[SC_CHUNKS];
  bool full[SC_CHUNKS];
};
// issues the loads of the thread's events of one sub-chunk; a sub-chunk that is not entirely inside the window (or unaligned
// arrays) is left to the slow path of chunk_rank, which reads its events with scalar loads
template <typename TT>
__device__ __forceinline__ void chunk_fetch_sub(ChunkRegs<TT>& r, const int sub, const ChunkHdr& h, const uint16_t* __restrict__ x,
                                                const uint16_t* __restrict__ y, const TT* __restrict__ t, const int8_t* __restrict__ p, bool vec,
                                                int tid) {
  const int64_t g0 = h.c0 + (int64_t)sub * CHUNK + (int64_t)tid * EPT;
  const int idx0 = (int)(g0 - h.start);  // may be < 0 at the head of the window
  r.full[sub] = vec && idx0 >= 0 && idx0 + EPT <= h.n;
  if (r.full[sub]) {
    r.qx[sub] = __ldg(reinterpret_cast<const uint4*>(x + g0));
    r.qy[sub] = __ldg(reinterpret_cast<const uint4*>(y + g0));
    r.qp[sub] = __ldg(reinterpret_cast<const uint2*>(p + g0));
    if (sizeof(TT) == 4 || sub == 0) r.qt[sub].load_vec(t, g0);  // 64-bit timestamps of later sub-chunks: fetched when needed (registers)
    r.t_before[sub] = __ldg(t + g0 - (idx0 >= 1 ? 1 : 0));       // the window's first event is compared with itself
  }
}
template <typename TT>
__device__ __forceinline__ void chunk_fetch(ChunkRegs<TT>& r, const ChunkHdr& h, const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                                            const TT* __restrict__ t, const int8_t* __restrict__ p, bool vec, int tid) {
#pragma unroll
  for (int sub = 0; sub < SC_CHUNKS; ++sub) chunk_fetch_sub<TT>(r, sub, h, x, y, t, p, vec, tid);
}

struct BinAcc {  // per-thread partial results of the window scalars
  int tmin = INT_MAX, tmax = INT_MIN;
  uint32_t flags = 0, m1 = 0;
};
struct BinSmem {  // 32-bit shared-window addresses (opaque to the compiler: no rebuild of the window base per access)
  uint32_t stage, sbkt, lcur;
};

// ranks the thread's events of one super-chunk inside their buckets and stages their records
template <typename TT, int MODE, bool SPLIT, bool DIV>
__device__ __forceinline__ void chunk_rank(ChunkRegs<TT>& r, const ChunkHdr& h, const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                                           const TT* __restrict__ t, const int8_t* __restrict__ p, const Geom& g, const BinSmem sm,
                                           const int32_t* sh_snap_idx, const int ns, BinAcc& acc, int tid) {
  const int n = h.n;
  const int64_t start = h.start, t_base = h.t_base;
  const int32_t tlast_rel = h.tlast_rel;
  // index boundaries of the SBN windows (mixed_density_event_stack.py:55-74)
  const int n3 = n / 3, s4 = n / 2, s5 = s4 + n / 4, s6 = s5 + n / 8;
  const uint32_t Wd = (uint32_t)g.W, Hd = (uint32_t)g.H;
  const uint32_t pix_mask = (uint32_t)(g.tile_px - 1);
  const int tile_shift = g.tile_shift;

  // SBN window mask of an index / first time-surface snapshot an index feeds; both change monotonically with the
  // index, so a thread's EPT consecutive events nearly always share one value
  auto aux_of = [&](int idx) -> uint32_t {
    if (MODE == REC_T_WMASK)
      return 1u | (idx < n3 ? 2u : (idx < 2 * n3 ? 4u : (idx < 3 * n3 ? 8u : 0u))) | (idx >= s4 ? 16u : 0u) | (idx >= s5 ? 32u : 0u) |
             (idx >= s6 ? 64u : 0u);
    if (MODE == REC_T_SNAP) {
      int s = 0;
      while (s < ns && idx > sh_snap_idx[s]) ++s;
      return (uint32_t)s;
    }
    return 0u;
  };
  // one event with a valid pixel and polarity pv in {-1, 0, 1}: rank it inside its bucket and stage its record.
  // `keep` = false leaves a null record in the slot (the event was counted by k_hist).
  auto place = [&](uint32_t lin, int pv, int32_t t_rel, int idx, uint32_t aux, bool keep) {
    uint32_t bin = lin >> tile_shift;
    if (SPLIT) bin = bin + bin + (pv > 0 ? 0u : 1u);
    const uint32_t slot = smem_fetch_inc(sm.lcur + (bin << 2));
    uint32_t k = (uint32_t)t_rel;
    if (MODE == REC_IDX) k = (uint32_t)idx;
    if (MODE == REC_T_SNAP && aux >= (uint32_t)ns) keep = false;  // after the last emitted surface: feeds nothing
    if (MODE == REC_T_TORE && t_rel >= tlast_rel) keep = false;   // strict `<` against the sample time (tore.py:17)
    uint32_t meta;
    if (MODE == REC_T_IDX)
      meta = fused_meta(lin & pix_mask, (uint32_t)idx, (uint32_t)pv & 3u);
    else
      meta = rec_meta(lin & pix_mask, aux, (uint32_t)pv & 3u);
    if (keep) {
      acc.tmin = min(acc.tmin, t_rel);
      acc.tmax = max(acc.tmax, t_rel);
      if (MODE == REC_T_WMASK) acc.m1 |= pv == -1 ? aux : 0u;
    } else {
      k = 0u;
      meta = MODE == REC_T_IDX ? FUSED_NULL_META : REC_NULL_META;
    }
    if (SPLIT) {
      // split buckets come with 1024-pixel tiles and at most 4096 buckets: the staged record carries its own bucket in the meta
      // bits the pixel (10 of 16) and the top of the word leave free, and the copy-out strips them again - one random 2-byte
      // store and one load per event less on the shared-memory pipe that bounds this kernel
      meta |= ((bin & 63u) << 10) | ((bin >> 6) << 26);
      asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sm.stage + (slot << 3)), "r"(k), "r"(meta) : "memory");
    } else {
      asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sm.stage + (slot << 3)), "r"(k), "r"(meta) : "memory");
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(sm.sbkt + (slot << 1)), "h"((uint16_t)bin) : "memory");
    }
  };

#pragma unroll
  for (int sub = 0; sub < SC_CHUNKS; ++sub) {
    const int64_t g0 = h.c0 + (int64_t)sub * CHUNK + (int64_t)tid * EPT;
    const int idx0 = (int)(g0 - start);
    bool fast = r.full[sub];
    if (fast) {
      if (sizeof(TT) == 8 && sub > 0) r.qt[sub].load_vec(t, g0);
      // sorted (also against the previous event), valid polarities, first and last inside the 31-bit range
      TT prev = r.t_before[sub];
      bool sorted = true;
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const TT te = r.qt[sub].get(e);
        sorted &= !(te < prev);
        prev = te;
      }
      const int64_t d0 = (int64_t)r.qt[sub].get(0) - t_base, d7 = (int64_t)r.qt[sub].get(EPT - 1) - t_base;
      const bool in_range = d0 > -(int64_t)T_REL_LIMIT && d7 < (int64_t)T_REL_LIMIT;
      fast = sorted && in_range && polarities_valid4(r.qp[sub].x) && polarities_valid4(r.qp[sub].y);
    }
    const uint32_t aux_first = aux_of(idx0);
    if ((MODE == REC_T_WMASK || MODE == REC_T_SNAP) && aux_first != aux_of(idx0 + EPT - 1)) fast = false;  // a window boundary inside
    uint32_t lin[EPT];
    if (fast) {  // every pixel valid (an invalid one has no slot: k_hist did not count it)
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        uint32_t xe = raw_u16(r.qx[sub], e), ye = raw_u16(r.qy[sub], e);
        if (DIV) {
          xe /= (uint32_t)g.div_x;
          ye /= (uint32_t)g.div_y;
        }
        fast &= (xe < Wd) & (ye < Hd);
        lin[e] = ye * Wd + xe;
      }
    }
    if (fast) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int32_t t_rel = sizeof(TT) == 4 ? (int32_t)((uint32_t)r.qt[sub].get(e) - (uint32_t)t_base) : (int32_t)((int64_t)r.qt[sub].get(e) - t_base);
        place(lin[e], raw_i8(r.qp[sub], e), t_rel, idx0 + e, aux_first, true);
      }
    } else {
      if (idx0 >= n || idx0 + EPT <= 0) continue;
      TT t_prev = 0;
      bool have_prev = idx0 >= 1;
      if (have_prev) t_prev = __ldg(t + g0 - 1);
#pragma unroll 1
      for (int e = 0; e < EPT; ++e) {
        const int idx = idx0 + e;
        if ((uint32_t)idx >= (uint32_t)n) continue;
        const TT te = __ldg(t + g0 + e);
        uint32_t xe = __ldg(x + g0 + e), ye = __ldg(y + g0 + e);
        if (DIV) {
          xe /= (uint32_t)g.div_x;
          ye /= (uint32_t)g.div_y;
        }
        int pv = __ldg(p + g0 + e);
        if (have_prev && te < t_prev) acc.flags |= EVREP_WF_UNSORTED;
        t_prev = te;
        have_prev = true;
        if ((xe >= Wd) | (ye >= Hd)) { acc.flags |= EVREP_WF_OUT_OF_RANGE; continue; }  // not counted by k_hist: no slot
        bool keep = true;
        const int64_t d = (int64_t)te - t_base;
        if (d >= T_REL_LIMIT || d <= -T_REL_LIMIT) { acc.flags |= EVREP_WF_T_RANGE; keep = false; }
        // index-keyed records (EventStack, the filters) carry no timestamp: a stream longer than 2^30 us only raises the flag
        if (MODE == REC_IDX) keep = true;
        if (pv > 1 || pv < -1) { acc.flags |= EVREP_WF_BAD_POLARITY; pv = pv > 0 ? 1 : -1; }
        place(ye * Wd + xe, pv, (int32_t)d, idx, aux_of(idx), keep);
      }
    }
  }
}

// CTA-wide reduction of the per-window scalars into wp[b] (ends with the threads past a __syncthreads)
__device__ __forceinline__ void bin_publish(BinAcc acc, WinParams* __restrict__ wpb, int* sh_tmin, int* sh_tmax, uint32_t* sh_flags, uint32_t* sh_m1,
                                            int tid) {
  acc.tmin = __reduce_min_sync(0xffffffffu, acc.tmin);
  acc.tmax = __reduce_max_sync(0xffffffffu, acc.tmax);
  acc.flags = __reduce_or_sync(0xffffffffu, acc.flags);
  acc.m1 = __reduce_or_sync(0xffffffffu, acc.m1);
  if ((tid & 31) == 0) {
    if (acc.tmin != INT_MAX) { atomicMin(sh_tmin, acc.tmin); atomicMax(sh_tmax, acc.tmax); }
    if (acc.flags) atomicOr(sh_flags, acc.flags);
    if (acc.m1) atomicOr(sh_m1, acc.m1);
  }
  __syncthreads();
  if (tid == 0) {
    if (*sh_tmin != INT_MAX) { atomicMin(&wpb->tmin_rel, *sh_tmin); atomicMax(&wpb->tmax_rel, *sh_tmax); }
    if (*sh_flags) atomicOr(&wpb->flags, *sh_flags);
    if (*sh_m1) atomicOr(&wpb->has_m1, *sh_m1);
  }
}

// one CTA per super-chunk.  (A persistent, software-pipelined variant - next super-chunk's events in registers and its table
// rows on their way by cp.async during the copy-out - was measured slower, 0.212 against 0.169 ms: the kernel is bound by the
// shared-memory pipe (l1tex 70 % busy), not by exposed global-load latency, and the second CTA of the SM already fills
// the head of the first; profiles/README.md.)
template <typename TT, int MODE, bool SPLIT, bool DIV>
__global__ void __launch_bounds__(BIN_THREADS, EVREP_BIN_CTAS) k_bin(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                                                     const TT* __restrict__ t, const int8_t* __restrict__ p,
                                                     WinParams* __restrict__ wp, const SnapParams* __restrict__ snap,
                                                     const int32_t* __restrict__ sc_prefix, const int32_t* __restrict__ sc_win,
                                                     const Geom g, const bool vec, const uint32_t* __restrict__ base,
                                                     const uint16_t* __restrict__ cc, const uint32_t* __restrict__ cp,
                                                     uint2* __restrict__ records) {
  extern __shared__ __align__(16) unsigned char sh_raw[];
  uint2* stage = reinterpret_cast<uint2*>(sh_raw);                 // SUPER records, sorted by bucket
  uint16_t* sbkt = reinterpret_cast<uint16_t*>(stage + SUPER);     // SUPER: the bucket of every staged record
  uint32_t* lcur = reinterpret_cast<uint32_t*>(sbkt + SUPER);      // Tb: next free staged slot of every bucket
  uint32_t* delta = lcur + g.Tb;                                   // Tb: destination minus staged slot
  __shared__ int sh_tmin, sh_tmax;
  __shared__ uint32_t sh_flags, sh_m1;
  __shared__ int32_t sh_snap_idx[MAX_SNAP];
  __shared__ int sh_nsnap;

  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  const ChunkHdr h = chunk_hdr(blockIdx.x, wp, sc_prefix, sc_win);
  ChunkRegs<TT> r;
  chunk_fetch<TT>(r, h, x, y, t, p, vec, tid);  // every event of the thread is requested before the bucket tables are built

  const uint16_t* crow = cc + (size_t)blockIdx.x * cc_stride(g.Tb);  // first staged slot of every bucket (k_hist), then the total
  const uint32_t* brow = base + (size_t)h.b * g.Tb;
  const uint32_t* prow = cp + (size_t)blockIdx.x * g.Tb;
  for (int i = tid; i < g.Tb; i += BIN_THREADS) {
    const uint32_t lo = (uint32_t)__ldg(crow + i);
    lcur[i] = lo;
    delta[i] = __ldg(brow + i) + __ldg(prow + i) - lo;
  }
  const uint32_t total = (uint32_t)__ldg(crow + g.Tb);
  if (tid == 0) { sh_tmin = INT_MAX; sh_tmax = INT_MIN; sh_flags = 0; sh_m1 = 0; sh_nsnap = 0; }
  if (MODE == REC_T_SNAP) {
    if (tid < MAX_SNAP) sh_snap_idx[tid] = snap[h.b].idx[tid];
    if (tid == 0) sh_nsnap = snap[h.b].n_valid;
  }
  __syncthreads();

  BinSmem sm;
  sm.stage = (uint32_t)__cvta_generic_to_shared(stage);
  sm.sbkt = (uint32_t)__cvta_generic_to_shared(sbkt);
  sm.lcur = (uint32_t)__cvta_generic_to_shared(lcur);
  asm volatile("" : "+r"(sm.stage), "+r"(sm.sbkt), "+r"(sm.lcur));
  BinAcc acc;
  chunk_rank<TT, MODE, SPLIT, DIV>(r, h, x, y, t, p, g, sm, sh_snap_idx, MODE == REC_T_SNAP ? sh_nsnap : 0, acc, tid);
  bin_publish(acc, wp + h.b, &sh_tmin, &sh_tmax, &sh_flags, &sh_m1, tid);
  // copy out: consecutive threads hold consecutive records of a run
  uint2* dst = records + h.start;
  if (SPLIT) {
#pragma unroll 4
    for (uint32_t i = tid; i < total; i += BIN_THREADS) {
      uint2 rec = stage[i];
      const uint32_t bin = ((rec.y >> 10) & 63u) | ((rec.y >> 26) << 6);
      rec.y &= 0x03ff03ffu;
      dst[i + delta[bin]] = rec;
    }
  } else {
#pragma unroll 4
    for (uint32_t i = tid; i < total; i += BIN_THREADS) dst[i + delta[sbkt[i]]] = stage[i];
  }
}

// A single-pass alternative was measured and dropped (profiles/README.md, round 2: "k_sortbin"): one kernel that sorts a
// super-chunk by bucket on chip and writes the records back in place (0.152 ms for the 32 x 1 M step against 0.244 ms for
// k_hist + scans + k_bin).  Its records stay fragmented per super-chunk (about 2.3 records per bucket and super-chunk at
// 1 Mpx), so the tile kernel would have to gather ~120 18-byte pieces per bucket; the estimated net gain (-0.05 ms) did not
// justify a second record layout through every consumer.

template <typename TT, int MODE, bool SPLIT, bool DIV = false>
static int launch_bin(const Events& ev, const Geom& g, const Workspace& ws, int n_sc, bool vec, cudaStream_t stream) {
  const size_t smem = (size_t)SUPER * (sizeof(uint2) + sizeof(uint16_t)) + 2 * sizeof(uint32_t) * (size_t)g.Tb;
  EVREP_CUDA_OK(cudaFuncSetAttribute(k_bin<TT, MODE, SPLIT, DIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EVREP_CUDA_OK(launch_pdl(k_bin<TT, MODE, SPLIT, DIV>, n_sc, BIN_THREADS, smem, stream, ev.x, ev.y, (const TT*)ev.t, ev.p, ws.wp, ws.snap, ws.sc_prefix,
                           ws.sc_win, g, vec, ws.base, ws.cc, ws.cp, ws.records));
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}
template <typename TT>
static int launch_bin_mode(int mode, const Events& ev, const Geom& g, const Workspace& ws, int n_sc, bool vec, cudaStream_t stream) {
  switch (mode) {
    case REC_T_WMASK:
      return g.split ? launch_bin<TT, REC_T_WMASK, true>(ev, g, ws, n_sc, vec, stream) : launch_bin<TT, REC_T_WMASK, false>(ev, g, ws, n_sc, vec, stream);
    case REC_IDX:
      return (g.div_x > 1 || g.div_y > 1) ? launch_bin<TT, REC_IDX, false, true>(ev, g, ws, n_sc, vec, stream)
                                          : launch_bin<TT, REC_IDX, false>(ev, g, ws, n_sc, vec, stream);
    case REC_T_SNAP: return launch_bin<TT, REC_T_SNAP, false>(ev, g, ws, n_sc, vec, stream);
    case REC_T_TORE: return launch_bin<TT, REC_T_TORE, false>(ev, g, ws, n_sc, vec, stream);
    case REC_T_IDX: return launch_bin<TT, REC_T_IDX, false>(ev, g, ws, n_sc, vec, stream);
    default:
      return g.split ? launch_bin<TT, REC_T_ONLY, true>(ev, g, ws, n_sc, vec, stream) : launch_bin<TT, REC_T_ONLY, false>(ev, g, ws, n_sc, vec, stream);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int build_sc_prefix(const int64_t* win_offsets_host, int B, std::vector<int32_t>& prefix) {
  prefix.assign((size_t)B + 1, 0);
  int64_t acc = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t s = win_offsets_host[b], e = win_offsets_host[b + 1];
    if (e < s || s < 0) {
      set_error("win_offsets must be non-decreasing and non-negative (window %d: %lld..%lld)", b, (long long)s, (long long)e);
      return EVREP_EINVAL;
    }
    if (e - s >= (int64_t)INT_MAX - 8) {
      set_error("window %d holds %lld events; the limit is 2^31 - 9", b, (long long)(e - s));
      return EVREP_EUNSUPPORTED;
    }
    const int64_t a = s & ~(int64_t)(EPT - 1);
    acc += (e > s) ? (e - a + SUPER - 1) / SUPER : 0;
    if (acc > INT_MAX) {
      set_error("batch too large: more than 2^31 super-chunks");
      return EVREP_EUNSUPPORTED;
    }
    prefix[(size_t)b + 1] = (int32_t)acc;
  }
  return EVREP_OK;
}

// uploads offsets + super-chunk prefix, initialises the per-window parameters.  Shared by the tile pipeline
// and the direct-scatter ops (which pass g.Tb == 0: no super-chunk table).  Returns the number of super-chunks.
int prepare_windows(const Events& ev, const int64_t* win_offsets_host, const Geom& g, const Workspace& ws, int* n_super_chunks,
                    cudaStream_t stream) {
  std::vector<int32_t> prefix;
  int rc = build_sc_prefix(win_offsets_host, g.B, prefix);
  if (rc) return rc;
  if (win_offsets_host[g.B] > g.total) {
    set_error("win_offsets[B] = %lld exceeds total_events = %lld", (long long)win_offsets_host[g.B], (long long)g.total);
    return EVREP_EINVAL;
  }
  *n_super_chunks = prefix[(size_t)g.B];
  if ((int64_t)*n_super_chunks > max_super_chunks(g.B, g.total)) {  // cannot happen: ceil((n + 7) / SUPER) <= n / SUPER + 2
    set_error("internal: super-chunk bound exceeded");
    return EVREP_EINVAL;
  }
  int32_t* sc_win = g.Tb > 0 ? ws.sc_win : nullptr;
  if (g.B + 1 <= INIT_TABLE_MAX) {
    InitTable tab;
    memcpy(tab.offsets, win_offsets_host, sizeof(int64_t) * (size_t)(g.B + 1));
    memcpy(tab.sc_prefix, prefix.data(), sizeof(int32_t) * (size_t)(g.B + 1));
    if (ev.t_bytes == 4)
      EVREP_CUDA_OK(launch_pdl(k_init<int32_t, true>, g.B, 64, 0, stream, (const int32_t*)ev.t, ws.offsets, ws.sc_prefix, tab, g.B, sc_win, ws.wp, ws.ticket));
    else
      EVREP_CUDA_OK(launch_pdl(k_init<int64_t, true>, g.B, 64, 0, stream, (const int64_t*)ev.t, ws.offsets, ws.sc_prefix, tab, g.B, sc_win, ws.wp, ws.ticket));
  } else {
    // pageable-source async copies are staged before the call returns, so the host vectors may die here
    // one copy for both tables (each pageable-source copy costs several microseconds of staging on the host)
    std::vector<int64_t> pack((size_t)(g.B + 1) + ((size_t)(g.B + 1) + 1) / 2);
    memcpy(pack.data(), win_offsets_host, sizeof(int64_t) * (size_t)(g.B + 1));
    memcpy(pack.data() + (g.B + 1), prefix.data(), sizeof(int32_t) * (size_t)(g.B + 1));
    EVREP_CUDA_OK(cudaMemcpyAsync(ws.offsets, pack.data(), sizeof(int64_t) * (size_t)(g.B + 1) + sizeof(int32_t) * (size_t)(g.B + 1),
                                  cudaMemcpyHostToDevice, stream));
    InitTable tab;  // unused
    tab.offsets[0] = 0;
    if (ev.t_bytes == 4)
      k_init<int32_t, false><<<g.B, 64, 0, stream>>>((const int32_t*)ev.t, ws.offsets, ws.sc_prefix, tab, g.B, sc_win, ws.wp, ws.ticket);
    else
      k_init<int64_t, false><<<g.B, 64, 0, stream>>>((const int64_t*)ev.t, ws.offsets, ws.sc_prefix, tab, g.B, sc_win, ws.wp, ws.ticket);
  }
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("EVREP_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

bool events_vectorisable(const Events& ev) { return aligned16(ev.x) && aligned16(ev.y) && aligned16(ev.t) && aligned16(ev.p); }

int run_binning(const Events& ev, const int64_t* win_offsets_host, const Geom& g, const Workspace& ws, int rec_mode, int n_snap,
                const int64_t* snap_indices_host, cudaStream_t stream) {
  int n_sc = 0;
  prof_next_call();
  if (g.Tb < 1 || g.Tb > MAX_TILES) {
    set_error("internal: %d buckets per window", g.Tb);
    return EVREP_EINVAL;
  }
  int rc = prepare_windows(ev, win_offsets_host, g, ws, &n_sc, stream);
  if (rc) return rc;
  const bool vec = events_vectorisable(ev);  // (k_init zeroed the ticket counters)

  if (rec_mode == REC_T_SNAP || rec_mode == REC_T_IDX) {
    const int64_t* user = nullptr;
    if (snap_indices_host) {
      EVREP_CUDA_OK(cudaMemcpyAsync(ws.snap_in, snap_indices_host, sizeof(int64_t) * (size_t)g.B * n_snap, cudaMemcpyHostToDevice, stream));
      user = ws.snap_in;
    }
    if (ev.t_bytes == 4)
      k_snap_init<int32_t><<<g.B, 32, 0, stream>>>((const int32_t*)ev.t, ws.wp, user, n_snap, ws.snap);
    else
      k_snap_init<int64_t><<<g.B, 32, 0, stream>>>((const int64_t*)ev.t, ws.wp, user, n_snap, ws.snap);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  if (n_sc > 0) {
    prof_begin(EVREP_K_COUNT, stream);
    if (g.split)
      EVREP_CUDA_OK(launch_pdl(k_hist<true, false>, n_sc, BIN_THREADS, sizeof(uint32_t) * (size_t)g.Tb, stream, ev.x, ev.y, ev.p, ws.wp, ws.sc_prefix, ws.sc_win, g, vec, ws.cc));
    else if (g.div_x > 1 || g.div_y > 1)
      EVREP_CUDA_OK(launch_pdl(k_hist<false, true>, n_sc, BIN_THREADS, sizeof(uint32_t) * (size_t)g.Tb, stream, ev.x, ev.y, ev.p, ws.wp, ws.sc_prefix, ws.sc_win, g, vec, ws.cc));
    else
      EVREP_CUDA_OK(launch_pdl(k_hist<false, false>, n_sc, BIN_THREADS, sizeof(uint32_t) * (size_t)g.Tb, stream, ev.x, ev.y, ev.p, ws.wp, ws.sc_prefix, ws.sc_win, g, vec, ws.cc));
    prof_end(EVREP_K_COUNT, stream);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  prof_begin(EVREP_K_SCAN, stream);
  // bucket sizes (also for windows without events: an empty column range gives 0), then bucket starts
  for (int b0 = 0; b0 < g.B; b0 += 65535) {
    const int nb = std::min(65535, g.B - b0);
    EVREP_CUDA_OK(launch_pdl(k_colscan, dim3((unsigned)((g.Tb + 31) / 32), (unsigned)nb), 256, 0, stream, ws.cc, ws.sc_prefix + b0, g.Tb, ws.cp + 0, ws.hist + (size_t)b0 * g.Tb));
  }
  EVREP_CUDA_OK(launch_pdl(k_scan, g.B, BIN_THREADS, 0, stream, ws.hist, ws.base, g.Tb));
  prof_end(EVREP_K_SCAN, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  if (n_sc > 0) {
    prof_begin(EVREP_K_BIN, stream);
    const int rc2 = ev.t_bytes == 4 ? launch_bin_mode<int32_t>(rec_mode, ev, g, ws, n_sc, vec, stream)
                                    : launch_bin_mode<int64_t>(rec_mode, ev, g, ws, n_sc, vec, stream);
    if (rc2) return rc2;
    prof_end(EVREP_K_BIN, stream);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  return EVREP_OK;
}

}  // namespace evrep
