// ev-licious' stateful per-pixel event filters, which the reference runs as sequential numba
// loops over the whole stream (ev-licious/src/evlicious/tools/utils.py):
//   _refractory_period          :193-200  keep an event iff t - last_kept[y, x] >= period, then last_kept = t
//   _contrast_threshold_control :184-191  activity[y, x] += p; keep and reset when |activity| >= factor
//   _filter_events_resize       :143-158  per fx x fy cell: change += p / (fx fy); keep when |change| >= 1, then change -= p
// Each is a state machine per pixel (or cell) over that pixel's events IN STREAM ORDER; pixels are independent.
//   _background_activity_filter :169-178  reads timestamps[y, x], then writes t into the (2 radius)^2 block
//                                         [y - radius, y + radius) x [x - radius, x + radius).  Pixels are NOT independent
//                                         there, but every pixel's own sequence of reads and writes is: the stream is
//                                         expanded into one record per written pixel (k_ba_expand; the event's own pixel
//                                         first, flagged "read, then write"; writes clipped by the sensor border become
//                                         repeated writes of the own pixel) and the expanded stream runs through the same
//                                         per-pixel machine with state = the last timestamp written.
//
// On the GPU the events are bucketed by tile with key = stream index (binning.cu, REC_IDX); a CTA owns one (window,
// tile) bucket, takes it in segments made of whole super-chunk runs (runs of different super-chunks are in stream
// order, the records inside one run are not), sorts each segment by (pixel, stream index) with a block radix sort in
// shared memory (cub::BlockRadixSort inside this kernel), and then one thread per pixel walks that pixel's events in
// order with the state in a register.  State arrays come from and go back to the caller, so successive calls continue a
// stream exactly like successive `insert` calls of the reference's filter objects (tools/filters.py:57-109).
#include <algorithm>

#include <cub/block/block_radix_sort.cuh>

#include "evrep_common.cuh"

namespace evrep {

constexpr int FT_THREADS = 512;
constexpr int FT_ITEMS = 16;
constexpr int FT_ITEMS_SMALL = 4;
constexpr int FT_CAP = FT_THREADS * FT_ITEMS;  // 8192 records per segment = the longest possible run of one super-chunk
constexpr int FT_MAX_SC = 4096;                // super-chunks per window whose run table fits shared memory (33 M events)
static_assert(FT_CAP >= SUPER, "a segment must hold a whole super-chunk run");

template <int FILTER>
struct FilterState;
template <>
struct FilterState<EVREP_FILTER_REFRACTORY> { using type = double; };
template <>
struct FilterState<EVREP_FILTER_CONTRAST> { using type = int32_t; };
template <>
struct FilterState<EVREP_FILTER_RESIZE> { using type = float; };
template <>
struct FilterState<EVREP_FILTER_BACKGROUND> { using type = double; };

template <int FILTER, typename TT>
__global__ void __launch_bounds__(FT_THREADS) k_filter_tile(const uint2* __restrict__ records, const uint32_t* __restrict__ base,
                                                            const uint32_t* __restrict__ hist, const uint32_t* __restrict__ cp,
                                                            const int32_t* __restrict__ sc_prefix, const WinParams* __restrict__ wp,
                                                            const Geom g, const TT* __restrict__ t, double param,
                                                            typename FilterState<FILTER>::type* __restrict__ state,
                                                            unsigned char* __restrict__ mask) {
  using ST = typename FilterState<FILTER>::type;
  using Sort = cub::BlockRadixSort<unsigned long long, FT_THREADS, FT_ITEMS>;
  using SortSmall = cub::BlockRadixSort<unsigned long long, FT_THREADS, FT_ITEMS_SMALL>;
  static_assert(sizeof(typename SortSmall::TempStorage) <= sizeof(typename Sort::TempStorage), "the small sort reuses the large one's scratch");
  extern __shared__ __align__(16) unsigned char ft_raw[];
  // the sort's scratch and the sorted keys share one region: the keys are written back after the sort is done with it
  constexpr size_t SORT_BYTES = sizeof(typename Sort::TempStorage) > sizeof(unsigned long long) * FT_CAP ? sizeof(typename Sort::TempStorage)
                                                                                                        : sizeof(unsigned long long) * FT_CAP;
  typename Sort::TempStorage& sort_tmp = *reinterpret_cast<typename Sort::TempStorage*>(ft_raw);
  typename SortSmall::TempStorage& sort_small = *reinterpret_cast<typename SortSmall::TempStorage*>(ft_raw);
  unsigned long long* sorted = reinterpret_cast<unsigned long long*>(ft_raw);                    // FT_CAP keys
  uint32_t* runs = reinterpret_cast<uint32_t*>(ft_raw + ((SORT_BYTES + 15) & ~(size_t)15));      // FT_MAX_SC + 1 run starts
  int* first = reinterpret_cast<int*>(runs + FT_MAX_SC + 1);                                    // TP: first sorted slot of a pixel
  ST* st_s = reinterpret_cast<ST*>(first + g.tile_px + ((g.tile_px + FT_MAX_SC + 1) & 1));      // TP states (8-byte aligned)
  __shared__ int s_seg_end;

  const int tid = threadIdx.x;
  const int b = blockIdx.x / g.T, tile = blockIdx.x - b * g.T;
  const int TP = g.tile_px, pix0 = tile << g.tile_shift, npix = min(TP, g.HW - pix0);
  const WinParams w = wp[b];
  const uint32_t count = hist[blockIdx.x];
  if (count == 0) return;  // nothing to decide, state unchanged
  const uint2* rec = records + w.start + base[blockIdx.x];
  const int sc0 = sc_prefix[b], nsc = sc_prefix[b + 1] - sc0;
  ST* gstate = state + (size_t)b * g.HW + pix0;
  for (int p = tid; p < npix; p += FT_THREADS) st_s[p] = gstate[p];
  // where every super-chunk's run starts inside this bucket; runs[nsc] = bucket size
  for (int s = tid; s < nsc; s += FT_THREADS) runs[s] = __ldg(cp + (size_t)(sc0 + s) * g.Tb + tile);
  if (tid == 0) runs[nsc] = count;
  __syncthreads();

  const double cell = (double)g.div_x * (double)g.div_y;
  int s_begin = 0;  // first super-chunk run of the current segment
  while (s_begin < nsc) {
    // the segment: as many whole runs as fit FT_CAP (one run always fits)
    if (tid == 0) {
      int e = s_begin + 1;
      while (e < nsc && runs[e + 1] - runs[s_begin] <= (uint32_t)FT_CAP) ++e;
      s_seg_end = e;
    }
    for (int p = tid; p < TP; p += FT_THREADS) first[p] = -1;
    __syncthreads();
    const int s_end = s_seg_end;
    const uint32_t a = runs[s_begin], n = runs[s_end] - a;
    auto load_key = [&](uint32_t i) -> unsigned long long {
      if (i >= n) return ~0ull;  // padding and null records sort to the end
      const uint2 r = __ldg(rec + a + i);
      if (rec_is_null(r.y)) return ~0ull;
      return ((unsigned long long)(r.y & 0xffffu) << 33) | ((unsigned long long)r.x << 2) | ((r.y >> 24) & 3u);
    };
    // sort by (pixel, stream index); the index is unique, so the polarity bits need no pass.  Most buckets hold about a
    // thousand records: a quarter-size sort (4 keys per thread) covers them, the full one is for hot tiles.
    if (n <= (uint32_t)(FT_THREADS * FT_ITEMS_SMALL)) {
      unsigned long long keys[FT_ITEMS_SMALL];
#pragma unroll
      for (int k = 0; k < FT_ITEMS_SMALL; ++k) keys[k] = load_key((uint32_t)(tid * FT_ITEMS_SMALL + k));
      SortSmall(sort_small).Sort(keys, 2, 49);
      __syncthreads();  // the scratch is about to be overwritten with the sorted keys
#pragma unroll
      for (int k = 0; k < FT_ITEMS_SMALL; ++k) sorted[tid * FT_ITEMS_SMALL + k] = keys[k];
    } else {
      unsigned long long keys[FT_ITEMS];
#pragma unroll
      for (int k = 0; k < FT_ITEMS; ++k) keys[k] = load_key((uint32_t)(tid * FT_ITEMS + k));
      Sort(sort_tmp).Sort(keys, 2, 49);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < FT_ITEMS; ++k) sorted[tid * FT_ITEMS + k] = keys[k];
    }
    __syncthreads();
    for (uint32_t j = tid; j < n; j += FT_THREADS) {
      const unsigned long long key = sorted[j];
      if (key == ~0ull) continue;
      const uint32_t pix = (uint32_t)(key >> 33);
      if (j == 0 || (uint32_t)(sorted[j - 1] >> 33) != pix) first[pix] = (int)j;
    }
    __syncthreads();
    for (int p = tid; p < npix; p += FT_THREADS) {
      int j = first[p];
      if (j < 0) continue;
      ST st = st_s[p];
      for (; (uint32_t)j < n; ++j) {
        const unsigned long long key = sorted[j];
        if (key == ~0ull || (uint32_t)(key >> 33) != (uint32_t)p) break;
        const uint32_t idx = (uint32_t)(key >> 2) & 0x7fffffffu;
        const uint32_t pc = (uint32_t)key & 3u;
        const int pv = pc == 1u ? 1 : (pc == 3u ? -1 : 0);
        bool keep;
        if (FILTER == EVREP_FILTER_REFRACTORY) {
          const double tt = (double)t[w.start + idx];
          keep = !(tt - (double)st < param);
          if (keep) st = (ST)tt;
        } else if (FILTER == EVREP_FILTER_BACKGROUND) {  // pc == 1: the event's own pixel (read, then write); else a neighbour's write
          const double tt = (double)t[w.start + idx];
          keep = pc == 1u && !((double)st > 0.0 && tt - (double)st > param);
          st = (ST)tt;
        } else if (FILTER == EVREP_FILTER_CONTRAST) {
          st = (ST)((int32_t)st + pv);
          keep = fabs((double)(int32_t)st) >= param;
          if (keep) st = (ST)0;
        } else {
          float c = (float)((double)(float)st + (double)pv * 1.0 / cell);  // float32 cell += float64 p * 1.0 / (fx * fy)
          keep = fabsf(c) >= 1.f;
          if (keep) c = c - (float)pv;
          st = (ST)c;
        }
        mask[w.start + idx] = keep ? 1 : 0;
      }
      st_s[p] = st;
    }
    __syncthreads();
    s_begin = s_end;
  }
  for (int p = tid; p < npix; p += FT_THREADS) gstate[p] = st_s[p];
}

template <int FILTER, typename TT>
static int launch_filter(const Geom& g, const Workspace& ws, const Events& ev, double param, void* state, unsigned char* mask,
                         cudaStream_t stream) {
  using ST = typename FilterState<FILTER>::type;
  using Sort = cub::BlockRadixSort<unsigned long long, FT_THREADS, FT_ITEMS>;
  const size_t sort_bytes = std::max(sizeof(typename Sort::TempStorage), sizeof(unsigned long long) * (size_t)FT_CAP);
  const size_t smem = ((sort_bytes + 15) & ~(size_t)15) + sizeof(uint32_t) * (FT_MAX_SC + 1) + sizeof(int) * (size_t)(g.tile_px + 1) +
                      sizeof(double) * (size_t)g.tile_px + 16;
  auto kern = k_filter_tile<FILTER, TT>;
  EVREP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(EVREP_K_TILE, stream);
  kern<<<g.B * g.T, FT_THREADS, smem, stream>>>(ws.records, ws.base, ws.hist, ws.cp, ws.sc_prefix, ws.wp, g, (const TT*)ev.t, param, (ST*)state, mask);
  prof_end(EVREP_K_TILE, stream);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

int launch_filter_tile(const Geom& g, const Workspace& ws, const Events& ev, int filter, double param, void* state, unsigned char* mask,
                       cudaStream_t stream) {
  if (ev.t_bytes == 4) {
    switch (filter) {
      case EVREP_FILTER_REFRACTORY: return launch_filter<EVREP_FILTER_REFRACTORY, int32_t>(g, ws, ev, param, state, mask, stream);
      case EVREP_FILTER_CONTRAST: return launch_filter<EVREP_FILTER_CONTRAST, int32_t>(g, ws, ev, param, state, mask, stream);
      case EVREP_FILTER_BACKGROUND: return launch_filter<EVREP_FILTER_BACKGROUND, int32_t>(g, ws, ev, param, state, mask, stream);
      default: return launch_filter<EVREP_FILTER_RESIZE, int32_t>(g, ws, ev, param, state, mask, stream);
    }
  }
  switch (filter) {
    case EVREP_FILTER_REFRACTORY: return launch_filter<EVREP_FILTER_REFRACTORY, int64_t>(g, ws, ev, param, state, mask, stream);
    case EVREP_FILTER_CONTRAST: return launch_filter<EVREP_FILTER_CONTRAST, int64_t>(g, ws, ev, param, state, mask, stream);
    case EVREP_FILTER_BACKGROUND: return launch_filter<EVREP_FILTER_BACKGROUND, int64_t>(g, ws, ev, param, state, mask, stream);
    default: return launch_filter<EVREP_FILTER_RESIZE, int64_t>(g, ws, ev, param, state, mask, stream);
  }
}

// ---------------------------------------------------------------------------------------------
// background-activity filter: expansion of the stream into per-pixel write records, and the mask read-back
// ---------------------------------------------------------------------------------------------
// Record k of event i (K = (2 radius)^2 records per event, expanded index K i + k): k = 0 is the event's own pixel with
// p = +1 ("read the state, decide, then write t"); k >= 1 walks the other offsets of the block with p = -1 ("write t").
// An offset outside the sensor is what numpy's slice clips away: it becomes another write of t to the own pixel, which
// sorts after record 0 and changes nothing.  Events outside the sensor get coordinates the binning drops.
template <typename TT>
__global__ void __launch_bounds__(256) k_ba_expand(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, const TT* __restrict__ t,
                                                   int64_t total, int H, int W, int radius, uint16_t* __restrict__ xe,
                                                   uint16_t* __restrict__ ye, TT* __restrict__ te, int8_t* __restrict__ pe) {
  const int side = 2 * radius, K = side * side, own = radius * side + radius;
  const int64_t n = total * K;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j / K;
    const int k = (int)(j - i * K);
    const int xi = (int)__ldg(x + i), yi = (int)__ldg(y + i);
    int px = xi, py = yi;
    if (xi >= W || yi >= H) {
      px = py = 0xffff;
    } else if (k > 0) {
      const int o = (k - 1 < own) ? k - 1 : k;
      const int qx = xi + (o % side) - radius, qy = yi + (o / side) - radius;
      if (qx >= 0 && qx < W && qy >= 0 && qy < H) { px = qx; py = qy; }
    }
    xe[j] = (uint16_t)px;
    ye[j] = (uint16_t)py;
    te[j] = __ldg(t + i);
    pe[j] = k == 0 ? (int8_t)1 : (int8_t)-1;
  }
}

__global__ void __launch_bounds__(256) k_ba_collect(const unsigned char* __restrict__ mask_e, int64_t total, int K, unsigned char* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) mask[i] = mask_e[i * K];
}

int launch_ba_expand(const Events& ev, int64_t total, int H, int W, int radius, uint16_t* xe, uint16_t* ye, void* te, int8_t* pe, cudaStream_t stream) {
  if (total == 0) return EVREP_OK;
  const int K = 4 * radius * radius;
  const int64_t n = total * K;
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 32);
  if (ev.t_bytes == 4)
    k_ba_expand<int32_t><<<grid, 256, 0, stream>>>(ev.x, ev.y, (const int32_t*)ev.t, total, H, W, radius, xe, ye, (int32_t*)te, pe);
  else
    k_ba_expand<int64_t><<<grid, 256, 0, stream>>>(ev.x, ev.y, (const int64_t*)ev.t, total, H, W, radius, xe, ye, (int64_t*)te, pe);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

int launch_ba_collect(const unsigned char* mask_e, int64_t total, int K, unsigned char* mask, cudaStream_t stream) {
  if (total == 0) return EVREP_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 32);
  k_ba_collect<<<grid, 256, 0, stream>>>(mask_e, total, K, mask);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

}  // namespace evrep
