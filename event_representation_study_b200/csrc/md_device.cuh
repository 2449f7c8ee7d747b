// Device-side vocabulary shared by the ahead-of-time build (nvcc, through evrep_common.cuh) and the run-time specialised
// mixed-density kernels (NVRTC, md_jit.cu): constants, record layout, per-window parameters, tile geometry and the
// accumulator plan.  Self-contained on purpose - NVRTC sees no system or CUDA headers, only the three in-memory sources
// md_device.cuh, md_plan.cuh and md_tile_static.cuh.
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif

// MixedDensityEventStack vocabulary: include/evrep.h when it has been seen, the same values otherwise (NVRTC)
#ifndef EVREP_H
#define EVREP_FUNC_TIMESTAMP 0
#define EVREP_FUNC_POLARITY 1
#define EVREP_FUNC_COUNT 2
#define EVREP_FUNC_TIMESTAMP_POS 3
#define EVREP_FUNC_TIMESTAMP_NEG 4
#define EVREP_FUNC_COUNT_POS 5
#define EVREP_FUNC_COUNT_NEG 6
#define EVREP_AGG_SUM 0
#define EVREP_AGG_MEAN 1
#define EVREP_AGG_MAX 2
#define EVREP_AGG_VARIANCE 3
#define EVREP_AGG_MIN 4
#define EVREP_STACK_SBN 0
#define EVREP_STACK_SBT 1
#define EVREP_MAX_CHANNELS 32
#endif
static_assert(EVREP_FUNC_TIMESTAMP == 0 && EVREP_FUNC_POLARITY == 1 && EVREP_FUNC_COUNT == 2 && EVREP_FUNC_TIMESTAMP_POS == 3 &&
                  EVREP_FUNC_TIMESTAMP_NEG == 4 && EVREP_FUNC_COUNT_POS == 5 && EVREP_FUNC_COUNT_NEG == 6 && EVREP_AGG_SUM == 0 &&
                  EVREP_AGG_MEAN == 1 && EVREP_AGG_MAX == 2 && EVREP_AGG_VARIANCE == 3 && EVREP_AGG_MIN == 4 && EVREP_STACK_SBN == 0 &&
                  EVREP_STACK_SBT == 1 && EVREP_MAX_CHANNELS == 32,
              "md_device.cuh restates the vocabulary of include/evrep.h for NVRTC");

namespace evrep {

// ------------------------------------------------------------------------------------------------
// Binning geometry: the events of one window are cut into chunks of CHUNK events; one CTA bins one
// chunk.  Chunks start at an absolute event index that is a multiple of EPT so every thread's EPT
// consecutive events can be fetched with aligned 16-byte loads.
// ------------------------------------------------------------------------------------------------
constexpr int BIN_THREADS = 512;
constexpr int EPT = 8;
constexpr int CHUNK = BIN_THREADS * EPT;  // 4096 events per CTA iteration
#ifndef EVREP_SC_CHUNKS
#define EVREP_SC_CHUNKS 2
#endif
constexpr int SC_CHUNKS = EVREP_SC_CHUNKS;
constexpr int SUPER = CHUNK * SC_CHUNKS;  // 8192 events per CTA ("super-chunk"): the unit of the counting / scatter passes
constexpr int MAX_TILES = 4096;           // buckets per window (shared-memory histogram size bound)
constexpr int MIN_TILE_PX = 256;
constexpr int MAX_SNAP = 16;              // time-surface snapshots per window
constexpr int TILE_THREADS = 512;
constexpr int T_REL_LIMIT = 1 << 30;      // |t - t_first| must stay below this (microseconds)

// record payload written by the binning pass
enum RecMode : int {
  REC_T_WMASK = 0,  // key = t_rel, aux = SBN window mask           (mixed density)
  REC_IDX = 1,      // key = index inside window, aux = 0            (event stack)
  REC_T_SNAP = 2,   // key = t_rel, aux = first snapshot it feeds    (time surface)
  REC_T_TORE = 3,   // key = t_rel, events with t >= t_last dropped  (TORE)
  REC_T_ONLY = 4,   // key = t_rel, aux = 0: windows are derived from t later (mixed density, SBT)
  REC_T_IDX = 5,    // key = t_rel, meta = pixel (10 b) | stream index (20 b) | polarity code (2 b): EventStack + TimeSurface + TORE from one pass
};
// the fused record of REC_T_IDX (tiles of at most 1024 pixels, windows of fewer than 2^20 events)
__host__ __device__ inline uint32_t fused_meta(uint32_t pix, uint32_t idx, uint32_t pc) { return pix | (idx << 10) | (pc << 30); }
__host__ __device__ inline uint32_t fused_pix(uint32_t m) { return m & 0x3ffu; }
__host__ __device__ inline uint32_t fused_idx(uint32_t m) { return (m >> 10) & 0xfffffu; }
__host__ __device__ inline uint32_t fused_pc(uint32_t m) { return m >> 30; }  // 2 = null record
constexpr uint32_t FUSED_NULL_META = 2u << 30;
// meta word: [15:0] pixel inside tile, [23:16] aux, [25:24] polarity code (p & 3: 0 -> 0, 1 -> +1, 3 -> -1; 2 = null record)
__host__ __device__ inline uint32_t rec_meta(uint32_t pix, uint32_t aux, uint32_t pc) { return pix | (aux << 16) | (pc << 24); }
// A null record fills the slot of an event that was counted (valid x, y) but then dropped (timestamp out of range, after
// the last time-surface snapshot, at the TORE sample time): polarity code 2, member of no window.  Tile kernels skip it.
constexpr uint32_t REC_NULL_META = 2u << 24;
__host__ __device__ inline bool rec_is_null(uint32_t meta) { return ((meta >> 24) & 3u) == 2u; }

struct WinParams {  // one per window, lives at the start of the workspace
  int64_t start;    // absolute index of the first event
  int64_t n;        // number of events
  int64_t t_base;   // timestamp of the first event (t_rel = t - t_base)
  int32_t tmin_rel, tmax_rel;  // over accepted events
  int32_t tlast_rel;           // timestamp of the last event, relative
  uint32_t flags;              // EVREP_WF_*
  uint32_t has_m1;             // bit w set: window w of the mixed-density split holds an event with p == -1
  uint32_t pad;
};

struct SnapParams {  // time surface, one per window
  int32_t idx[MAX_SNAP];    // snapshot event indices (valid prefix only)
  int32_t t_rel[MAX_SNAP];  // timestamps at those indices
  int32_t n_valid;          // surfaces that the reference actually emits
  int32_t pad[3];
};

struct Geom {
  int B, H, W, HW;
  int tile_shift, tile_px, T;  // tile = contiguous range of tile_px linear pixel indices; T tiles per window
  int div_x, div_y;            // > 1: pixels are cells of div_x x div_y sensor pixels (x / div_x, y / div_y); W, H count cells
  int split;                   // 1: every tile has two buckets, p > 0 first, then the rest (mixed-density static kernels)
  int Tb;                      // buckets per window = T << split
  int64_t total;               // total events in the batch
  int64_t n_max;               // events of the largest window (0 = unknown)
  unsigned long long t_magic;  // ceil(2^44 / T): id / T == (id * t_magic) >> 44 for id < 2^32, T <= 4096
};

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ------------------------------------------------------------------------------------------------
// Mixed-density accumulator plan (built on the host from the (window, function, aggregation) tuple)
// ------------------------------------------------------------------------------------------------
enum { G_CNT = 1, G_PRES = 2, G_MAX = 4, G_ST = 8, G_ST2 = 16, G_MIN = 32 };
constexpr int MD_MAX_GROUPS = 32;

struct MdGroup {   // one (window, polarity class) pair that some channel reads
  uint8_t bit;     // membership bit: class * 8 + window; class 0 = all, 1 = p == 1, 2 = "negative", 3 = neither
  uint8_t flags;   // G_*
  uint8_t w_cnt, w_max, w_st, w_st2;  // accumulator word indices
  uint8_t pres_bit;
  uint8_t cnt_shift;  // packed plans keep two 16-bit counters per word: 0 or 16
  uint8_t w_min;      // earliest timestamp, kept as the maximum of ~t (0 = untouched)
};
struct MdChan {
  uint8_t func, agg, win, valid;
  int8_t g_main;               // group holding the sums / latest timestamp / presence bit / single-class count
  int8_t g_pos, g_neg, g_oth;  // class counters; "all events" counts are their sum (an event bumps exactly one of them)
};
struct MdPlan {
  int32_t C, G, words, stride, nl1, nl2, lw, w_pres, stacking;
  int32_t static_id;  // 0, or version * 100 + limb width of a compile-time specialised ERGO-12 kernel
  int32_t packed;     // 1: 16-bit counters and 16-bit limbs, valid for buckets of fewer than 65536 events
  int32_t pad;
  MdGroup grp[MD_MAX_GROUPS];
  MdChan ch[EVREP_MAX_CHANNELS];
};
constexpr uint32_t MD_PACKED_LIMIT = 65536;  // a packed plan may only see buckets with fewer events than this

}  // namespace evrep
