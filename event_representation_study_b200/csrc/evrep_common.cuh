// libevrep internals shared by every translation unit.  Nothing in here is part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/evrep.h"

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

struct Workspace {  // device pointers carved out of the caller's buffer
  WinParams* wp;
  int64_t* offsets;
  int32_t* sc_prefix;  // B+1   first super-chunk of each window
  int32_t* sc_win;     // n_sc  window of each super-chunk
  uint32_t* ticket;    // 64 words: work counters of persistent kernels
  uint32_t* hist;      // B*Tb  bucket sizes (events with valid x, y; dropped ones keep a null record)
  uint32_t* base;      // B*Tb  bucket start, relative to the window's first record
  uint16_t* cc;        // n_sc*cc_stride(Tb)  per super-chunk: exclusive scan of its bucket counts, then the total (<= SUPER)
  uint32_t* cp;        // n_sc*Tb  exclusive prefix of cc over the window's super-chunks: where the super-chunk's run starts inside the bucket
  SnapParams* snap;    // B
  int64_t* snap_in;  // B*MAX_SNAP caller-supplied snapshot indices
  double* stats;     // B*4   voxel normalisation sums
  uint2* records;    // total events
  size_t bytes;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// row stride of `cc` in entries: Tb offsets + the total, padded to an even count so that every row starts 4-byte aligned
__host__ __device__ inline int cc_stride(int Tb) { return (Tb + 2) & ~1; }

// Carves the workspace; with base == nullptr only computes the size.
inline int64_t max_super_chunks(int B, int64_t total) { return total / SUPER + 2 * (int64_t)B + 1; }
inline Workspace carve(void* basep, int B, int64_t total, int T) {
  Workspace w;
  size_t off = 0;
  char* base = (char*)basep;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.wp = (WinParams*)take(sizeof(WinParams) * (size_t)B);
  // offsets and the super-chunk prefix are uploaded with ONE host-to-device copy: keep them contiguous
  w.offsets = (int64_t*)take(sizeof(int64_t) * (size_t)(B + 1) + sizeof(int32_t) * (size_t)(B + 1));
  w.sc_prefix = base ? (int32_t*)(w.offsets + (B + 1)) : nullptr;
  const size_t n_sc = total > 0 ? (size_t)max_super_chunks(B, total) : 1;
  w.sc_win = (int32_t*)take(sizeof(int32_t) * n_sc);
  w.ticket = (uint32_t*)take(sizeof(uint32_t) * 64);
  w.hist = (uint32_t*)take(sizeof(uint32_t) * (size_t)B * T);
  w.base = (uint32_t*)take(sizeof(uint32_t) * (size_t)B * T);
  w.cc = (uint16_t*)take(sizeof(uint16_t) * n_sc * (size_t)cc_stride(T));
  w.cp = (uint32_t*)take(sizeof(uint32_t) * n_sc * (size_t)T);
  w.snap = (SnapParams*)take(sizeof(SnapParams) * (size_t)B);
  w.snap_in = (int64_t*)take(sizeof(int64_t) * (size_t)B * MAX_SNAP);
  w.stats = (double*)take(sizeof(double) * 4 * (size_t)B);
  w.records = (uint2*)take(sizeof(uint2) * (size_t)(total > 0 ? total : 1));
  w.bytes = off;
  return w;
}

// thread-local error text (api.cu)
void set_error(const char* fmt, ...);

// Optional per-kernel timing with CUDA events on the caller's stream (api.cu; evrep_profile_* in evrep.h).
// No-ops unless profiling was enabled; never synchronise.
void prof_next_call();
// Programmatic dependent launch between the kernels of one call (and from one call to the next): a kernel launched through
// launch_pdl may be scheduled while its predecessor in the stream drains - its CTAs become resident as the predecessor's
// retire and run up to pdl_wait(), which returns once the predecessor has completed and its writes are visible.  Every
// kernel of the chain calls pdl_wait() before its first global-memory access and pdl_trigger() right after it (so a
// dependent never runs ahead of a kernel that has not itself seen ITS predecessor complete); launched the ordinary way both
// are no-ops.  What it hides is the launch latency and the ramp of each of the six kernels of a call (~2 us apiece).
// EVREP_NO_PDL=1 falls back to ordinary launches (A/B measurements).
bool pdl_enabled();
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

void prof_begin(int kernel_id, cudaStream_t stream);
void prof_end(int kernel_id, cudaStream_t stream);

#define EVREP_CUDA_OK(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      evrep::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return EVREP_ECUDA;                                                                    \
    }                                                                                        \
  } while (0)

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

// ------------------------------------------------------------------------------------------------
// Launchers (each returns an EVREP_* code and enqueues on `stream`)
// ------------------------------------------------------------------------------------------------
struct Events {
  const uint16_t* x;
  const uint16_t* y;
  const void* t;
  int t_bytes;
  const int8_t* p;
};

// memset + offsets upload + per-window init + count + scan + bin.  After it returns (stream order)
// ws.records holds the tile-bucketed records, ws.cursor the bucket sizes, ws.base the bucket starts.
int run_binning(const Events& ev, const int64_t* win_offsets_host, const Geom& g, const Workspace& ws, int rec_mode,
                int n_snap, const int64_t* snap_indices_host, cudaStream_t stream);

int prepare_windows(const Events& ev, const int64_t* win_offsets_host, const Geom& g, const Workspace& ws, int* n_super_chunks,
                    cudaStream_t stream);
bool events_vectorisable(const Events& ev);

int launch_md_tile(const Geom& g, const Workspace& ws, const MdPlan& plan, const Events& ev, float* out, cudaStream_t stream);
int launch_event_stack_tile(const Geom& g, const Workspace& ws, int stack_size, float* out, cudaStream_t stream);
int launch_time_surface_tile(const Geom& g, const Workspace& ws, int S, double tau, float* out, cudaStream_t stream);
int launch_tore_tile(const Geom& g, const Workspace& ws, int k, float* out, cudaStream_t stream);
int launch_order_ops_fused(const Geom& g, const Workspace& ws, double tau, float* out_es, float* out_ts, float* out_tore, cudaStream_t stream);
int launch_filter_tile(const Geom& g, const Workspace& ws, const Events& ev, int filter, double param, void* state, unsigned char* mask,
                       cudaStream_t stream);

int launch_ba_expand(const Events& ev, int64_t total, int H, int W, int radius, uint16_t* xe, uint16_t* ye, void* te, int8_t* pe, cudaStream_t stream);
int launch_ba_collect(const unsigned char* mask_e, int64_t total, int K, unsigned char* mask, cudaStream_t stream);

int launch_voxel(const Events& ev, const int64_t* win_offsets_host, const Geom& g, const Workspace& ws, int flavour,
                 int n_bins, int normalize, const int64_t* t0_t1_host, int divider, float* out, cudaStream_t stream);
int launch_histogram(const Events& ev, const int64_t* win_offsets_host, const Geom& g, const Workspace& ws, float* out,
                     cudaStream_t stream);

size_t md_tile_smem_bytes(const MdPlan& plan, int tile_px);

}  // namespace evrep
