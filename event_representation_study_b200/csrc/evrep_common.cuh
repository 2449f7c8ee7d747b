// libevrep internals shared by every translation unit.  Nothing in here is part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/evrep.h"
#include "md_device.cuh"

namespace evrep {

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
int launch_sbt_negsel(const Geom& g, const Workspace& ws, const Events& ev, cudaStream_t stream);  // SBT: which time windows hold a p == -1 event
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
