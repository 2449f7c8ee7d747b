// The C ABI of libevrep (include/evrep.h): argument validation, tile geometry, workspace carving and
// kernel sequencing.  No exceptions leave this file; every entry point returns an EVREP_* code.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include <cmath>

#include "evrep_common.cuh"
#include "md_jit.h"

namespace evrep {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- optional kernel timing -------------------------------------------------------------------
// thread-local like the error text: a host thread that profiles its own calls neither sees nor disturbs another thread's
struct Prof {
  int max_calls = 0, call = -1;
  cudaEvent_t* ev = nullptr;  // [max_calls][EVREP_K_N][2]
  unsigned char* used = nullptr;
};
static thread_local Prof g_prof;

static void prof_free() {
  if (g_prof.ev) {
    for (int i = 0; i < g_prof.max_calls * EVREP_K_N * 2; ++i) cudaEventDestroy(g_prof.ev[i]);
    delete[] g_prof.ev;
    delete[] g_prof.used;
  }
  g_prof = Prof();
}
void prof_next_call() {
  if (g_prof.max_calls) ++g_prof.call;
}
void prof_begin(int k, cudaStream_t stream) {
  if (!g_prof.max_calls || g_prof.call < 0 || g_prof.call >= g_prof.max_calls) return;
  cudaEventRecord(g_prof.ev[(g_prof.call * EVREP_K_N + k) * 2], stream);
}
void prof_end(int k, cudaStream_t stream) {
  if (!g_prof.max_calls || g_prof.call < 0 || g_prof.call >= g_prof.max_calls) return;
  cudaEventRecord(g_prof.ev[(g_prof.call * EVREP_K_N + k) * 2 + 1], stream);
  g_prof.used[g_prof.call * EVREP_K_N + k] = 1;
}

int build_md_plan(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int64_t n_max, MdPlan* out);
size_t gwd_workspace_bytes(const int64_t* so, const int64_t* to, int n_pairs);
int launch_gwd(const double* Xs, const int64_t* so, int ds, const double* Xt, const int64_t* to, int dt, int n_pairs, double h,
               double* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);

size_t gw_kl_workspace_bytes(int n, int m);
int run_gw_kl(const double* Xs, int n, int ds, const double* Xt, int m, int dt, double h, int max_iter, double tol_rel, double tol_abs,
              int lmo, double* gw_dist_host, float* T_out, int* iters_host, int* lmo_stats_host, void* workspace, size_t workspace_bytes,
              cudaStream_t stream);
size_t gemm_workspace_bytes(int M, int N, int K);
int launch_gemm_nt_3xtf32_ws(const float* A, const float* B, float* C, int M, int N, int K, float alpha, const float* rv, const float* cv,
                             void* workspace, size_t workspace_bytes, cudaStream_t stream);

size_t est_workspace_bytes(int B);
int launch_est(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets_host, int B, int H, int W, int C,
               const double* breaks, const double* slope, const double* icpt, int K, float* out, void* workspace, size_t workspace_bytes,
               cudaStream_t stream);
int launch_est_backward(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets_host, int B, int H, int W,
                        int C, const double* breaks, int K, const float* grad_out, double* seg_sums, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream);
int launch_auction(const float* cost, int n, double eps_rel, int* sigma, int* stats, cudaStream_t stream);
int transport_plan_host(const float* cost, int n, int m, int cap, int* row_ptr, int* col, double* weight, int* nnz_out);
size_t unpack_workspace_bytes(int B, int64_t total);
int launch_warp_affine(const float* in, int B, int C, int in_h, int in_w, const double* M_host, const int* flips_host, int out_h, int out_w,
                       const float* border4, int reverse, float scale_out, float* out, cudaStream_t stream);
size_t unpack_delta_workspace_bytes(int B);
int launch_unpack_delta(const uint8_t* rec3, const int32_t* tbase, const uint32_t* esc_prefix, const uint32_t* esc_dt, const int64_t* win_offsets_host,
                        int B, int xb, int yb, uint16_t* x, uint16_t* y, int32_t* t, int8_t* p, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream);
size_t otmi_workspace_bytes(long long N, int R);
int launch_otmi_prepare(const void* ev, int ev_type, long long N, const double* rep, int R, int C, int height, int width, double* Xs, long long xs_cap,
                        double* Xt, long long xt_cap, long long* info_host, void* workspace, cudaStream_t stream);
int launch_unpack(const uint32_t* word, const uint16_t* dt16, const int32_t* tbase, const int64_t* win_offsets_host, int B, int fmt, int xb, int yb,
                  int blk_shift, uint16_t* x, uint16_t* y, int32_t* t, int8_t* p, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_image_pipeline(const float* rep, int B, int H, int W, int C, int img_size, int mode, int interp, float scale_in, float scale_out,
                          float pad, int reverse, float* out, cudaStream_t stream);

constexpr size_t TILE_SMEM_TARGET = 100 * 1024;  // two tile CTAs per SM
constexpr size_t TILE_SMEM_MAX = 220 * 1024;

// tile_px = largest power of two (<= 4096) whose accumulators fit the target; T = tiles per window
static int choose_tile(int H, int W, size_t bytes_per_px, Geom* g, int split = 0) {
  if (H < 1 || W < 1 || H > 65535 || W > 65535 || (int64_t)H * W > ((int64_t)1 << 28)) {
    set_error("sensor size %d x %d unsupported", W, H);
    return EVREP_EINVAL;
  }
  g->H = H;
  g->W = W;
  g->HW = H * W;
  size_t target = TILE_SMEM_TARGET;
  if (const char* e = getenv("EVREP_TILE_SMEM_KB")) {  // tuning knob for experiments
    const long kb = atol(e);
    if (kb >= 8 && kb <= 220) target = (size_t)kb * 1024;
  }
  int shift = 12;
  while (shift > 8 && ((size_t)1 << shift) * bytes_per_px > target) --shift;
  // a very large sensor needs bigger tiles than the target allows: trade occupancy for reach
  while ((((g->HW + (1 << shift) - 1) >> shift) << split) > MAX_TILES && shift < 16) ++shift;
  if (((size_t)1 << shift) * bytes_per_px > TILE_SMEM_MAX || (((g->HW + (1 << shift) - 1) >> shift) << split) > MAX_TILES) {
    set_error("%d x %d pixels with %zu accumulator bytes per pixel does not fit the tile pipeline", W, H, bytes_per_px);
    return EVREP_EUNSUPPORTED;
  }
  g->tile_shift = shift;
  g->tile_px = 1 << shift;
  g->T = (g->HW + g->tile_px - 1) >> shift;
  g->split = split;
  g->div_x = g->div_y = 1;
  g->Tb = g->T << split;
  g->t_magic = ((1ull << 44) + (unsigned long long)g->T - 1) / (unsigned long long)g->T;
  return EVREP_OK;
}

static int check_events(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                        int B, const void* out, Events* ev, int64_t* total, int64_t* n_max) {
  if (B < 0 || !win_offsets) { set_error("B must be >= 0 and win_offsets non-null"); return EVREP_EINVAL; }
  if (B >= (1 << 20)) { set_error("at most 2^20 - 1 windows per call"); return EVREP_EUNSUPPORTED; }
  if (t_bytes != 4 && t_bytes != 8) { set_error("t_bytes must be 4 (int32) or 8 (int64), got %d", t_bytes); return EVREP_EINVAL; }
  *total = win_offsets[B];
  *n_max = 0;
  for (int b = 0; b < B; ++b) *n_max = std::max<int64_t>(*n_max, win_offsets[b + 1] - win_offsets[b]);
  if (*total > 0 && (!x || !y || !t || !p)) { set_error("null event array"); return EVREP_EINVAL; }
  if (B > 0 && !out) { set_error("null output"); return EVREP_EINVAL; }
  ev->x = x; ev->y = y; ev->t = t; ev->t_bytes = t_bytes; ev->p = p;
  return EVREP_OK;
}

static int carve_checked(void* workspace, size_t workspace_bytes, int B, int64_t total, int T, Workspace* ws) {
  *ws = carve(workspace, B, total, T);
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
    set_error("workspace must be non-null and 256-byte aligned");
    return EVREP_EWORKSPACE;
  }
  if (ws->bytes > workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %zu", ws->bytes, workspace_bytes);
    return EVREP_EWORKSPACE;
  }
  return EVREP_OK;
}

}  // namespace evrep

using namespace evrep;

#define EVREP_TRY(expr)            \
  do {                             \
    int _rc = (expr);              \
    if (_rc != EVREP_OK) return _rc; \
  } while (0)

#define EVREP_GUARD_BEGIN try {
#define EVREP_GUARD_END                                   \
  }                                                       \
  catch (const std::bad_alloc&) {                         \
    set_error("host allocation failed");                  \
    return EVREP_EINVAL;                                  \
  }                                                       \
  catch (...) {                                           \
    set_error("unexpected C++ exception");                \
    return EVREP_EINVAL;                                  \
  }

extern "C" {

int evrep_version(void) { return EVREP_VERSION; }

const char* evrep_last_error(void) { return g_err; }

int evrep_profile_enable(int max_calls) {
  prof_free();
  if (max_calls <= 0) return EVREP_OK;
  if (max_calls > 4096) { set_error("max_calls > 4096"); return EVREP_EINVAL; }
  g_prof.ev = new (std::nothrow) cudaEvent_t[(size_t)max_calls * EVREP_K_N * 2];
  g_prof.used = new (std::nothrow) unsigned char[(size_t)max_calls * EVREP_K_N]();
  if (!g_prof.ev || !g_prof.used) { set_error("host allocation failed"); return EVREP_EINVAL; }
  g_prof.max_calls = max_calls;
  g_prof.call = -1;
  for (int i = 0; i < max_calls * EVREP_K_N * 2; ++i) EVREP_CUDA_OK(cudaEventCreate(&g_prof.ev[i]));
  return EVREP_OK;
}

int evrep_profile_read(int kernel_id, float* total_ms, int* launches) {
  if (kernel_id < 0 || kernel_id >= EVREP_K_N || !total_ms || !launches) { set_error("bad argument"); return EVREP_EINVAL; }
  *total_ms = 0.f;
  *launches = 0;
  for (int c = 0; c < g_prof.max_calls && c <= g_prof.call; ++c) {
    if (!g_prof.used[c * EVREP_K_N + kernel_id]) continue;
    cudaEvent_t a = g_prof.ev[(c * EVREP_K_N + kernel_id) * 2], b = g_prof.ev[(c * EVREP_K_N + kernel_id) * 2 + 1];
    EVREP_CUDA_OK(cudaEventSynchronize(b));
    float ms = 0.f;
    EVREP_CUDA_OK(cudaEventElapsedTime(&ms, a, b));
    *total_ms += ms;
    ++*launches;
  }
  return EVREP_OK;
}

size_t evrep_workspace_bytes(int op, int B, int64_t total_events, int H, int W, int C) {
  (void)C;
  if (op < EVREP_OP_MIXED_DENSITY || op > EVREP_OP_FILTER || B < 0 || total_events < 0 || H < 1 || W < 1) return 0;
  const int64_t hw = (int64_t)H * W;
  int64_t T = 2 * ((hw + MIN_TILE_PX - 1) / MIN_TILE_PX);  // buckets per window: at most two per tile
  if (T > MAX_TILES) T = MAX_TILES;
  const bool tiles = op != EVREP_OP_VOXEL && op != EVREP_OP_HISTOGRAM;
  return carve(nullptr, B > 0 ? B : 1, tiles ? total_events : 0, tiles ? (int)T : 1).bytes;
}

int evrep_window_flags(const void* workspace, int B, uint32_t* flags_host, evrep_stream_t stream) {
  if (!workspace || !flags_host || B < 0) { set_error("bad argument"); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  const WinParams* wp = (const WinParams*)workspace;
  EVREP_CUDA_OK(cudaMemcpy2DAsync(flags_host, sizeof(uint32_t), &wp[0].flags, sizeof(WinParams), sizeof(uint32_t), (size_t)B,
                                  cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  EVREP_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return EVREP_OK;
}

int evrep_mixed_density_plan_info(int H, int W, const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking,
                                  int64_t max_events_per_window, int* info) {
  EVREP_GUARD_BEGIN
  if (!win || !func || !agg || !info) { set_error("null argument"); return EVREP_EINVAL; }
  MdPlan plan;
  EVREP_TRY(build_md_plan(win, func, agg, C, stacking, max_events_per_window, &plan));
  Geom g;
  memset(&g, 0, sizeof(g));
  EVREP_TRY(choose_tile(H, W, (size_t)plan.stride * 4, &g));
  info[0] = plan.stride * 4;
  info[1] = g.tile_px;
  info[2] = g.T;
  info[3] = plan.words;
  info[4] = (int)md_tile_smem_bytes(plan, g.tile_px);
  return EVREP_OK;
  EVREP_GUARD_END
}

// tiles of 1024 or 512 pixels, two buckets per tile (p > 0 first): the geometry of the compile-time specialised kernels
static bool tile_split(int H, int W, int tile_px, Geom* g) {
  if (H < 1 || W < 1 || H > 65535 || W > 65535) return false;
  const int shift = tile_px == 512 ? 9 : 10;
  const int64_t hw = (int64_t)H * W;
  const int64_t T = (hw + tile_px - 1) >> shift;
  if (2 * T > MAX_TILES) return false;
  g->H = H;
  g->W = W;
  g->HW = (int)hw;
  g->tile_shift = shift;
  g->tile_px = tile_px;
  g->T = (int)T;
  g->split = 1;
  g->div_x = g->div_y = 1;
  g->Tb = g->T << 1;
  g->t_magic = ((1ull << 44) + (unsigned long long)g->T - 1) / (unsigned long long)g->T;
  return true;
}

int evrep_mixed_density_specialize(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int64_t max_events_per_window) {
  EVREP_GUARD_BEGIN
  if (!win || !func || !agg) { set_error("null channel description"); return EVREP_EINVAL; }
  MdPlan plan;
  EVREP_TRY(build_md_plan(win, func, agg, C, stacking, max_events_per_window > 0 ? max_events_per_window : 1, &plan));
  if (plan.static_id) return EVREP_OK;  // an ERGO-12 tuple: its kernels were built ahead of time
  return evrep::md_jit_specialize(win, func, agg, C, stacking, max_events_per_window, true, nullptr);
  EVREP_GUARD_END
}

int evrep_mixed_density_specialize_async(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int64_t max_events_per_window) {
  EVREP_GUARD_BEGIN
  if (!win || !func || !agg) { set_error("null channel description"); return EVREP_EINVAL; }
  MdPlan plan;
  EVREP_TRY(build_md_plan(win, func, agg, C, stacking, max_events_per_window > 0 ? max_events_per_window : 1, &plan));
  if (plan.static_id) return EVREP_OK;
  return evrep::md_jit_specialize(win, func, agg, C, stacking, max_events_per_window, false, nullptr, false);
  EVREP_GUARD_END
}

int evrep_mixed_density_specialize_compile_only(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking,
                                                int64_t max_events_per_window, size_t* cubin_bytes) {
  EVREP_GUARD_BEGIN
  if (!win || !func || !agg) { set_error("null channel description"); return EVREP_EINVAL; }
  return evrep::md_jit_specialize(win, func, agg, C, stacking, max_events_per_window, false, cubin_bytes);
  EVREP_GUARD_END
}

int evrep_mixed_density_is_specialized(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int64_t max_events_per_window) {
  if (!win || !func || !agg || C < 1 || C > EVREP_MAX_CHANNELS) return 0;
  return evrep::md_jit_find(win, func, agg, C, stacking, max_events_per_window) ? 1 : 0;
}

int evrep_mixed_density_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                                const int64_t* win_offsets, int B, int H, int W, const int8_t* win, const int8_t* func,
                                const int8_t* agg, int C, int stacking, float* out, void* workspace, size_t workspace_bytes,
                                evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  Events ev;
  int64_t total = 0, n_max = 0;
  EVREP_TRY(check_events(x, y, t, t_bytes, p, win_offsets, B, out, &ev, &total, &n_max));
  if (!win || !func || !agg) { set_error("null channel description"); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  MdPlan plan;
  EVREP_TRY(build_md_plan(win, func, agg, C, stacking, n_max, &plan));
  Geom g;
  memset(&g, 0, sizeof(g));
  // a tuple the caller has specialised (evrep_mixed_density_specialize): the ERGO-12 pipeline with that tuple's kernels
  if (!plan.static_id && evrep::md_jit_count() > 0) {
    if (evrep::JitProgram* jp = evrep::md_jit_find(win, func, agg, C, stacking, n_max)) {
      // (bulk stores move whole 16-byte units: any C on sensors with H * W * C a multiple of 4)
      // a program that cannot be loaded on this device (compiled elsewhere, driver library missing) leaves the call to the interpreted kernel
      if (((int64_t)H * W * C) % 4 == 0 && tile_split(H, W, evrep::md_jit_tile_px(jp), &g) && evrep::md_jit_prepare(jp) == EVREP_OK) {
        g.B = B;
        g.total = total;
        g.n_max = n_max;
        Workspace ws;
        EVREP_TRY(carve_checked(workspace, workspace_bytes, B, total, g.Tb, &ws));
        EVREP_TRY(run_binning(ev, win_offsets, g, ws, stacking == EVREP_STACK_SBN ? REC_T_WMASK : REC_T_ONLY, 0, nullptr, (cudaStream_t)stream));
        if (stacking == EVREP_STACK_SBT) EVREP_TRY(evrep::launch_sbt_negsel(g, ws, ev, (cudaStream_t)stream));
        return evrep::md_jit_launch(jp, g, ws, out, (cudaStream_t)stream);
      }
      memset(&g, 0, sizeof(g));
    }
  }
  EVREP_TRY(choose_tile(H, W, (size_t)plan.stride * 4, &g));
  // the compile-time specialised ERGO-12 kernels (1024-pixel tiles) want every tile's events split by polarity
  if (plan.static_id && g.tile_px == 1024 && 2 * g.T <= MAX_TILES) EVREP_TRY(choose_tile(H, W, (size_t)plan.stride * 4, &g, 1));
  if (g.tile_px != 1024) g.split = 0, g.Tb = g.T;
  g.B = B;
  g.total = total;
  g.n_max = n_max;
  Workspace ws;
  EVREP_TRY(carve_checked(workspace, workspace_bytes, B, total, g.Tb, &ws));
  EVREP_TRY(run_binning(ev, win_offsets, g, ws, stacking == EVREP_STACK_SBN ? REC_T_WMASK : REC_T_ONLY, 0, nullptr, (cudaStream_t)stream));
  return launch_md_tile(g, ws, plan, ev, out, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_ergo12_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                         int B, int H, int W, int version, float* out, void* workspace, size_t workspace_bytes, evrep_stream_t stream) {
  // representations/optimized_representation.py:86-115 (v2, active) and :16-66 (v1, commented out)
  enum { T_ = EVREP_FUNC_TIMESTAMP, P_ = EVREP_FUNC_POLARITY, C_ = EVREP_FUNC_COUNT, TP = EVREP_FUNC_TIMESTAMP_POS,
         TN = EVREP_FUNC_TIMESTAMP_NEG, CP = EVREP_FUNC_COUNT_POS, CN = EVREP_FUNC_COUNT_NEG };
  enum { SUM = EVREP_AGG_SUM, MEAN = EVREP_AGG_MEAN, MAX = EVREP_AGG_MAX, VAR = EVREP_AGG_VARIANCE };
  static const int8_t w2[12] = {0, 3, 2, 6, 5, 6, 2, 5, 1, 0, 4, 1};
  static const int8_t f2[12] = {P_, TN, CN, P_, CP, C_, TP, CN, TN, TP, T_, C_};
  static const int8_t a2[12] = {VAR, VAR, MEAN, SUM, MEAN, SUM, MEAN, MEAN, MAX, MAX, MAX, MEAN};
  static const int8_t w1[12] = {0, 2, 2, 3, 5, 0, 0, 4, 2, 6, 1, 1};
  static const int8_t f1[12] = {T_, TP, TN, CN, CP, P_, T_, C_, TP, C_, TP, TN};
  static const int8_t a1[12] = {MAX, SUM, MEAN, SUM, MEAN, VAR, VAR, SUM, MEAN, SUM, SUM, SUM};
  if (version != 1 && version != 2) { set_error("ERGO-12 version must be 1 or 2"); return EVREP_EINVAL; }
  return evrep_mixed_density_batched(x, y, t, t_bytes, p, win_offsets, B, H, W, version == 2 ? w2 : w1, version == 2 ? f2 : f1,
                                     version == 2 ? a2 : a1, 12, EVREP_STACK_SBN, out, workspace, workspace_bytes, stream);
}

int evrep_event_stack_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                              const int64_t* win_offsets, int B, int H, int W, int stack_size, float* out, void* workspace,
                              size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  Events ev;
  int64_t total = 0, n_max = 0;
  EVREP_TRY(check_events(x, y, t, t_bytes, p, win_offsets, B, out, &ev, &total, &n_max));
  if (stack_size < 1 || stack_size > EVREP_MAX_CHANNELS) { set_error("stack_size outside 1..%d", EVREP_MAX_CHANNELS); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  Geom g;
  memset(&g, 0, sizeof(g));
  EVREP_TRY(choose_tile(H, W, 4, &g));
  g.B = B;
  g.total = total;
  Workspace ws;
  EVREP_TRY(carve_checked(workspace, workspace_bytes, B, total, g.Tb, &ws));
  EVREP_TRY(run_binning(ev, win_offsets, g, ws, REC_IDX, 0, nullptr, (cudaStream_t)stream));
  return launch_event_stack_tile(g, ws, stack_size, out, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_time_surface_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                               const int64_t* win_offsets, int B, int H, int W, const int64_t* indices, int S, double tau, float* out,
                               void* workspace, size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  Events ev;
  int64_t total = 0, n_max = 0;
  EVREP_TRY(check_events(x, y, t, t_bytes, p, win_offsets, B, out, &ev, &total, &n_max));
  if (S < 1 || S > MAX_SNAP) { set_error("S outside 1..%d", MAX_SNAP); return EVREP_EINVAL; }
  if (!(tau > 0.0)) { set_error("tau must be positive"); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  Geom g;
  memset(&g, 0, sizeof(g));
  EVREP_TRY(choose_tile(H, W, (size_t)8 * S * (S == 6 ? 2 : 1), &g));  // S == 6: 1024-pixel tiles, four CTAs per SM (measured: 0.431 -> 0.399 ms at 1 Mpx)
  g.B = B;
  g.total = total;
  Workspace ws;
  EVREP_TRY(carve_checked(workspace, workspace_bytes, B, total, g.Tb, &ws));
  EVREP_TRY(run_binning(ev, win_offsets, g, ws, REC_T_SNAP, S, indices, (cudaStream_t)stream));
  return launch_time_surface_tile(g, ws, S, tau, out, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_tore_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                       int B, int H, int W, int k, float* out, void* workspace, size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  Events ev;
  int64_t total = 0, n_max = 0;
  EVREP_TRY(check_events(x, y, t, t_bytes, p, win_offsets, B, out, &ev, &total, &n_max));
  if (k < 1 || 2 * k > EVREP_MAX_CHANNELS) { set_error("k outside 1..%d", EVREP_MAX_CHANNELS / 2); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  Geom g;
  memset(&g, 0, sizeof(g));
  // 8 k bytes of accumulators per pixel; the k = 6 kernel also stages 24 KB of output, which would leave one 2048-pixel
  // CTA per SM: ask for 1024-pixel tiles there (72 KB, three CTAs per SM)
  EVREP_TRY(choose_tile(H, W, (size_t)8 * k + (k == 6 ? 12 : 0), &g));
  g.B = B;
  g.total = total;
  Workspace ws;
  EVREP_TRY(carve_checked(workspace, workspace_bytes, B, total, g.Tb, &ws));
  EVREP_TRY(run_binning(ev, win_offsets, g, ws, REC_T_TORE, 0, nullptr, (cudaStream_t)stream));
  return launch_tore_tile(g, ws, k, out, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_order_ops_fused_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                                  int B, int H, int W, double tau, float* out_es, float* out_ts, float* out_tore, void* workspace,
                                  size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  Events ev;
  int64_t total = 0, n_max = 0;
  EVREP_TRY(check_events(x, y, t, t_bytes, p, win_offsets, B, out_es, &ev, &total, &n_max));
  if (!(tau > 0.0)) { set_error("tau must be positive"); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  if (!out_ts || !out_tore) { set_error("null output"); return EVREP_EINVAL; }
  if (n_max >= ((int64_t)1 << 20)) { set_error("fused order ops: windows must hold fewer than 2^20 events (got %lld)", (long long)n_max); return EVREP_EUNSUPPORTED; }
  Geom g;
  memset(&g, 0, sizeof(g));
  EVREP_TRY(choose_tile(H, W, 60, &g));  // 1024-pixel tiles: what TORE's staged kernel wants and the fused record's 10 pixel bits allow
  if (g.tile_px > 1024) { set_error("fused order ops: sensor too large for 1024-pixel tiles"); return EVREP_EUNSUPPORTED; }
  g.B = B;
  g.total = total;
  Workspace ws;
  EVREP_TRY(carve_checked(workspace, workspace_bytes, B, total, g.Tb, &ws));
  EVREP_TRY(run_binning(ev, win_offsets, g, ws, REC_T_IDX, 6, nullptr, (cudaStream_t)stream));
  return launch_order_ops_fused(g, ws, tau, out_es, out_ts, out_tore, (cudaStream_t)stream);
  EVREP_GUARD_END
}

static int voxel_common(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets, int B, int H,
                        int W, int flavour, int n_bins, int normalize, const int64_t* t0_t1_us, int divider, float* out, void* workspace,
                        size_t workspace_bytes, evrep_stream_t stream) {
  Events ev;
  int64_t total = 0, n_max = 0;
  EVREP_TRY(check_events(x, y, t, t_bytes, p, win_offsets, B, out, &ev, &total, &n_max));
  if (flavour < EVREP_VOXEL_TONIC || flavour > EVREP_VOXEL_GWD) { set_error("unknown voxel flavour %d", flavour); return EVREP_EINVAL; }
  if (n_bins < 1 || n_bins > 1024) { set_error("n_bins outside 1..1024"); return EVREP_EINVAL; }
  if (divider < 1 || divider > 65535) { set_error("divider outside 1..65535"); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  Geom g;
  memset(&g, 0, sizeof(g));
  EVREP_TRY(choose_tile(H, W, 4, &g));
  g.B = B;
  g.total = total;
  Workspace ws;
  g.Tb = 0;  // direct scatter: no buckets
  EVREP_TRY(carve_checked(workspace, workspace_bytes, B, 0, 1, &ws));
  return launch_voxel(ev, win_offsets, g, ws, flavour, n_bins, normalize, t0_t1_us, divider, out, (cudaStream_t)stream);
}

int evrep_voxel_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                        int B, int H, int W, int flavour, int n_bins, int normalize, const int64_t* t0_t1_us, float* out,
                        void* workspace, size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  return voxel_common(x, y, t, t_bytes, p, win_offsets, B, H, W, flavour, n_bins, normalize, t0_t1_us, 1, out, workspace, workspace_bytes, stream);
  EVREP_GUARD_END
}

int evrep_voxel_subpixel_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                                 int B, int H, int W, int divider, int n_bins, int normalize, const int64_t* t0_t1_us, float* out,
                                 void* workspace, size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  return voxel_common(x, y, t, t_bytes, p, win_offsets, B, H, W, EVREP_VOXEL_EVLICIOUS, n_bins, normalize, t0_t1_us, divider, out, workspace,
                      workspace_bytes, stream);
  EVREP_GUARD_END
}

int evrep_histogram_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                            int B, int H, int W, float* out, void* workspace, size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  Events ev;
  int64_t total = 0, n_max = 0;
  EVREP_TRY(check_events(x, y, t, t_bytes, p, win_offsets, B, out, &ev, &total, &n_max));
  if (B == 0) return EVREP_OK;
  Geom g;
  memset(&g, 0, sizeof(g));
  EVREP_TRY(choose_tile(H, W, 4, &g));
  g.B = B;
  g.total = total;
  Workspace ws;
  g.Tb = 0;  // direct scatter: no buckets
  EVREP_TRY(carve_checked(workspace, workspace_bytes, B, 0, 1, &ws));
  return launch_histogram(ev, win_offsets, g, ws, out, (cudaStream_t)stream);
  EVREP_GUARD_END
}

size_t evrep_gwd_workspace_bytes(const int64_t* s_offsets, const int64_t* t_offsets, int n_pairs) {
  if (!s_offsets || !t_offsets) return 0;
  return gwd_workspace_bytes(s_offsets, t_offsets, n_pairs);
}

int evrep_gwd_kernel_l1(const double* Xs, const int64_t* s_offsets, int ds, const double* Xt, const int64_t* t_offsets, int dt,
                        int n_pairs, double h, double* out, void* workspace, size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (n_pairs < 0 || !s_offsets || !t_offsets) { set_error("bad pair description"); return EVREP_EINVAL; }
  if (n_pairs == 0) return EVREP_OK;
  if (!Xs || !Xt || !out) { set_error("null array"); return EVREP_EINVAL; }
  if (!(h > 0.0)) { set_error("h must be positive"); return EVREP_EINVAL; }
  return launch_gwd(Xs, s_offsets, ds, Xt, t_offsets, dt, n_pairs, h, out, workspace, workspace_bytes, (cudaStream_t)stream);
  EVREP_GUARD_END
}

size_t evrep_gw_kl_workspace_bytes(int n, int m) { return gw_kl_workspace_bytes(n, m); }

int evrep_gw_kl(const double* Xs, int n, int ds, const double* Xt, int m, int dt, double h, int max_iter, double tol_rel, double tol_abs,
                int lmo, double* gw_dist, float* T_out, int* iters, int* lmo_stats, void* workspace, size_t workspace_bytes,
                evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (!Xs || !Xt || !gw_dist) { set_error("null array"); return EVREP_EINVAL; }
  if (!(h > 0.0)) { set_error("h must be positive"); return EVREP_EINVAL; }
  if (max_iter < 0) { set_error("max_iter must be >= 0"); return EVREP_EINVAL; }
  if (lmo != EVREP_LMO_AUCTION && lmo != EVREP_LMO_HOST) { set_error("unknown lmo %d", lmo); return EVREP_EINVAL; }
  return run_gw_kl(Xs, n, ds, Xt, m, dt, h, max_iter, tol_rel, tol_abs, lmo, gw_dist, T_out, iters, lmo_stats, workspace, workspace_bytes,
                   (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_filter_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets, int B,
                         int H, int W, int filter, double param, int fx, int fy, void* state, unsigned char* mask, void* workspace,
                         size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  Events ev;
  int64_t total = 0, n_max = 0;
  EVREP_TRY(check_events(x, y, t, t_bytes, p, win_offsets, B, mask, &ev, &total, &n_max));
  if (filter < EVREP_FILTER_REFRACTORY || filter > EVREP_FILTER_RESIZE) { set_error("unknown filter %d", filter); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  if (!state) { set_error("null state"); return EVREP_EINVAL; }
  if (filter != EVREP_FILTER_RESIZE) fx = fy = 1;
  if (fx < 1 || fy < 1) { set_error("fx, fy must be >= 1"); return EVREP_EINVAL; }
  if (n_max > (int64_t)4096 * SUPER - 16) { set_error("filters: at most %lld events per window", (long long)4096 * SUPER - 16); return EVREP_EUNSUPPORTED; }
  Geom g;
  memset(&g, 0, sizeof(g));
  EVREP_TRY(choose_tile(H, W, 96, &g));  // 1024-pixel tiles: the sort scratch is what fills shared memory
  g.B = B;
  g.total = total;
  g.div_x = fx;
  g.div_y = fy;
  Workspace ws;
  EVREP_TRY(carve_checked(workspace, workspace_bytes, B, total, g.Tb, &ws));
  EVREP_CUDA_OK(cudaMemsetAsync(mask, 0, (size_t)total, (cudaStream_t)stream));
  EVREP_TRY(run_binning(ev, win_offsets, g, ws, REC_IDX, 0, nullptr, (cudaStream_t)stream));
  return launch_filter_tile(g, ws, ev, filter, param, state, mask, (cudaStream_t)stream);
  EVREP_GUARD_END
}

// workspace of the background-activity filter: [binning workspace of the expanded stream][x_e][y_e][t_e][p_e][mask_e]
namespace {
struct BaCarve {
  size_t bin_bytes, off_x, off_y, off_t, off_p, off_m, bytes;
};
bool ba_carve(int B, int64_t total, int H, int W, int radius, int t_bytes, BaCarve* c) {
  if (B < 0 || total < 0 || H < 1 || W < 1 || radius < 1 || radius > 4 || (t_bytes != 4 && t_bytes != 8)) return false;
  const int64_t K = 4 * (int64_t)radius * radius;
  if (total > ((int64_t)1 << 40) / K) return false;
  const int64_t ne = total * K;
  c->bin_bytes = evrep_workspace_bytes(EVREP_OP_FILTER, B, ne, H, W, 1);
  if (!c->bin_bytes) return false;
  size_t o = evrep::align_up(c->bin_bytes, 256);
  c->off_x = o; o = evrep::align_up(o + sizeof(uint16_t) * (size_t)ne, 256);
  c->off_y = o; o = evrep::align_up(o + sizeof(uint16_t) * (size_t)ne, 256);
  c->off_t = o; o = evrep::align_up(o + (size_t)t_bytes * (size_t)ne, 256);
  c->off_p = o; o = evrep::align_up(o + (size_t)ne, 256);
  c->off_m = o; o = evrep::align_up(o + (size_t)ne, 256);
  c->bytes = o;
  return true;
}
}  // namespace

size_t evrep_filter_background_workspace_bytes(int B, int64_t total_events, int H, int W, int radius, int t_bytes) {
  BaCarve c;
  return ba_carve(B, total_events, H, W, radius, t_bytes, &c) ? c.bytes : 0;
}

int evrep_filter_background_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int64_t* win_offsets, int B, int H,
                                    int W, double depth_us, int radius, double* state, unsigned char* mask, void* workspace,
                                    size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  Events ev;
  int64_t total = 0, n_max = 0;
  static const int8_t dummy_p = 0;  // the filter never reads polarities
  EVREP_TRY(check_events(x, y, t, t_bytes, &dummy_p, win_offsets, B, mask, &ev, &total, &n_max));
  if (radius < 1 || radius > 4) { set_error("background-activity filter: radius must be in 1..4, got %d", radius); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  if (!state) { set_error("null state"); return EVREP_EINVAL; }
  const int K = 4 * radius * radius;
  if (n_max * K > (int64_t)4096 * SUPER - 16) {
    set_error("background-activity filter: at most %lld events per window at radius %d", (long long)(((int64_t)4096 * SUPER - 16) / K), radius);
    return EVREP_EUNSUPPORTED;
  }
  BaCarve c;
  if (!ba_carve(B, total, H, W, radius, t_bytes, &c)) { set_error("sensor size %d x %d or batch unsupported", W, H); return EVREP_EINVAL; }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) { set_error("workspace must be non-null and 256-byte aligned"); return EVREP_EWORKSPACE; }
  if (c.bytes > workspace_bytes) { set_error("workspace too small: need %zu bytes, got %zu", c.bytes, workspace_bytes); return EVREP_EWORKSPACE; }
  for (int b = 0; b < B; ++b)
    if (win_offsets[b + 1] < win_offsets[b] || win_offsets[b] < 0) { set_error("win_offsets must be non-decreasing and non-negative"); return EVREP_EINVAL; }
  unsigned char* wsb = (unsigned char*)workspace;
  uint16_t* xe = (uint16_t*)(wsb + c.off_x);
  uint16_t* ye = (uint16_t*)(wsb + c.off_y);
  void* te = wsb + c.off_t;
  int8_t* pe = (int8_t*)(wsb + c.off_p);
  unsigned char* me = wsb + c.off_m;
  const int64_t ne = total * K;
  std::vector<int64_t> offs_e((size_t)B + 1);
  for (int b = 0; b <= B; ++b) offs_e[(size_t)b] = win_offsets[b] * K;
  cudaStream_t st = (cudaStream_t)stream;
  EVREP_TRY(launch_ba_expand(ev, total, H, W, radius, xe, ye, te, pe, st));
  Events eve;
  eve.x = xe; eve.y = ye; eve.t = te; eve.t_bytes = t_bytes; eve.p = pe;
  Geom g;
  memset(&g, 0, sizeof(g));
  EVREP_TRY(choose_tile(H, W, 96, &g));
  g.B = B;
  g.total = ne;
  Workspace ws;
  EVREP_TRY(carve_checked(workspace, c.bin_bytes, B, ne, g.Tb, &ws));
  EVREP_CUDA_OK(cudaMemsetAsync(me, 0, (size_t)ne, st));
  EVREP_CUDA_OK(cudaMemsetAsync(mask, 0, (size_t)total, st));
  EVREP_TRY(run_binning(eve, offs_e.data(), g, ws, REC_IDX, 0, nullptr, st));
  EVREP_TRY(launch_filter_tile(g, ws, eve, EVREP_FILTER_BACKGROUND, depth_us, state, me, st));
  return launch_ba_collect(me, total, K, mask, st);
  EVREP_GUARD_END
}

size_t evrep_est_workspace_bytes(int B) { return B < 0 ? 0 : est_workspace_bytes(B > 0 ? B : 1); }

int evrep_est_quantize_batched(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets, int B, int H,
                               int W, int C, const double* breaks, const double* slope, const double* icpt, int K, float* out, void* workspace,
                               size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (B < 0 || !win_offsets) { set_error("B must be >= 0 and win_offsets non-null"); return EVREP_EINVAL; }
  if (B > 65535) { set_error("at most 65535 windows per call"); return EVREP_EUNSUPPORTED; }
  if (H < 1 || W < 1 || H > 65535 || W > 65535) { set_error("sensor size %d x %d unsupported", W, H); return EVREP_EINVAL; }
  if (C < 2 || C > 64) { set_error("EST needs 2 <= C <= 64 temporal bins (the reference divides by C - 1)"); return EVREP_EINVAL; }
  if (K < 0 || !slope || !icpt || (K > 0 && !breaks)) { set_error("bad piecewise-linear table"); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  if (!out || (win_offsets[B] > 0 && (!x || !y || !t || !p))) { set_error("null array"); return EVREP_EINVAL; }
  return launch_est(x, y, t, p, win_offsets, B, H, W, C, breaks, slope, icpt, K, out, workspace, workspace_bytes, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_est_backward_batched(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets, int B, int H,
                               int W, int C, const double* breaks, int K, const float* grad_out, double* seg_sums, void* workspace,
                               size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (B < 0 || !win_offsets) { set_error("B must be >= 0 and win_offsets non-null"); return EVREP_EINVAL; }
  if (B > 65535) { set_error("at most 65535 windows per call"); return EVREP_EUNSUPPORTED; }
  if (H < 1 || W < 1 || H > 65535 || W > 65535) { set_error("sensor size %d x %d unsupported", W, H); return EVREP_EINVAL; }
  if (C < 2 || C > 64) { set_error("EST needs 2 <= C <= 64 temporal bins"); return EVREP_EINVAL; }
  if (K < 0 || (K > 0 && !breaks) || !seg_sums) { set_error("bad piecewise-linear table / null seg_sums"); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  if (!grad_out || (win_offsets[B] > 0 && (!x || !y || !t || !p))) { set_error("null array"); return EVREP_EINVAL; }
  return launch_est_backward(x, y, t, p, win_offsets, B, H, W, C, breaks, K, grad_out, seg_sums, workspace, workspace_bytes, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_assignment_auction(const float* cost, int n, double eps_rel, int* sigma, int* stats, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (!cost || !sigma || !stats) { set_error("null argument"); return EVREP_EINVAL; }
  return launch_auction(cost, n, eps_rel, sigma, stats, (cudaStream_t)stream);
  EVREP_GUARD_END
}

size_t evrep_unpack_workspace_bytes(int B, int64_t total_events) { return (B < 0 || total_events < 0) ? 0 : unpack_workspace_bytes(B, total_events); }

int evrep_unpack_events(const uint32_t* word, const uint16_t* dt16, const int32_t* tbase, const int64_t* win_offsets, int B, int format, int x_bits,
                        int y_bits, int block_shift, uint16_t* x, uint16_t* y, int32_t* t, int8_t* p, void* workspace, size_t workspace_bytes,
                        evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (B < 0 || !win_offsets) { set_error("B must be >= 0 and win_offsets non-null"); return EVREP_EINVAL; }
  if (B == 0 || win_offsets[B] == 0) return EVREP_OK;
  if (!word || !tbase || !x || !y || !t || !p || (format == 6 && !dt16)) { set_error("null array"); return EVREP_EINVAL; }
  return launch_unpack(word, dt16, tbase, win_offsets, B, format, x_bits, y_bits, block_shift, x, y, t, p, workspace, workspace_bytes, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_transport_plan_host(const float* cost, int n, int m, int cap, int* row_ptr, int* col, double* weight, int* nnz) {
  EVREP_GUARD_BEGIN
  if (!cost || !row_ptr || !col || !weight || cap < 1) { set_error("null argument or cap < 1"); return EVREP_EINVAL; }
  return transport_plan_host(cost, n, m, cap, row_ptr, col, weight, nnz);
  EVREP_GUARD_END
}

size_t evrep_unpack_delta_workspace_bytes(int B) { return B < 0 ? 0 : unpack_delta_workspace_bytes(B); }

int evrep_unpack_events_delta(const uint8_t* rec3, const int32_t* tbase, const uint32_t* esc_prefix, const uint32_t* esc_dt, const int64_t* win_offsets,
                              int B, int x_bits, int y_bits, uint16_t* x, uint16_t* y, int32_t* t, int8_t* p, void* workspace, size_t workspace_bytes,
                              evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (B < 0 || !win_offsets) { set_error("bad argument"); return EVREP_EINVAL; }
  if (B == 0 || win_offsets[B] == 0) return EVREP_OK;
  if (!rec3 || !tbase || !esc_prefix || !x || !y || !t || !p) { set_error("null argument"); return EVREP_EINVAL; }
  return launch_unpack_delta(rec3, tbase, esc_prefix, esc_dt, win_offsets, B, x_bits, y_bits, x, y, t, p, workspace, workspace_bytes, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_warp_affine_batched(const float* img, int B, int C, int in_h, int in_w, const double* M, const int* flips, int out_h, int out_w,
                              const float* border4, int reverse_channels, float scale_out, float* out, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (B < 0 || C < 1 || C > 4096 || in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1 || in_h > 32767 || in_w > 32767 || out_h > 65535 || out_w > 65535) {
    set_error("bad warp geometry");
    return EVREP_EINVAL;
  }
  if (B == 0) return EVREP_OK;
  if (!img || !M || !border4 || !out) { set_error("null argument"); return EVREP_EINVAL; }
  for (int64_t k = 0; k < (int64_t)B * 6; ++k)
    if (!std::isfinite(M[k])) { set_error("warp matrix %lld has a non-finite entry", (long long)(k / 6)); return EVREP_EINVAL; }
  return launch_warp_affine(img, B, C, in_h, in_w, M, flips, out_h, out_w, border4, reverse_channels, scale_out, out, (cudaStream_t)stream);
  EVREP_GUARD_END
}

size_t evrep_otmi_workspace_bytes(int64_t n_events, int rep_size) {
  if (n_events < 0 || rep_size < 1 || rep_size > 65535) return 0;
  return otmi_workspace_bytes((long long)n_events, rep_size);
}

int evrep_otmi_prepare(const void* events, int ev_type, int64_t n_events, const double* rep, int rep_size, int C, int height, int width, double* Xs,
                       int64_t xs_capacity, double* Xt, int64_t xt_capacity, int64_t* info, void* workspace, size_t workspace_bytes,
                       evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (!events || !rep || !Xs || !Xt || !info || !workspace) { set_error("null argument"); return EVREP_EINVAL; }
  if (ev_type < EVREP_OTMI_INT32 || ev_type > EVREP_OTMI_FLOAT64) { set_error("ev_type must be EVREP_OTMI_INT32 / FLOAT32 / FLOAT64"); return EVREP_EINVAL; }
  if (n_events < 0 || rep_size < 1 || rep_size > 65535 || C < 1 || C > 62 || height < 2 || width < 2) { set_error("bad otmi geometry"); return EVREP_EINVAL; }
  const int64_t half = (int64_t)rep_size / 2 + 2;
  if (xs_capacity < n_events || xt_capacity < half * half) {
    set_error("otmi: capacity per point set too small (need %lld event rows and %lld pixel rows)", (long long)n_events, (long long)(half * half));
    return EVREP_EWORKSPACE;
  }
  if (workspace_bytes < otmi_workspace_bytes((long long)n_events, rep_size) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
    set_error("otmi: workspace too small or not 256-byte aligned");
    return EVREP_EWORKSPACE;
  }
  static_assert(sizeof(long long) == sizeof(int64_t), "info layout");
  return launch_otmi_prepare(events, ev_type, (long long)n_events, rep, rep_size, C, height, width, Xs, (long long)xs_capacity, Xt, (long long)xt_capacity,
                             (long long*)info, workspace, (cudaStream_t)stream);
  EVREP_GUARD_END
}

size_t evrep_gemm_workspace_bytes(int M, int N, int K) { return gemm_workspace_bytes(M, N, K); }

int evrep_gemm_nt_3xtf32(const float* A, const float* B, float* C, int M, int N, int K, float alpha, const float* rv, const float* cv,
                         void* workspace, size_t workspace_bytes, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (!A || !B || !C) { set_error("null matrix"); return EVREP_EINVAL; }
  return launch_gemm_nt_3xtf32_ws(A, B, C, M, N, K, alpha, rv, cv, workspace, workspace_bytes, (cudaStream_t)stream);
  EVREP_GUARD_END
}

int evrep_image_pipeline_batched(const float* rep, int B, int H, int W, int C, int img_size, int mode, int interp, float scale_in,
                                 float scale_out, float pad_value, int reverse_channels, float* out, evrep_stream_t stream) {
  EVREP_GUARD_BEGIN
  if (B < 0 || H < 1 || W < 1 || C < 1 || C > 4096 || img_size < 1 || img_size > 65535) { set_error("bad image geometry"); return EVREP_EINVAL; }
  if (B > 65535) { set_error("at most 65535 windows per call"); return EVREP_EUNSUPPORTED; }
  if (mode != EVREP_IMG_LETTERBOX && mode != EVREP_IMG_SQUASH) { set_error("unknown mode %d", mode); return EVREP_EINVAL; }
  if (interp < EVREP_INTERP_AUTO || interp > EVREP_INTERP_LINEAR_TORCH) { set_error("unknown interpolation %d", interp); return EVREP_EINVAL; }
  if (B == 0) return EVREP_OK;
  if (!rep || !out) { set_error("null image"); return EVREP_EINVAL; }
  return launch_image_pipeline(rep, B, H, W, C, img_size, mode, interp, scale_in, scale_out, pad_value, reverse_channels, out, (cudaStream_t)stream);
  EVREP_GUARD_END
}

}  // extern "C"
