/*
 * libevrep - B200-native event-representation engine: C ABI.
 *
 * This header is the drop-in boundary.  Every entry point takes plain pointers
 * and sizes (no C++ / torch types), enqueues all device work on the caller's stream, allocates nothing
 * persistent (scratch comes from the caller's workspace, sized by evrep_workspace_bytes), overwrites
 * its output completely and returns 0 or a negative EVREP_E* code; evrep_last_error() gives the
 * thread-local message.  Nothing here throws or calls exit().
 *
 * Event batches are SoA + CSR: x,y (uint16), t (int32 or int64 microseconds, see t_bytes), p (int8 in
 * {-1,0,+1}) are DEVICE pointers to `total_events` elements; window b owns
 * [win_offsets[b], win_offsets[b+1]).  `win_offsets` is a HOST pointer to B+1 int64 (the collate step
 * that builds a batch knows the window sizes on the host; the library copies them to the device).
 * Events inside one window must be in stream order (the order the reference's loaders deliver them);
 * representations that are order dependent say so below.
 *
 * Outputs are DEVICE float32, batch-major, in the layout the reference produces for one window.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   evrep_mixed_density_batched  representations/representation_search/mixed_density_event_stack.py:25-46
 *                                (+ operations.py:15-89), one call per window there
 *   evrep_mixed_density_specialize   the same interface for a tuple other than ERGO-12 (any result or candidate of the representation
 *                                search, representation_search/optimization.py:36-64): kernels compiled for that tuple at run time
 *   evrep_ergo12_batched         representations/optimized_representation.py:86-134
 *   evrep_event_stack_batched    representations/event_stack.py:15-63 as called at gen1_transforms.py:33-42
 *   evrep_time_surface_batched   representations/time_surface.py:25-74 as called at gen1_transforms.py:69-87
 *   evrep_tore_batched           representations/tore.py:6-83 as called at gen1_transforms.py:51-67
 *   evrep_voxel_batched          tonic ToVoxelGrid (gen1_transforms.py:21-25);
 *                                ev-licious/src/evlicious/tools/utils.py:51-85 (events_to_voxel_grid);
 *                                representation_search/gromov_wasserstein.py:72-82 (compute_repr)
 *   evrep_histogram_batched      tonic ToImage (gen1_transforms.py:44-49)
 *   evrep_gwd_kernel_l1          representation_search/compute_otmi.py:50-93 (OTMI.__init__ + solve, GWD-A)
 *   evrep_otmi_prepare           representation_search/compute_otmi.py:96-203 (otmi: quadrant split, normalisation, crops -> point sets)
 *   evrep_gw_kl                  representation_search/gromov_wasserstein.py:39-69 (OTMI.__init__ + solve, GWD-B:
 *                                POT ot.gromov.gromov_wasserstein(Ks, Kt, p, q, "kl_loss"))
 *   evrep_gemm_nt_3xtf32         the tensor product constC - hC1 T hC2^T inside that solve (POT ot/gromov tensor_product)
 *   evrep_filter_batched         ev-licious/src/evlicious/tools/utils.py:143-158, 184-200 (_filter_events_resize,
 *                                _contrast_threshold_control, _refractory_period) as used by tools/filters.py:57-109
 *   evrep_filter_background_batched  ev-licious/src/evlicious/tools/utils.py:169-178 (_background_activity_filter) as used by
 *                                tools/filters.py:57-69 (BackgroundActivity.insert)
 *   evrep_est_quantize_batched   ev-YOLOv6/yolov6/models/learned_repr.py:143-172 (QuantizationLayer.forward);
 *   evrep_est_backward_batched   its backward pass with respect to the ValueLayer weights (:9-77 under autograd)
 *   evrep_warp_affine_batched    ev-YOLOv6/yolov6/data/data_augment.py:110-123 (random_affine: cv2.warpAffine) + gen1_2yolo.py:210-228
 *                                (general_augment: flips), the training-time augmentation between letterbox and CHW (gen1_2yolo.py:365-391)
 *   evrep_image_pipeline_batched ev-YOLOv6/yolov6/data/gen1_2yolo.py:230-265,321-341,397 (resize_image, letterbox, CHW + reversal),
 *                                gen4/precompute_reps.py:216-251 (resize_image_process), yolov6/core/engine.py:629-635 (/ 255)
 */
#ifndef EVREP_H
#define EVREP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVREP_VERSION 100 /* 0.1.0 */

/* return codes */
#define EVREP_OK 0
#define EVREP_EINVAL (-1)       /* bad argument */
#define EVREP_EWORKSPACE (-2)   /* workspace too small / misaligned */
#define EVREP_ECUDA (-3)        /* a CUDA runtime call failed */
#define EVREP_EUNSUPPORTED (-4) /* valid request outside the implemented envelope */

/* ops, for evrep_workspace_bytes */
#define EVREP_OP_MIXED_DENSITY 1
#define EVREP_OP_EVENT_STACK 2
#define EVREP_OP_TIME_SURFACE 3
#define EVREP_OP_TORE 4
#define EVREP_OP_VOXEL 5
#define EVREP_OP_HISTOGRAM 6
#define EVREP_OP_FILTER 7

/* MixedDensityEventStack vocabulary (operations.py:39-89; mixed_density_event_stack.py:48-109) */
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
#define EVREP_AGG_MIN 4 /* torch_scatter reduce="min" (untouched pixels 0): Operations passes the string straight to scatter (operations.py:30-35); used by the N-ImageNet scatter_min planes (imagenet.py:241-244, 383-386) */
#define EVREP_STACK_SBN 0 /* windows by number of events (the one ERGO-12 uses) */
#define EVREP_STACK_SBT 1 /* windows by time */
#define EVREP_MAX_CHANNELS 32

/* voxel-grid flavours */
#define EVREP_VOXEL_TONIC 0     /* (B, n_bins, H, W): bilinear in time over [t_first, t_last], p==0 -> -1 */
#define EVREP_VOXEL_EVLICIOUS 1 /* (B, n_bins, H, W): floor-bin polarity histogram (+ optional normalisation) */
#define EVREP_VOXEL_GWD 2       /* (B, H, W, n_bins): compute_repr, bilinear over t in [0,1] = (t - t_first)/(t_last - t_first) */

/* per-window status bits written by every batched op (read back with evrep_window_flags) */
#define EVREP_WF_OUT_OF_RANGE 0x100u /* an event had x >= W or y >= H; it was dropped */
#define EVREP_WF_UNSORTED 0x200u     /* timestamps decrease somewhere inside the window */
#define EVREP_WF_T_RANGE 0x400u      /* |t - t_first| does not fit 31 bits; the event was dropped (EventStack and the filters do
                                        not key on time: there the flag is only informative and every event is kept) */
#define EVREP_WF_BAD_POLARITY 0x800u /* p outside {-1,0,1}; treated as sign(p) */

/* kernels of the tile pipeline, for evrep_profile_read */
#define EVREP_K_COUNT 0 /* bucket sizes (binning pass 1) */
#define EVREP_K_SCAN 1  /* bucket starts */
#define EVREP_K_BIN 2   /* scatter into tile buckets (binning pass 2) */
#define EVREP_K_TILE 3  /* per-tile reduction + finalise: the kernel that writes the output */
#define EVREP_K_N 4

typedef void* evrep_stream_t; /* a cudaStream_t */

int evrep_version(void);
const char* evrep_last_error(void);

/* Per-kernel device timing for benchmarks.  evrep_profile_enable(n) makes the next n batched calls record
 * CUDA events around their kernels on the caller's stream (no synchronisation, a few microseconds of
 * stream time per call); n = 0 switches it off and frees the events.  evrep_profile_read synchronises on
 * the recorded events and returns, for one kernel, the summed milliseconds and the number of launches
 * timed since the last enable.  The state is per host thread (like evrep_last_error): a thread times its own calls and
 * never sees another thread's; benchmark use only. */
int evrep_profile_enable(int max_calls);
int evrep_profile_read(int kernel_id, float* total_ms, int* launches);

/* Upper bound of the scratch an op needs for a batch of B windows with total_events events on an
 * H x W sensor producing C channels.  0 on invalid arguments. */
size_t evrep_workspace_bytes(int op, int B, int64_t total_events, int H, int W, int C);

/* Copies the B per-window status words (EVREP_WF_*) of the last op that used `workspace` to the host
 * and synchronises `stream`. */
int evrep_window_flags(const void* workspace, int B, uint32_t* flags_host, evrep_stream_t stream);

/* Host-only: the shared-memory plan the mixed-density tile kernel would use.  Writes
 * info[0]=bytes of accumulator per pixel, info[1]=pixels per tile, info[2]=tiles per window,
 * info[3]=number of accumulator words, info[4]=dynamic shared memory bytes per CTA. */
int evrep_mixed_density_plan_info(int H, int W, const int8_t* win, const int8_t* func, const int8_t* agg, int C,
                                  int stacking, int64_t max_events_per_window, int* info);

/* out: (B, H, W, C) float32.  win/func/agg: HOST arrays of C entries.  Window indices 0..6 (SBN) or
 * 0..7 (SBT); an index outside that range yields an all-zero channel like the reference's
 * swallow-and-zero.  A window with zero events yields zeros. */
int evrep_mixed_density_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                                const int64_t* win_offsets, int B, int H, int W, const int8_t* win, const int8_t* func,
                                const int8_t* agg, int C, int stacking, float* out, void* workspace,
                                size_t workspace_bytes, evrep_stream_t stream);

/* Run-time specialisation of evrep_mixed_density_batched for ONE tuple - for running a tuple other than ERGO-12 at scale: a
 * representation found by the search (mixed_density_event_stack.py:25-46 with the tuple optimization.py:36-64 assembles) used to
 * train or evaluate a model over a dataset, or a search that scores candidates on large batches.  (The reference's own search
 * scores a candidate on two samples, optimization.py:131-141: that stays on the interpreted kernel and loses nothing.)
 * The two ERGO-12 tuples have kernels built ahead of time; any other tuple runs an interpreted kernel that is several times
 * slower.  This call compiles the same kernel templates for the given tuple with NVRTC (sm_100a; a few seconds, once per
 * tuple and process) and makes them resident on the current device; from then on evrep_mixed_density_batched launches them
 * whenever it is called with exactly this tuple (same results within the documented tolerance; integer channels bit exact).
 * max_events_per_window bounds the largest window of later calls (it picks the limb width of the hot-tile kernel; a later
 * call with larger windows silently uses the interpreted kernel).  Envelope: SBN or SBT stacking, per-pixel accumulators that fit
 * a tile's shared memory (at most 55 packed words: every 12-channel tuple of the reference's vocabulary does); at call time a
 * sensor with H * W * C a multiple of 4 and at most 2 Mpx (1 Mpx for tuples above 27 words, which take 512-pixel tiles).
 * EVREP_EUNSUPPORTED otherwise (and when libnvrtc / the driver library cannot be opened) - the interpreted
 * kernel keeps serving such tuples.  Host call; thread safe. */
int evrep_mixed_density_specialize(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking,
                                   int64_t max_events_per_window);
/* The same, without waiting: the compilation runs on a background host thread and the call returns at once; calls of
 * evrep_mixed_density_batched keep using the interpreted kernel until the program is ready and switch to the specialised
 * kernels from then on (results agree within the documented tolerance, integer-valued channels exactly).  For streams of
 * per-window calls that should never stall (the drop-in class uses it after 64 calls with one tuple). */
int evrep_mixed_density_specialize_async(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking,
                                         int64_t max_events_per_window);
/* The same compilation without loading anything on a device (works on a machine without a GPU): *cubin_bytes receives the
 * size of the sm_100a image.  Used by the build check and the CPU tests. */
int evrep_mixed_density_specialize_compile_only(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking,
                                                int64_t max_events_per_window, size_t* cubin_bytes);
/* 1 when a call with this tuple and largest window would run specialised kernels, else 0. */
int evrep_mixed_density_is_specialized(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking,
                                       int64_t max_events_per_window);

/* ERGO-12: version 2 = the active tuple (optimized_representation.py:86-115), 1 = the commented one (:16-66). */
int evrep_ergo12_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                         const int64_t* win_offsets, int B, int H, int W, int version, float* out, void* workspace,
                         size_t workspace_bytes, evrep_stream_t stream);

/* out: (B, H, W, stack_size) float32 in {-1,0,1}: polarity sign (p > 0 -> +1 else -1) of the latest
 * event of the pixel if its index is inside nested suffix window k, else 0. */
int evrep_event_stack_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                              const int64_t* win_offsets, int B, int H, int W, int stack_size, float* out,
                              void* workspace, size_t workspace_bytes, evrep_stream_t stream);

/* out: (B, S, 2, H, W) float32.  indices: HOST int64 (B*S) snapshot event indices, or NULL for the
 * gen1_transforms.py:78-80 rule (searchsorted of S equal time steps).  Polarity plane = p > 0.
 * Requires time-sorted windows (EVREP_WF_UNSORTED is raised otherwise). */
int evrep_time_surface_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                               const int64_t* win_offsets, int B, int H, int W, const int64_t* indices, int S,
                               double tau, float* out, void* workspace, size_t workspace_bytes,
                               evrep_stream_t stream);

/* out: (B, H, W, 2k) float32.  Sample time = timestamp of the last event of each window; events with
 * t >= sample time are ignored; channels [0,k) = k most recent ages of p > 0 events ascending, [k,2k)
 * the same for p <= 0; log compression of tore.py:69-79.  Pixels are the 0-based x,y. */
int evrep_tore_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                       const int64_t* win_offsets, int B, int H, int W, int k, float* out, void* workspace,
                       size_t workspace_bytes, evrep_stream_t stream);

/* BASELINE configs[2] in one call: EventStack (stack_size 12), TimeSurface (6 snapshots by the gen1_transforms.py rule, tau)
 * and TORE (k = 6) of the same windows from a single bucketing pass.  out_es: (B, H, W, 12), out_ts: (B, 6, 2, H, W),
 * out_tore: (B, H, W, 12); each is bit-identical to what the separate entry points write.  Windows of fewer than 2^20
 * events (EVREP_EUNSUPPORTED otherwise).  Workspace: evrep_workspace_bytes(EVREP_OP_TORE, ...). */
int evrep_order_ops_fused_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                                  const int64_t* win_offsets, int B, int H, int W, double tau, float* out_es, float* out_ts,
                                  float* out_tore, void* workspace, size_t workspace_bytes, evrep_stream_t stream);

/* flavour = EVREP_VOXEL_*.  normalize and t0_t1_us (HOST, 2 entries: the t0_us / t1_us arguments of
 * events_to_voxel_grid, or NULL for first / last timestamp of each window) apply to the ev-licious flavour
 * only.  A window with fewer than 2 events gives zeros in the ev-licious flavour (utils.py:52-53). */
int evrep_voxel_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                        const int64_t* win_offsets, int B, int H, int W, int flavour, int n_bins, int normalize,
                        const int64_t* t0_t1_us, float* out, void* workspace, size_t workspace_bytes,
                        evrep_stream_t stream);

/* The ev-licious voxel grid for events with SUB-PIXEL coordinates (Events.divider > 1, events.py:37-47: the event sits at
 * (x / divider, y / divider) in float32; utils.py:70-76, 93-103): x, y hold the raw integers, H x W is the grid (the scaled
 * sensor), every event is spread over the four pixels around it with bilinear weights, taps outside the grid are dropped;
 * time binning, normalize and t0_t1_us as in the ev-licious flavour of evrep_voxel_batched.  out: (B, n_bins, H, W).
 * Float reductions: reproducible to the last ulp, not bit for bit. */
int evrep_voxel_subpixel_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                                 const int64_t* win_offsets, int B, int H, int W, int divider, int n_bins, int normalize,
                                 const int64_t* t0_t1_us, float* out, void* workspace, size_t workspace_bytes,
                                 evrep_stream_t stream);

/* out: (B, 2, H, W) float32 event counts, plane 0 = p <= 0, plane 1 = p > 0. */
int evrep_histogram_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                            const int64_t* win_offsets, int B, int H, int W, float* out, void* workspace,
                            size_t workspace_bytes, evrep_stream_t stream);

/* GWD-A for n_pairs independent (Xs, Xt) pairs: out[i] = mean |pad(Ks) - pad(Kt)| with Gaussian
 * kernels of bandwidth h * std (compute_otmi.py:6-32, 61-93).  Xs: DEVICE float64 rows packed back to
 * back, pair i = rows [s_offsets[i], s_offsets[i+1]) of width ds; same for Xt with dt.  s_offsets /
 * t_offsets: HOST int64 (n_pairs+1).  out: DEVICE float64 (n_pairs).  workspace >=
 * evrep_gwd_workspace_bytes(s_offsets, t_offsets, n_pairs), 256-byte aligned.  ds, dt <= 64. */
size_t evrep_gwd_workspace_bytes(const int64_t* s_offsets, const int64_t* t_offsets, int n_pairs);
int evrep_gwd_kernel_l1(const double* Xs, const int64_t* s_offsets, int ds, const double* Xt, const int64_t* t_offsets,
                        int dt, int n_pairs, double h, double* out, void* workspace, size_t workspace_bytes,
                        evrep_stream_t stream);

/* The data preparation of otmi() (compute_otmi.py:96-203) for ONE sample: `events` (DEVICE, n_events x 4 row major [x, y, t,
 * p]; ev_type EVREP_OTMI_INT32 - what the reference's caller passes, a torch int32 tensor (gen1_compute.py:57-59): exact
 * integer differences, float32 quotients like torch's int / int - or EVREP_OTMI_FLOAT32 / FLOAT64: the array's own precision) are split
 * into the four sensor quadrants, the densest is dropped (first maximum), the other three are rebased (not the first), scaled
 * by (W - 1) // 2 and (H - 1) // 2, their t and p normalised over the quadrant and the rows with rebased x, y below those
 * divisors kept, in stream order -> Xs, three point sets of 4 float64 columns; `rep` (DEVICE float64, rep_size x rep_size x
 * C, HWC) is cropped per quadrant (the reference's int() of its float bounds), two positional channels i / (a - 1), j /
 * (b - 1) are appended and the pixels whose C channels are all zero dropped, in row-major order -> Xt, three point sets of C
 * + 2 columns.  Point set s (s = 0..2: the kept quadrants in quadrant order) starts at Xs + s * xs_capacity * 4 and Xt +
 * s * xt_capacity * (C + 2); capacities are in rows (xs_capacity >= n_events, xt_capacity >= (rep_size / 2 + 2)^2).
 * info (HOST, 12 int64): [0..2] rows of the three Xs, [3..5] rows of the three Xt, [6] the dropped quadrant, [7] 1 + index of
 * an empty quadrant among 1..3 (the reference raises on it; 0 if none), [8..11] events per quadrant before the final mask.
 * The call SYNCHRONISES the stream (the caller needs the row counts).  The point sets feed evrep_gwd_kernel_l1 directly. */
#define EVREP_OTMI_INT32 0
#define EVREP_OTMI_FLOAT32 1
#define EVREP_OTMI_FLOAT64 2
size_t evrep_otmi_workspace_bytes(int64_t n_events, int rep_size);
int evrep_otmi_prepare(const void* events, int ev_type, int64_t n_events, const double* rep, int rep_size, int C, int height, int width,
                       double* Xs, int64_t xs_capacity, double* Xt, int64_t xt_capacity, int64_t* info, void* workspace,
                       size_t workspace_bytes, evrep_stream_t stream);

/* GWD-B for ONE pair: builds the Gaussian kernels Ks (n x n) and Kt (m x m) of Xs (n x ds) and Xt (m x dt) (DEVICE
 * float64, row major, bandwidth h * std as in compute_kernel), then runs conditional-gradient Gromov-Wasserstein with
 * the KL loss from T = p q^T (uniform p, q) until the loss changes by less than tol_abs or tol_rel (POT: 1e-9, 1e-9;
 * both are floored at 4 float32 ulp of the loss, the resolution of the float32 gradient) or max_iter (POT: 10000) steps.  Writes the loss at the final plan to *gw_dist (HOST), the iteration count to *iters
 * (HOST, may be NULL) and the plan to T_out (DEVICE float32 n x m, may be NULL).  The dense contraction of every step
 * runs on the tensor cores (evrep_gemm_nt_3xtf32).  The linear minimisation oracle (an assignment problem) is, with lmo =
 * EVREP_LMO_AUCTION, Bertsekas' forward auction with epsilon scaling on the GPU (optimal up to n * 1e-9 * cost range; n <=
 * 4096; a step that hits the round limit is redone by the host solver) or, with EVREP_LMO_HOST, an exact shortest-
 * augmenting-path solve on the host.  Either way a few scalars come back every iteration, so the call SYNCHRONISES the
 * stream.  lmo_stats (HOST, 3 ints, may be NULL): auction rounds, auction bids, steps solved on the host.  For n != m (the
 * reference's own call shape: n events against m pixels) every vertex is a transportation plan with up to n + m - 1
 * entries: the LMO is then always evrep_transport_plan_host (the gradient is read back, the plan uploaded in CSR form)
 * and the step's contraction is hC1 (Gc hC2^T) with the sparse product formed by a small kernel.  ds, dt <= 64. */
#define EVREP_LMO_AUCTION 0
#define EVREP_LMO_HOST 1
size_t evrep_gw_kl_workspace_bytes(int n, int m);
int evrep_gw_kl(const double* Xs, int n, int ds, const double* Xt, int m, int dt, double h, int max_iter, double tol_rel,
                double tol_abs, int lmo, double* gw_dist, float* T_out, int* iters, int* lmo_stats, void* workspace,
                size_t workspace_bytes, evrep_stream_t stream);

/* The LMO of evrep_gw_kl on its own: min-cost assignment of the n x n DEVICE float32 matrix `cost` (row major) by a
 * forward auction with epsilon scaling.  sigma (DEVICE, n ints): column assigned to every row; the total cost is within
 * n * eps_rel * (max cost - min cost) of the optimum.  stats (DEVICE, 3 ints): rounds, bids, status (0 ok, 1 round limit,
 * 2 non-finite costs; sigma is not a permutation unless status is 0).  n <= 4096.  Enqueued on `stream`, no sync. */
int evrep_assignment_auction(const float* cost, int n, double eps_rel, int* sigma, int* stats, evrep_stream_t stream);

/* Packed host wire format (the end-to-end path is bound by the host link: 9 B/event of SoA arrays at ~54 GB/s).  The loader
 * packs an event into one 32-bit word  x | y << x_bits | (p & 3) << (x_bits + y_bits) | dt << (x_bits + y_bits + 2)  with
 * dt = t - tbase[block], blocks of 2^block_shift consecutive events of a window, block bases int32 relative to the
 * window's first timestamp (format 4: 4 B/event; needs every dt < 2^(30 - x_bits - y_bits)), or with dt in a separate uint16
 * array (format 6: 6 B/event, dt < 65536).  evrep_unpack_events decodes a batch into the SoA arrays every other entry
 * point takes: x, y (uint16), t (int32, equal to the original timestamps up to one constant per window - representations
 * only use differences inside a window), p (int8 in {-1, 0, +1}).  word, dt16 (format 6 only, else NULL), tbase (one entry
 * per block, windows concatenated: window b owns ceil(n_b / 2^block_shift) blocks) and the outputs are DEVICE pointers;
 * win_offsets is HOST (B + 1); workspace: DEVICE, evrep_unpack_workspace_bytes.  Enqueued on `stream`.
 * Python side: event_representation_study_b200.packed (pack_host / upload). */
size_t evrep_unpack_workspace_bytes(int B, int64_t total_events);
int evrep_unpack_events(const uint32_t* word, const uint16_t* dt16, const int32_t* tbase, const int64_t* win_offsets, int B, int format,
                        int x_bits, int y_bits, int block_shift, uint16_t* x, uint16_t* y, int32_t* t, int8_t* p, void* workspace,
                        size_t workspace_bytes, evrep_stream_t stream);

/* Format 3 of the packed host wire format: 3 bytes per event, for time-sorted windows with polarities in {-1, +1} on sensors
 * with x_bits + y_bits <= 21 (1280 x 720: 11 + 10).  record = x | y << x_bits | (p > 0) << (x_bits + y_bits) | code <<
 * (x_bits + y_bits + 1), little endian, where code is the event's timestamp MINUS ITS PREDECESSOR'S (0, 1, 2) or 3 = "the
 * difference is the next entry of esc_dt".  Events are grouped in blocks of 64 of one window; a block occupies exactly 192
 * bytes of rec3 (the last block of a window is padded with anything), tbase[block] is the timestamp of its first event
 * relative to the window's first (that event's code is 0), esc_prefix[block] the number of escapes in the blocks before it
 * (any common offset is removed: the tables of a group of windows may be slices of a batch's), esc_dt the escape values in
 * stream order.  All DEVICE pointers except win_offsets (HOST, B + 1); x, y, t, p as in evrep_unpack_events.  3.13 B/event
 * on the host link against 4.06 (format 4) and 9 (SoA arrays).  Lossless (tests/test_packed.py). */
size_t evrep_unpack_delta_workspace_bytes(int B);
int evrep_unpack_events_delta(const uint8_t* rec3, const int32_t* tbase, const uint32_t* esc_prefix, const uint32_t* esc_dt,
                              const int64_t* win_offsets, int B, int x_bits, int y_bits, uint16_t* x, uint16_t* y, int32_t* t, int8_t* p,
                              void* workspace, size_t workspace_bytes, evrep_stream_t stream);

/* The LMO of evrep_gw_kl for rectangular plans (n != m): an optimal vertex of the transportation problem
 *     min <cost, G>  s.t.  G 1 = 1/n,  G^T 1 = 1/m,  G >= 0
 * (what POT's ot.emd returns inside ot.gromov.gromov_wasserstein; gromov_wasserstein.py:62-69).  HOST function, HOST
 * pointers, no GPU involved: `cost` is n x m float32 row major; the plan comes back in CSR form - row_ptr (n + 1), col and
 * weight (capacity `cap`, n + m suffices unless costs tie), *nnz entries (may be NULL).  Exact: successive shortest paths
 * with integer flows (unit gcd(n, m) / (n m)); EVREP_EWORKSPACE when the plan does not fit `cap`, EVREP_EINVAL for
 * non-finite costs. */
int evrep_transport_plan_host(const float* cost, int n, int m, int cap, int* row_ptr, int* col, double* weight, int* nnz);

/* C[M x N] = alpha * A[M x K] * B[N x K]^T + rv[i] + cv[j] on the tcgen05 tensor cores: every fp32 operand is split
 * into two TF32 terms (round to nearest) and lo*hi + hi*lo + hi*hi is accumulated, k-blocks of 32 summed in fp32 with
 * round-to-nearest (error about 2^-22 relative to |A| |B|).  All pointers DEVICE float32, row major, K contiguous in A
 * and B; rv (M) and cv (N) may be NULL.  With a workspace of evrep_gemm_workspace_bytes(M, N, K) (DEVICE, 256-byte
 * aligned) the operands are first packed into swizzled tile images and the kernel is fed by TMA bulk copies; with
 * workspace == NULL a slower variant stages the operands through registers. */
size_t evrep_gemm_workspace_bytes(int M, int N, int K);
int evrep_gemm_nt_3xtf32(const float* A, const float* B, float* C, int M, int N, int K, float alpha, const float* rv,
                         const float* cv, void* workspace, size_t workspace_bytes, evrep_stream_t stream);

/* The image pipeline that follows a representation in the detector's datasets, fused: out = letterbox(resize(rep *
 * scale_in)) * scale_out, HWC -> CHW, optionally with the channel order reversed (`img.transpose(2, 0, 1)[::-1]`).
 * rep: DEVICE float32 (B, H, W, C); out: DEVICE float32 (B, C, img_size, img_size).
 * mode EVREP_IMG_LETTERBOX: gen1_2yolo.py - resize to (int(W r), int(H r)), r = img_size / max(H, W), then pad to the
 * square with pad_value (114) like letterbox(auto=False, scaleup=False); mode EVREP_IMG_SQUASH: precompute_reps.py -
 * resize straight to img_size x img_size.  interp EVREP_INTERP_AUTO follows the reference (INTER_AREA when r < 1, the
 * non-augmented branch, else INTER_LINEAR); the arithmetic follows cv::resize on float images tap for tap.
 * INTER_AREA is implemented for shrinking axes at any factor (tap tables up to 4, taps generated on the fly beyond); an
 * enlarging axis, where cv::resize switches to a linear variant, returns EVREP_EUNSUPPORTED. */
#define EVREP_IMG_LETTERBOX 0
#define EVREP_IMG_SQUASH 1
#define EVREP_INTERP_AUTO 0
#define EVREP_INTERP_LINEAR 1
#define EVREP_INTERP_AREA 2
#define EVREP_INTERP_LINEAR_TORCH 3 /* torch interpolate(bilinear, align_corners=False) arithmetic + letterbox_image_batch placement (learned_repr.py:94-141) */
int evrep_image_pipeline_batched(const float* rep, int B, int H, int W, int C, int img_size, int mode, int interp, float scale_in,
                                 float scale_out, float pad_value, int reverse_channels, float* out, evrep_stream_t stream);

/* The training-time augmentation between the letterbox and the CHW transpose (gen1_2yolo.py:365-391): random_affine's
 * cv2.warpAffine(img, M[:2], dsize=(out_w, out_h), borderValue=(114, 114, 114)) (data_augment.py:110-123), then the flips of
 * general_augment (gen1_2yolo.py:210-228).  img: DEVICE float32 (B, C, in_h, in_w) planes - what evrep_image_pipeline_batched
 * writes with reverse_channels = 0 and scale_out = 1; out: DEVICE float32 (B, C, out_h, out_w) = warped, flipped, channel
 * order reversed if asked (the `[::-1]` of gen1_2yolo.py:397), times scale_out (1 / 255).  M: HOST, B x 6 doubles, the FORWARD
 * 2 x 3 matrices as handed to cv2.warpAffine (the caller draws them: get_transform_matrix uses Python's random); flips: HOST,
 * B ints, bit 0 = up-down, bit 1 = left-right, may be NULL; border4: HOST, 4 floats - cv2 turns the 3-tuple into the scalar
 * (114, 114, 114, 0) and channel k takes entry k & 3, so pass exactly that to reproduce the reference.  The arithmetic follows
 * cv::warpAffine INTER_LINEAR on float images: double inverse map, fixed-point destination coordinates rounded to 1/32 pixel,
 * float 32 x 32 weight table, double accumulation. */
int evrep_warp_affine_batched(const float* img, int B, int C, int in_h, int in_w, const double* M, const int* flips, int out_h, int out_w,
                              const float* border4, int reverse_channels, float scale_out, float* out, evrep_stream_t stream);

/* ev-licious' per-pixel stateful filters over B event windows (streams): mask[i] = 1 if event i passes, 0 otherwise
 * (DEVICE uint8, one per event, fully overwritten; events with x or y outside the sensor get 0 and raise
 * EVREP_WF_OUT_OF_RANGE).  `state` is DEVICE memory, (B, H, W) read AND written, so that a stream can be fed in pieces:
 *   EVREP_FILTER_REFRACTORY  float64 last kept timestamp per pixel (start at -inf); param = period.  Keeps an event iff
 *                            t - state >= period, then state = t (utils.py:193-200).
 *   EVREP_FILTER_CONTRAST    int32 activity per pixel (start at 0); param = factor.  activity += p; keeps the event and
 *                            resets the pixel when |activity| >= factor (utils.py:184-191).  p must be -1 / +1.
 *   EVREP_FILTER_RESIZE      float32 change map per CELL of fx x fy pixels; here H, W are the cell grid (height, width of
 *                            the change map) and x / fx, y / fy index it.  change += p / (fx fy); keeps the event when
 *                            |change| >= 1 and then subtracts p (utils.py:143-158).  fx, fy >= 1 (ignored by the others).
 * Events of a window must be in stream order.  Windows of at most 33 M events. */
#define EVREP_FILTER_REFRACTORY 0
#define EVREP_FILTER_CONTRAST 1
#define EVREP_FILTER_RESIZE 2
int evrep_filter_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p, const int64_t* win_offsets,
                         int B, int H, int W, int filter, double param, int fx, int fy, void* state, unsigned char* mask,
                         void* workspace, size_t workspace_bytes, evrep_stream_t stream);

/* ev-licious' background-activity filter (utils.py:169-178): for every event in stream order, t_last = state[y, x];
 * mask = !(t_last > 0 && t - t_last > depth_us); then state[y - radius .. y + radius - 1, x - radius .. x + radius - 1] = t
 * (clipped to the sensor, like the numpy slice).  `state` is DEVICE float64 (B, H, W), read and written (start at -inf,
 * tools/filters.py:64-65), so a stream can be fed in pieces; `mask` DEVICE uint8, one per event, fully overwritten (events
 * outside the sensor get 0, write nothing and raise EVREP_WF_OUT_OF_RANGE).  radius in 1..4.  The stream is expanded into
 * (2 radius)^2 per-pixel write records inside the workspace (evrep_filter_background_workspace_bytes), so a window may
 * hold at most 33 M / (2 radius)^2 events.  The filter id below is only valid through this entry point. */
#define EVREP_FILTER_BACKGROUND 3
size_t evrep_filter_background_workspace_bytes(int B, int64_t total_events, int H, int W, int radius, int t_bytes);
int evrep_filter_background_batched(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int64_t* win_offsets, int B,
                                    int H, int W, double depth_us, int radius, double* state, unsigned char* mask, void* workspace,
                                    size_t workspace_bytes, evrep_stream_t stream);

/* EST, the reference's learned quantisation layer, forward pass (ev-YOLOv6/yolov6/models/learned_repr.py:143-172):
 * out[b, y, x, p * C + i] = sum over the events of window b at (x, y, p) of tn * f(tn - i / (C - 1)), tn = t / max(t of the
 * window) in float32, f = the ValueLayer MLP given as the piecewise-linear function it is: `breaks` (K sorted float64
 * breakpoints), `slope` and `icpt` (K + 1 float64 each; segment j covers breaks[j-1] <= u < breaks[j]) - compile them from
 * the weights with event_representation_study_b200/est.py::compile_value_layer.  All three are DEVICE pointers.
 * x, y, p as everywhere; t: DEVICE float32 (the reference's event tensor is float).  out: DEVICE float32 (B, H, W, 2C),
 * i.e. torch.cat([vox[:, 0], vox[:, 1]], 1) in HWC.  p > 0 selects the second half.  C >= 2.  Accumulation uses float
 * atomics like the reference's put_(accumulate=True): reproducible to rounding only. */
size_t evrep_est_workspace_bytes(int B);
int evrep_est_quantize_batched(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets, int B,
                               int H, int W, int C, const double* breaks, const double* slope, const double* icpt, int K, float* out,
                               void* workspace, size_t workspace_bytes, evrep_stream_t stream);

/* Backward pass of the EST layer with respect to the ValueLayer weights (training: learned_repr.py:9-77, 143-172 under
 * autograd).  Inside segment j the layer is f(u) = a_j u + c_j, so dL/dtheta = sum_j (G1_j da_j/dtheta + G0_j dc_j/dtheta)
 * with G0_j = sum g, G1_j = sum g u over the (event, bin) samples of segment j and g = grad_out[b, y, x, p C + i] * tn.
 * seg_sums (DEVICE float64, 2 (K + 1) entries: G0 then G1) is overwritten; est.py::EstQuantize turns it into parameter
 * gradients.  Same events, tables and workspace as the forward call; grad_out: DEVICE float32 (B, H, W, 2C). */
int evrep_est_backward_batched(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets, int B,
                               int H, int W, int C, const double* breaks, int K, const float* grad_out, double* seg_sums,
                               void* workspace, size_t workspace_bytes, evrep_stream_t stream);

/* Host-side encoder of packed wire format 3, the loader's half of evrep_unpack_events_delta (no counterpart in the reference: its
 * loaders slice, concatenate and cast the event arrays per sample in Python, ev-YOLOv6/yolov6/data/gen1_2yolo.py:186-208).  Every
 * pointer is a HOST pointer.  x, y, t (t_bytes 4 or 8), p: SoA events of B windows (win_offsets, B + 1).  Outputs, caller
 * allocated: rec3 (192 bytes per block of 64 events; evrep_pack_delta_host_blocks gives the block count), tbase (one int32 per
 * block), esc_prefix (blocks + 1 entries) and esc_dt (esc_capacity entries; the number needed comes back in *n_escapes -
 * EVREP_EWORKSPACE when it does not fit, call again with a larger table).  One fused pass over the events on n_threads host
 * threads (< 1: as many as the machine has, at most 16); byte-identical to packed.py's numpy packer.  EVREP_EUNSUPPORTED when
 * the stream does not fit the format (a window not time sorted inside a block, a polarity other than -1 / +1, x and y needing
 * more than 21 bits, a window spanning 2^31 us): ship formats 4 / 6 or the SoA arrays instead.  zero_is_negative != 0 also
 * accepts p == 0 and writes it like p == -1 (the decoder then returns -1): for {0, 1} streams whose consumers treat 0 and -1
 * alike - every representation of this library does when a window holds no -1 (operations.py:59-61, the p > 0 tests of
 * event_stack.py / time_surface.py) - at the price of not being able to tell them apart afterwards. */
int64_t evrep_pack_delta_host_blocks(const int64_t* win_offsets, int B);
/* The same for wire formats 4 and 6 (the loader's half of evrep_unpack_events): word (one uint32 per event), dt16 (one uint16 per
 * event, format 6 only) and tbase (one int32 per block of 64 / 256 events; evrep_pack_host_blocks gives the count).  Any event
 * order, polarities in {-1, 0, 1}; EVREP_EUNSUPPORTED when a block spans more time than the format holds
 * (2^(30 - bits(W) - bits(H)) us in format 4, 65536 us in format 6). */
int64_t evrep_pack_host_blocks(const int64_t* win_offsets, int B, int fmt);
int evrep_pack_events_host(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                           const int64_t* win_offsets, int B, int H, int W, int fmt, uint32_t* word, uint16_t* dt16,
                           int32_t* tbase, int n_threads);
int evrep_pack_events_delta_host(const uint16_t* x, const uint16_t* y, const void* t, int t_bytes, const int8_t* p,
                                 const int64_t* win_offsets, int B, int H, int W, uint8_t* rec3, int32_t* tbase,
                                 uint32_t* esc_prefix, uint32_t* esc_dt, int64_t esc_capacity, int64_t* n_escapes, int zero_is_negative,
                                 int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* EVREP_H */
