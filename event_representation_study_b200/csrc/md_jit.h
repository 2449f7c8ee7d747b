// Run-time specialised mixed-density kernels (md_jit.cu).
#pragma once
#include "evrep_common.cuh"

namespace evrep {

struct JitProgram;

// Is the tuple inside the envelope of the specialised kernels?  Fills the packed plan (buckets below 65536 events), the wide plan
// (hot tiles), the limb width of the wide plan and the tile size.
int md_jit_envelope(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int64_t max_events, MdPlan* packed, MdPlan* wide,
                    int* lw_out, int* tp_out);
// Compiles (once per tuple and limb width) and, with `load`, makes the kernels resident on the current device.  wait = false:
// the compilation runs on a background thread and the call returns at once (md_jit_find sees the program when it is ready).
int md_jit_specialize(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int64_t max_events, bool load, size_t* cubin_bytes,
                      bool wait = true);
// The program a call with this tuple and largest window may use, or nullptr.
JitProgram* md_jit_find(const int8_t* win, const int8_t* func, const int8_t* agg, int C, int stacking, int64_t n_max);
int md_jit_prepare(JitProgram* jp);        // loads the kernels on the current device if they are not there yet
int md_jit_tile_px(const JitProgram* jp);  // 1024 or 512: the tile size its kernels were compiled for
// Light (packed plan, persistent) + heavy (wide plan) kernels on tiles of md_jit_tile_px pixels with polarity-split buckets.
int md_jit_launch(JitProgram* jp, const Geom& g, const Workspace& ws, float* out, cudaStream_t stream);
int md_jit_count();

}  // namespace evrep
