"""ctypes binding of libevrep.so (include/evrep.h).

There is no fallback: if the shared library is missing or a symbol cannot be resolved the import of
this module raises, and every compute entry point needs a CUDA device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EVREP_LIB selects another build of the same library (A/B tuning builds made with EVREP_NVCC_FLAGS / EVREP_LIB_OUT); it
# must exist - there is still no fallback
LIB_PATH = os.environ.get("EVREP_LIB") or os.path.join(_HERE, "lib", "libevrep.so")

# return codes (include/evrep.h)
OK, EINVAL, EWORKSPACE, ECUDA, EUNSUPPORTED = 0, -1, -2, -3, -4
MAX_TILES = 4096  # csrc/evrep_common.cuh: buckets per window
OP_MIXED_DENSITY, OP_EVENT_STACK, OP_TIME_SURFACE, OP_TORE, OP_VOXEL, OP_HISTOGRAM, OP_FILTER = 1, 2, 3, 4, 5, 6, 7
FUNCS = {"timestamp": 0, "polarity": 1, "count": 2, "timestamp_pos": 3, "timestamp_neg": 4, "count_pos": 5, "count_neg": 6}
AGGS = {"sum": 0, "mean": 1, "max": 2, "variance": 3, "min": 4}
STACKING = {"SBN": 0, "SBT": 1}
VOXEL_TONIC, VOXEL_EVLICIOUS, VOXEL_GWD = 0, 1, 2
K_COUNT, K_SCAN, K_BIN, K_TILE = 0, 1, 2, 3
KERNEL_NAMES = {K_COUNT: "k_hist", K_SCAN: "k_scan", K_BIN: "k_bin", K_TILE: "tile kernel"}
WF_OUT_OF_RANGE, WF_UNSORTED, WF_T_RANGE, WF_BAD_POLARITY = 0x100, 0x200, 0x400, 0x800

_c = ctypes
_vp, _i, _i64, _sz, _d = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t, _c.c_double
_EV = [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i]  # x, y, t, t_bytes, p, win_offsets, B, H, W
_TAIL = [_vp, _vp, _sz, _vp]                      # out, workspace, workspace_bytes, stream

SIGNATURES = {
    "evrep_version": (_i, []),
    "evrep_last_error": (_c.c_char_p, []),
    "evrep_profile_enable": (_i, [_i]),
    "evrep_profile_read": (_i, [_i, _vp, _vp]),
    "evrep_workspace_bytes": (_sz, [_i, _i, _i64, _i, _i, _i]),
    "evrep_window_flags": (_i, [_vp, _i, _vp, _vp]),
    "evrep_mixed_density_plan_info": (_i, [_i, _i, _vp, _vp, _vp, _i, _i, _i64, _vp]),
    "evrep_mixed_density_batched": (_i, _EV + [_vp, _vp, _vp, _i, _i] + _TAIL),
    "evrep_mixed_density_specialize": (_i, [_vp, _vp, _vp, _i, _i, _i64]),
    "evrep_mixed_density_specialize_async": (_i, [_vp, _vp, _vp, _i, _i, _i64]),
    "evrep_mixed_density_specialize_compile_only": (_i, [_vp, _vp, _vp, _i, _i, _i64, _vp]),
    "evrep_mixed_density_is_specialized": (_i, [_vp, _vp, _vp, _i, _i, _i64]),
    "evrep_ergo12_batched": (_i, _EV + [_i] + _TAIL),
    "evrep_event_stack_batched": (_i, _EV + [_i] + _TAIL),
    "evrep_time_surface_batched": (_i, _EV + [_vp, _i, _d] + _TAIL),
    "evrep_tore_batched": (_i, _EV + [_i] + _TAIL),
    "evrep_order_ops_fused_batched": (_i, _EV + [_d, _vp, _vp] + _TAIL),
    "evrep_voxel_batched": (_i, _EV + [_i, _i, _i, _vp] + _TAIL),
    "evrep_voxel_subpixel_batched": (_i, _EV + [_i, _i, _i, _vp] + _TAIL),
    "evrep_histogram_batched": (_i, _EV + _TAIL),
    "evrep_gwd_workspace_bytes": (_sz, [_vp, _vp, _i]),
    "evrep_gwd_kernel_l1": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _d, _vp, _vp, _sz, _vp]),
    "evrep_gw_kl_workspace_bytes": (_sz, [_i, _i]),
    "evrep_gw_kl": (_i, [_vp, _i, _i, _vp, _i, _i, _d, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "evrep_filter_batched": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _d, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "evrep_filter_background_workspace_bytes": (_sz, [_i, _i64, _i, _i, _i, _i]),
    "evrep_filter_background_batched": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _d, _i, _vp, _vp, _vp, _sz, _vp]),
    "evrep_est_workspace_bytes": (_sz, [_i]),
    "evrep_est_quantize_batched": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "evrep_est_backward_batched": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "evrep_assignment_auction": (_i, [_vp, _i, _d, _vp, _vp, _vp]),
    "evrep_unpack_workspace_bytes": (_sz, [_i, _i64]),
    "evrep_unpack_events": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "evrep_unpack_delta_workspace_bytes": (_sz, [_i]),
    "evrep_pack_delta_host_blocks": (_i64, [_vp, _i]),
    "evrep_pack_host_blocks": (_i64, [_vp, _i, _i]),
    "evrep_pack_events_host": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i]),
    "evrep_pack_events_delta_host": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _i, _i]),
    "evrep_unpack_events_delta": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "evrep_transport_plan_host": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "evrep_otmi_workspace_bytes": (_sz, [_i64, _i]),
    "evrep_otmi_prepare": (_i, [_vp, _i, _i64, _vp, _i, _i, _i, _i, _vp, _i64, _vp, _i64, _vp, _vp, _sz, _vp]),
    "evrep_gemm_workspace_bytes": (_sz, [_i, _i, _i]),
    "evrep_gemm_nt_3xtf32": (_i, [_vp, _vp, _vp, _i, _i, _i, _c.c_float, _vp, _vp, _vp, _sz, _vp]),
    "evrep_warp_affine_batched": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _c.c_float, _vp, _vp]),
    "evrep_image_pipeline_batched": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _c.c_float, _c.c_float, _c.c_float, _i, _vp, _vp]),
}


class EvrepError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libevrep error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.evrep_version() != 100:
        raise ImportError(f"libevrep version {lib.evrep_version()} does not match the Python layer (100)")
    return lib


lib = _load()


def check(rc):
    if rc != OK:
        raise EvrepError(rc, lib.evrep_last_error().decode("utf-8", "replace"))


def profile_enable(max_calls):
    check(lib.evrep_profile_enable(int(max_calls)))


def profile_read(kernel_id):
    """-> (total milliseconds, launches timed) for one kernel of the tile pipeline since profile_enable."""
    ms, n = ctypes.c_float(0), ctypes.c_int(0)
    check(lib.evrep_profile_read(int(kernel_id), ctypes.byref(ms), ctypes.byref(n)))
    return ms.value, n.value
