"""ctypes binding of the C ABI declared in include/mpp_b200.h.  There is no CPU fallback: if the CUDA library
is missing or fails to load, importing a device object raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libmpp_b200.so")

NO_OBJECT = 0xFFFFFFFF
MAX_TERMS = 8
WINDOW_STATS = 35
IPC_HANDLE_BYTES = 64
PRECISION_FP32, PRECISION_FP64 = 0, 1
SETUP_LEGACY, SETUP_NO_CALIBRATION, SETUP_TOY = 0, 1, 2
COMB_RAW_SUM, COMB_HIERARCHICAL, COMB_LOGISTIC, COMB_MANUAL_HIERARCHICAL = 0, 1, 2, 3

ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_OUT_OF_BOUNDS, ERR_CELL_FULL, ERR_NEIGHBOURHOOD, ERR_NOT_FOUND, ERR_TIMEOUT = -1, -2, -3, -4, -5, -6, -7, -8


class ModelParams(C.Structure):
    _fields_ = [("setup", C.c_int32), ("combinator", C.c_int32), ("ratio_prior", C.c_int32), ("rewarding", C.c_int32),
                ("pos_threshold", C.c_double), ("remap_coef", C.c_double * 3), ("remap_intercept", C.c_double * 3),
                ("min_area", C.c_double), ("max_area", C.c_double), ("target_ratio", C.c_double),
                ("overlap_max_dist", C.c_double), ("align_max_dist", C.c_double), ("comb_w", C.c_double * MAX_TERMS),
                ("comb_bias", C.c_double), ("comb_threshold", C.c_double), ("toy_unit_value", C.c_double),
                ("toy_pair_value", C.c_double), ("toy_pair_dist", C.c_double), ("toy_pair_strict", C.c_int32),
                ("marks_are_energies", C.c_int32)]


class KernelParams(C.Structure):
    _fields_ = [("p_kernel", C.c_double * 8), ("intensity", C.c_double), ("gauss_translation_sigma", C.c_double),
                ("data_translation_max_delta", C.c_int32), ("reserved", C.c_int32), ("gauss_transform_sigma", C.c_double)]


# numpy mirrors of mpp_proposal / mpp_step_result (C layout, natural alignment)
PROPOSAL_DTYPE = np.dtype([("kernel", "<i4"), ("rem_x", "<i4"), ("rem_y", "<i4"), ("rem_uid", "<u4"), ("add_x", "<i4"),
                           ("add_y", "<i4"), ("add_uid", "<u4"), ("add_cls", "<u4"), ("add_size", "<f8"),
                           ("add_ratio", "<f8"), ("add_angle", "<f8"), ("delta0", "<f8"), ("delta1", "<f8"),
                           ("param_id", "<i4"), ("new_class", "<i4"), ("u", "<f8")], align=True)
STEP_RESULT_DTYPE = np.dtype([("delta_e", "<f8"), ("fwd", "<f8"), ("bwd", "<f8"), ("log_alpha", "<f8"),
                              ("temperature", "<f8"), ("accepted", "<i4"), ("n_after", "<i4")], align=True)

# mpp_window_trace: one record per proposal of the window sampler (debug instantiation)
WINDOW_TRACE_DTYPE = np.dtype([("flags", "<u4"), ("rem_uid", "<u4"), ("add_uid", "<u4"), ("add_cls", "<u4"), ("add_x", "<i4"),
                               ("add_y", "<i4"), ("add_size", "<f4"), ("add_ratio", "<f4"), ("add_angle", "<f4"),
                               ("delta_e", "<f4"), ("log_ratio", "<f4"), ("temperature", "<f4"), ("u_accept", "<f4"),
                               ("q", "<u4", (3,))], align=True)
TRACE_WRITTEN, TRACE_EVALUATED, TRACE_ACCEPT, TRACE_IDENTITY, TRACE_HAS_ADD, TRACE_HAS_REM, TRACE_LEFT_WINDOW, TRACE_CELL_FULL = \
    1, 2, 4, 8, 16, 32, 64, 128

SPLIT_MERGE_DTYPE = np.dtype([("kind", "<i4"), ("n_neighbors", "<i4"), ("n_add", "<i4"), ("reserved", "<i4"), ("rem_uid", "<u4", (2,)),
                              ("rem_x", "<i4", (2,)), ("rem_y", "<i4", (2,)), ("add_x", "<i4", (2,)), ("add_y", "<i4", (2,)),
                              ("add_size", "<f8", (2,)), ("add_ratio", "<f8", (2,)), ("add_angle", "<f8", (2,)), ("pos_delta", "<f8", (2,)),
                              ("shape_delta", "<f8", (3,)), ("u", "<f8")], align=True)

# every symbol include/mpp_b200.h declares
SYMBOLS = ["mpp_abi_version", "mpp_last_error", "mpp_abi_struct_size", "mpp_ctx_create", "mpp_ctx_destroy", "mpp_set_maps",
           "mpp_set_model", "mpp_set_kernels", "mpp_add_objects", "mpp_remove_objects", "mpp_clear_objects",
           "mpp_num_objects", "mpp_read_objects", "mpp_energy_vectors", "mpp_delta_batch", "mpp_replay",
           "mpp_run_sweeps", "mpp_sample_births", "mpp_naive_init", "mpp_pack_rows", "mpp_unpack_rows",
           "mpp_query_neighbors", "mpp_copy_state", "mpp_pair_values", "mpp_run_chain", "mpp_sample_proposals",
           "mpp_proposal_probs", "mpp_combine", "mpp_run_windows", "mpp_ctx_reset", "mpp_run_window_rows", "mpp_window_grid",
           "mpp_set_window_trace", "mpp_window_stats", "mpp_run_windows_batch", "mpp_split_export", "mpp_split_attach",
           "mpp_split_attach_local", "mpp_split_detach", "mpp_set_maps_band", "mpp_sample_points_2d", "mpp_sample_split_merge", "mpp_split_merge_probs"]

_lib = None


class MPPError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[mpp_b200 {code}] {message}")
        self.code = code


def load():
    """Loads libmpp_b200.so (built in-tree by build.py).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    if os.environ.get("MPP_B200_DEBUG") == "1" and os.environ.get("MPP_B200_DEBUG_LIB"):
        # development only, and only when explicitly switched on: an instrumented build of the same library (tools/visit_timers.py)
        path = os.path.abspath(os.environ["MPP_B200_DEBUG_LIB"])
        if not path.startswith(os.path.dirname(PKG_DIR) + os.sep):
            raise RuntimeError("MPP_B200_DEBUG_LIB must point inside the repository")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: run `python -m mpp_cnn_rs_object_detection_b200.build` "
                           f"(the MPP sampler has no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, f64, u64 = C.c_void_p, C.c_int, C.c_double, C.c_uint64
    for name in SYMBOLS:
        getattr(lib, name).restype = C.c_int
    lib.mpp_last_error.restype = C.c_char_p
    lib.mpp_abi_struct_size.argtypes = [i32]
    lib.mpp_ctx_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, vp]
    lib.mpp_ctx_destroy.argtypes = [vp]
    lib.mpp_ctx_reset.argtypes = [vp, vp]
    lib.mpp_set_maps.argtypes = [vp, vp, vp, f64]
    lib.mpp_sample_split_merge.argtypes = [vp, i32, f64, C.POINTER(f64), u64, u64, vp]
    lib.mpp_split_merge_probs.argtypes = [vp, vp, f64, f64, f64, C.POINTER(f64), vp]
    lib.mpp_sample_points_2d.argtypes = [vp, i32, i32, i32, u64, vp, vp, i32, vp]
    lib.mpp_set_maps_band.argtypes = [vp, vp, vp, i32, i32, f64]
    lib.mpp_set_model.argtypes = [vp, C.POINTER(ModelParams)]
    lib.mpp_set_kernels.argtypes = [vp, C.POINTER(KernelParams)]
    lib.mpp_add_objects.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    lib.mpp_remove_objects.argtypes = [vp, vp, i32]
    lib.mpp_clear_objects.argtypes = [vp]
    lib.mpp_num_objects.argtypes = [vp, C.POINTER(i32)]
    lib.mpp_read_objects.argtypes = [vp, i32, vp, vp, vp, vp, C.POINTER(i32)]
    lib.mpp_energy_vectors.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.mpp_delta_batch.argtypes = [vp, vp, i32, vp]
    lib.mpp_replay.argtypes = [vp, vp, i32, f64, f64, f64, vp]
    lib.mpp_run_sweeps.argtypes = [vp, i32, i32, i32, f64, f64, f64, u64, u64, C.POINTER(C.c_ulonglong)]
    lib.mpp_sample_births.argtypes = [vp, i32, u64, vp]
    lib.mpp_naive_init.argtypes = [vp, f64, f64, C.POINTER(i32)]
    lib.mpp_pack_rows.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32)]
    lib.mpp_unpack_rows.argtypes = [vp, i32, i32, vp, i32]
    lib.mpp_query_neighbors.argtypes = [vp, i32, i32, f64, i32, C.c_uint32, i32, vp, C.POINTER(i32)]
    lib.mpp_copy_state.argtypes = [vp, vp]
    lib.mpp_pair_values.argtypes = [vp, vp, vp, i32, vp]
    lib.mpp_run_chain.argtypes = [vp, i32, f64, f64, f64, u64, u64, vp, C.POINTER(C.c_ulonglong)]
    lib.mpp_sample_proposals.argtypes = [vp, vp, i32, u64, u64, vp]
    lib.mpp_proposal_probs.argtypes = [vp, vp, i32, vp]
    lib.mpp_run_windows.argtypes = [vp, i32, i32, i32, i32, f64, f64, f64, u64, u64, C.POINTER(C.c_ulonglong), vp]
    lib.mpp_run_window_rows.argtypes = [vp, i32, i32, f64, u64, u64, i32, i32, i32]
    lib.mpp_window_grid.argtypes = [vp, u64, u64, C.POINTER(i32), C.POINTER(i32)]
    lib.mpp_set_window_trace.argtypes = [vp, vp, u64, u64]
    lib.mpp_window_stats.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    lib.mpp_run_windows_batch.argtypes = [C.POINTER(vp), C.POINTER(u64), i32, u64, i32, i32, i32, f64, f64, f64, u64, i32,
                                          C.POINTER(C.c_ulonglong), vp]
    lib.mpp_split_export.argtypes = [vp, C.c_char_p]
    lib.mpp_split_attach.argtypes = [vp, i32, i32, C.c_char_p, C.c_char_p]
    lib.mpp_split_attach_local.argtypes = [vp, i32, i32, vp, vp]
    lib.mpp_split_detach.argtypes = [vp]
    lib.mpp_combine.argtypes = [C.POINTER(ModelParams), vp, i32, vp, vp, i32, vp]
    if lib.mpp_abi_version() != 1:
        raise RuntimeError("libmpp_b200.so ABI version mismatch")
    sizes = [C.sizeof(ModelParams), C.sizeof(KernelParams), PROPOSAL_DTYPE.itemsize, STEP_RESULT_DTYPE.itemsize,
             WINDOW_TRACE_DTYPE.itemsize, SPLIT_MERGE_DTYPE.itemsize]
    for which, sz in enumerate(sizes):
        if lib.mpp_abi_struct_size(which) != sz:
            raise RuntimeError(f"ABI struct {which} size mismatch: C {lib.mpp_abi_struct_size(which)} vs python {sz}")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise MPPError(rc, load().mpp_last_error().decode())
