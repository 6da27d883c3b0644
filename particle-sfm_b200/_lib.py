"""ctypes binding of libpsfm_b200.so (the C ABI of include/psfm_b200.h).

The library is the product: there is no Python/NumPy fallback.  Loading fails loudly if
the shared object has not been built (python -m particlesfm_b200.build)."""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# PSFM_LIB: alternative build of the same library (kernel-tuning experiments only)
LIB_PATH = os.environ.get("PSFM_LIB") or os.path.join(_HERE, "libpsfm_b200.so")
_LIB = None

EXPORTS = [
    "psfm_last_error", "psfm_abi_version", "psfm_device_count", "psfm_set_device", "psfm_launch_count",
    "psfm_traj_default_options", "psfm_traj_optimize", "psfm_traj_optimize_device",
    "psfm_ba_default_options", "psfm_ba_global_options", "psfm_ba_solve", "psfm_ba_create",
    "psfm_ba_set_state", "psfm_ba_run", "psfm_ba_get_state", "psfm_ba_destroy", "psfm_ba_evaluate",
    "psfm_ba_linear_step", "psfm_ba_band_solve", "psfm_measure_dfma", "psfm_ba_default_refine_options",
    "psfm_ba_filter_negative_depth", "psfm_ba_filter_points", "psfm_ba_normalize", "psfm_ba_num_observations",
    "psfm_ba_get_observation_mask", "psfm_ba_get_point_errors", "psfm_ba_iterative_refinement",
    "psfm_grid_sample", "psfm_flow_check", "psfm_tracker_step", "psfm_tracker_buffer_inputs", "psfm_known_rotation_translations", "psfm_triangulate_tracks", "psfm_dist_get_unique_id", "psfm_dist_init", "psfm_dist_world_size",
    "psfm_dist_rank", "psfm_dist_finalize",
]


class PsfmError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise PsfmError("libpsfm_b200.so is not built — run `python -m particlesfm_b200.build` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
    L.psfm_last_error.restype = C.c_char_p
    L.psfm_launch_count.restype = C.c_int64
    L.psfm_traj_default_options.argtypes = [C.POINTER(_abi.TrajOptions)]
    L.psfm_traj_default_options.restype = None
    L.psfm_traj_optimize.argtypes = [dp, dp, dp, dp, fp, C.c_int32, C.c_int32, C.c_int32,
                                     C.POINTER(_abi.TrajOptions), dp, C.POINTER(_abi.TrajSummary)]
    L.psfm_traj_optimize_device.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int32, C.c_int32,
                                                               C.POINTER(_abi.TrajOptions), C.c_void_p,
                                                               C.POINTER(_abi.TrajSummary), C.c_void_p]
    L.psfm_ba_default_options.argtypes = [C.POINTER(_abi.BAOptions)]
    L.psfm_ba_default_options.restype = None
    L.psfm_ba_global_options.argtypes = [C.POINTER(_abi.BAOptions)]
    L.psfm_ba_global_options.restype = None
    L.psfm_ba_solve.argtypes = [C.POINTER(_abi.BAProblemStruct), C.POINTER(_abi.BAOptions),
                                C.POINTER(_abi.BASummary)]
    L.psfm_ba_create.argtypes = [C.POINTER(_abi.BAProblemStruct), C.POINTER(C.c_void_p)]
    L.psfm_ba_set_state.argtypes = [C.c_void_p, dp, dp, dp, dp]
    L.psfm_ba_get_state.argtypes = [C.c_void_p, dp, dp, dp, dp]
    L.psfm_ba_run.argtypes = [C.c_void_p, C.POINTER(_abi.BAOptions), C.POINTER(_abi.BASummary)]
    L.psfm_ba_destroy.argtypes = [C.c_void_p]
    L.psfm_ba_destroy.restype = None
    L.psfm_ba_evaluate.argtypes = [C.c_void_p, C.POINTER(_abi.BAOptions), dp, dp, dp, dp]
    L.psfm_ba_linear_step.argtypes = [C.c_void_p, C.POINTER(_abi.BAOptions), C.c_double, dp, dp, ip]
    L.psfm_measure_dfma.argtypes = [dp, dp]
    u8p = C.POINTER(C.c_uint8)
    L.psfm_grid_sample.argtypes = [fp, C.c_int32, C.c_int32, C.c_int32, dp, C.c_int32, fp]
    L.psfm_flow_check.argtypes = [fp, fp, C.c_int32, C.c_int32, C.c_float, fp, u8p]
    L.psfm_tracker_step.argtypes = [fp, u8p, C.c_int32, C.c_int32, dp, C.c_int32, C.c_int32, dp, u8p, u8p]
    L.psfm_tracker_buffer_inputs.argtypes = [fp, fp, u8p, C.c_int32, C.c_int32, dp, C.c_int32, C.c_double, dp, dp, dp]
    L.psfm_known_rotation_translations.argtypes = [dp, dp, ip, dp, dp, C.c_int32, dp, ip]
    L.psfm_triangulate_tracks.argtypes = [dp, dp, ip, C.c_int32, dp]
    i64p = C.POINTER(C.c_int64)
    L.psfm_ba_default_refine_options.argtypes = [C.POINTER(_abi.BARefineOptions)]
    L.psfm_ba_default_refine_options.restype = None
    L.psfm_ba_filter_negative_depth.argtypes = [C.c_void_p, i64p]
    L.psfm_ba_filter_points.argtypes = [C.c_void_p, C.c_double, C.c_double, i64p]
    L.psfm_ba_normalize.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, dp, dp]
    L.psfm_ba_num_observations.argtypes = [C.c_void_p, i64p]
    L.psfm_ba_get_observation_mask.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
    L.psfm_ba_get_point_errors.argtypes = [C.c_void_p, dp]
    L.psfm_ba_iterative_refinement.argtypes = [C.c_void_p, C.POINTER(_abi.BAOptions), C.POINTER(_abi.BARefineOptions),
                                               C.POINTER(_abi.BARefineReport)]
    L.psfm_ba_band_solve.argtypes = [dp, dp, C.c_int32, C.c_int32, dp]
    L.psfm_dist_get_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.psfm_dist_init.argtypes = [C.POINTER(C.c_uint8), C.c_int32, C.c_int32]
    L.psfm_dist_finalize.restype = None
    _LIB = L
    return L


def check(rc, what):
    if rc < 0:
        msg = lib().psfm_last_error().decode("utf-8", "replace")
        raise PsfmError(f"{what} failed with status {rc}: {msg}")
    return rc


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
