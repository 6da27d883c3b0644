"""HP1 through the C ABI: `optimize_location` with the reference's Python signature
(point_trajectory/optimize/src/bindings.cc:31, trajectory_optimize.h:35-42)."""
import ctypes as C

import numpy as np

from . import _abi, _lib


def default_options():
    o = _abi.TrajOptions()
    _lib.lib().psfm_traj_default_options(C.byref(o))
    return o


def optimize_location(uv12, uv_ref1, uv_ref2, ref2_scale, flow12_map, total_num=None, width=None,
                      height=None, options=None, return_summary=False):
    """Drop-in for particlesfm.optimize_location: returns the optimised [N, 4] array.
    Inputs are copied, never mutated (the reference takes its Eigen arguments by value)."""
    uv12 = np.ascontiguousarray(uv12, np.float64).reshape(-1, 4)
    n = uv12.shape[0] if total_num is None else int(total_num)
    ref1 = np.ascontiguousarray(uv_ref1, np.float64).reshape(-1, 2)
    ref2 = np.ascontiguousarray(uv_ref2, np.float64).reshape(-1, 2)
    scale = np.ascontiguousarray(ref2_scale, np.float64).reshape(-1)
    flow = np.ascontiguousarray(flow12_map, np.float32)   # f32 -> f64 widening on the device is exact
    h, w = flow.shape[:2]
    if width is not None and (int(width) != w or int(height) != h):
        raise ValueError("flow12_map shape does not match width/height")
    if min(uv12.shape[0], ref1.shape[0], ref2.shape[0], scale.shape[0]) < n:
        raise ValueError("total_num exceeds the number of rows")
    out = np.empty((n, 4), np.float64)
    s = _abi.TrajSummary()
    rc = _lib.lib().psfm_traj_optimize(_lib.dptr(uv12), _lib.dptr(ref1), _lib.dptr(ref2), _lib.dptr(scale),
                                       flow.ctypes.data_as(C.POINTER(C.c_float)), n, w, h,
                                       C.byref(options) if options is not None else None, _lib.dptr(out),
                                       C.byref(s))
    _lib.check(rc, "psfm_traj_optimize")
    return (out, s) if return_summary else out
