"""ctypes loader of the CPU oracle (TEST INFRASTRUCTURE — see oracle/psfm_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, _ROOT) if _ROOT not in sys.path else None
from particlesfm_b200 import _abi  # noqa: E402  (struct definitions only)

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libpsfm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ba_oracle.c", "ba_schur_blocks.h", "traj_oracle.c", "psfm_oracle.h",
                                             "ceres_semantics.h", "Makefile")]
    srcs.append(os.path.join(_ROOT, "include", "psfm_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "MAKEFLAGS": ""})
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.psfm_oracle_traj_optimize.argtypes = [dp, dp, dp, dp, fp, C.c_int32, C.c_int32, C.c_int32,
                                                C.POINTER(_abi.TrajOptions), dp,
                                                C.POINTER(_abi.TrajSummary), C.c_int32, C.c_int32]
        L.psfm_oracle_traj_evaluate.argtypes = [dp, dp, dp, dp, fp, C.c_int32, C.c_int32, C.c_int32, dp, dp]
        L.psfm_oracle_bilinear.argtypes = [fp, C.c_int32, C.c_int32, C.c_double, C.c_double, dp, dp, dp]
        L.psfm_oracle_bilinear.restype = None
        L.psfm_oracle_ba_solve.argtypes = [C.POINTER(_abi.BAProblemStruct), C.POINTER(_abi.BAOptions),
                                           C.POINTER(_abi.BASummary), C.c_int32]
        L.psfm_oracle_ba_evaluate.argtypes = [C.POINTER(_abi.BAProblemStruct), C.POINTER(_abi.BAOptions),
                                              dp, dp, dp, dp, C.c_int32]
        L.psfm_oracle_ba_jacobians.argtypes = [C.POINTER(_abi.BAProblemStruct), C.POINTER(_abi.BAOptions),
                                               dp, dp, dp]
        L.psfm_oracle_ba_linear_step.argtypes = [C.POINTER(_abi.BAProblemStruct), C.POINTER(_abi.BAOptions),
                                                 C.c_double, C.c_int32, dp, dp, ip, C.c_int32]
        L.psfm_oracle_ba_default_options.argtypes = [C.POINTER(_abi.BAOptions)]
        L.psfm_oracle_ba_default_options.restype = None
        L.psfm_oracle_ba_global_options.argtypes = [C.POINTER(_abi.BAOptions)]
        L.psfm_oracle_ba_global_options.restype = None
        L.psfm_oracle_traj_default_options.argtypes = [C.POINTER(_abi.TrajOptions)]
        L.psfm_oracle_traj_default_options.restype = None
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def num_threads():
    return int(lib().psfm_oracle_num_threads())


def ba_default_options():
    o = _abi.BAOptions()
    lib().psfm_oracle_ba_default_options(C.byref(o))
    return o


def ba_global_options(refine_rotation=False, refine_focal_length=False):
    """GlobalMapperOptions::GlobalBundleAdjustment (controllers/global_mapper.cc:41-71);
    pass B flips refine_rotation / refine_focal_length (:219-224)."""
    o = _abi.BAOptions()
    lib().psfm_oracle_ba_global_options(C.byref(o))
    o.refine_rotation = int(refine_rotation)
    o.refine_focal_length = int(refine_focal_length)
    o.minimizer_progress_to_stdout = 0
    o.print_summary = 0
    return o


def traj_default_options():
    o = _abi.TrajOptions()
    lib().psfm_oracle_traj_default_options(C.byref(o))
    return o


def traj_optimize(uv12, ref1, ref2, scale, flow12, opts=None, num_threads=0, reduction_mode=0):
    """Oracle twin of particlesfm.optimize_location: returns (out [N,4], summary)."""
    uv12 = np.ascontiguousarray(uv12, np.float64).reshape(-1, 4)
    n = uv12.shape[0]
    ref1 = np.ascontiguousarray(ref1, np.float64).reshape(n, 2)
    ref2 = np.ascontiguousarray(ref2, np.float64).reshape(n, 2)
    scale = np.ascontiguousarray(scale, np.float64).reshape(n)
    flow12 = np.ascontiguousarray(flow12, np.float32)
    h, w = flow12.shape[:2]
    out = np.empty_like(uv12)
    s = _abi.TrajSummary()
    rc = lib().psfm_oracle_traj_optimize(_d(uv12), _d(ref1), _d(ref2), _d(scale),
                                         flow12.ctypes.data_as(C.POINTER(C.c_float)), n, w, h,
                                         C.byref(opts) if opts is not None else None, _d(out),
                                         C.byref(s), num_threads, reduction_mode)
    assert rc == 0, rc
    return out, s


def traj_evaluate(uv12, ref1, ref2, scale, flow12):
    uv12 = np.ascontiguousarray(uv12, np.float64).reshape(-1, 4)
    n = uv12.shape[0]
    ref1 = np.ascontiguousarray(ref1, np.float64).reshape(n, 2)
    ref2 = np.ascontiguousarray(ref2, np.float64).reshape(n, 2)
    scale = np.ascontiguousarray(scale, np.float64).reshape(n)
    flow12 = np.ascontiguousarray(flow12, np.float32)
    h, w = flow12.shape[:2]
    r = np.empty((n, 6))
    J = np.empty((n, 6, 4))
    lib().psfm_oracle_traj_evaluate(_d(uv12), _d(ref1), _d(ref2), _d(scale),
                                    flow12.ctypes.data_as(C.POINTER(C.c_float)), n, w, h, _d(r), _d(J))
    return r, J


def bilinear(flow, r, c):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w = flow.shape[:2]
    f, dr, dc = np.empty(2), np.empty(2), np.empty(2)
    lib().psfm_oracle_bilinear(flow.ctypes.data_as(C.POINTER(C.c_float)), w, h, float(r), float(c),
                               _d(f), _d(dr), _d(dc))
    return f, dr, dc


def ba_solve(problem, opts, num_threads=0):
    """In-place solve of a particlesfm_b200.BAProblem; returns the BASummary."""
    s = _abi.BASummary()
    st = problem.struct()
    rc = lib().psfm_oracle_ba_solve(C.byref(st), C.byref(opts), C.byref(s), num_threads)
    assert rc in (0, 1), rc
    return s


def ba_evaluate(problem, opts, num_threads=0):
    F, P, M, Cn = problem.num_images, problem.num_points, problem.num_observations, problem.num_cameras
    cost = C.c_double()
    r = np.empty(2 * M)
    gc = np.empty(6 * F + 3 * Cn)
    gp = np.empty(3 * P)
    st = problem.struct()
    rc = lib().psfm_oracle_ba_evaluate(C.byref(st), C.byref(opts), C.byref(cost), _d(r), _d(gc), _d(gp),
                                       num_threads)
    assert rc == 0, rc
    return cost.value, r.reshape(M, 2), gc, gp.reshape(P, 3)


def ba_jacobians(problem, opts):
    M = problem.num_observations
    jc, jp, jk = np.empty((M, 2, 6)), np.empty((M, 2, 3)), np.empty((M, 2, 3))
    st = problem.struct()
    rc = lib().psfm_oracle_ba_jacobians(C.byref(st), C.byref(opts), _d(jc), _d(jp), _d(jk))
    assert rc == 0, rc
    return jc, jp, jk


def ba_linear_step(problem, opts, radius, solver, num_threads=0):
    F, P, Cn = problem.num_images, problem.num_points, problem.num_cameras
    sc = np.empty(6 * F + 3 * Cn)
    sp = np.empty(3 * P)
    it = C.c_int32()
    st = problem.struct()
    rc = lib().psfm_oracle_ba_linear_step(C.byref(st), C.byref(opts), radius, solver, _d(sc), _d(sp),
                                          C.byref(it), num_threads)
    assert rc == 0, rc
    return sc, sp.reshape(P, 3), it.value
