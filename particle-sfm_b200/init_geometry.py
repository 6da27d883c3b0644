"""Host mirror of the two batched initialisation ops of SURVEY.md 8(f) row f-4 (csrc/init_geometry.cu).

Names and argument meaning follow the reference:
* `optimize_relative_position_with_known_rotation(points1, points2, rotation1, rotation2)` —
  OptimizeRelativePositionWithKnownRotation, sfm/gmapper/src/global/known_rotation_util.cc:107-196
  (normalised image points, quaternions (w, x, y, z); returns the unit relative position);
* `batch_optimize_relative_position_with_known_rotation(pairs)` — BatchOptimize..., :198-229, every pair in
  one launch instead of one ThreadPool task per pair;
* `triangulate_multi_view_points(tracks)` — the multi-view DLT behind IncrementalTriangulator::Create
  (sfm/incremental_triangulator.cc:463-548), every track in one launch.
There is no CPU fallback: the library raises without a CUDA device."""
import ctypes as C

import numpy as np

from . import _lib


def _ptr(sizes):
    p = np.zeros(len(sizes) + 1, np.int32)
    np.cumsum(sizes, out=p[1:])
    return p


def batch_optimize_relative_position_with_known_rotation(pairs, return_iterations=False):
    """pairs: sequence of (points1 [n][2], points2 [n][2], rotation1 [4], rotation2 [4]); returns [len(pairs)][3]."""
    pairs = list(pairs)
    n = len(pairs)
    tvec = np.zeros((n, 3))
    its = np.zeros(n, np.int32)
    if n == 0:
        return (tvec, its) if return_iterations else tvec
    for a, b, _, _ in pairs:
        if np.shape(a) != np.shape(b):
            raise ValueError("points1 and points2 must have the same shape")       # CHECK_EQ, known_rotation_util.cc:115
    ptr = _ptr([len(a) for a, _, _, _ in pairs])
    p1 = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float64).reshape(-1, 2) for a, _, _, _ in pairs]))
    p2 = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float64).reshape(-1, 2) for _, b, _, _ in pairs]))
    q1 = np.ascontiguousarray(np.array([np.asarray(q, np.float64) for _, _, q, _ in pairs]).reshape(n, 4))
    q2 = np.ascontiguousarray(np.array([np.asarray(q, np.float64) for _, _, _, q in pairs]).reshape(n, 4))
    ip = C.POINTER(C.c_int32)
    _lib.check(_lib.lib().psfm_known_rotation_translations(_lib.dptr(p1), _lib.dptr(p2), ptr.ctypes.data_as(ip), _lib.dptr(q1),
                                                           _lib.dptr(q2), n, _lib.dptr(tvec), its.ctypes.data_as(ip)),
               "psfm_known_rotation_translations")
    return (tvec, its) if return_iterations else tvec


def optimize_relative_position_with_known_rotation(points1, points2, rotation1, rotation2):
    return batch_optimize_relative_position_with_known_rotation([(points1, points2, rotation1, rotation2)])[0]


def triangulate_multi_view_points(tracks):
    """tracks: sequence of (proj_matrices [v][3][4], points [v][2] normalised); returns [len(tracks)][3]."""
    tracks = list(tracks)
    n = len(tracks)
    xyz = np.zeros((n, 3))
    if n == 0:
        return xyz
    ptr = _ptr([len(p) for p, _ in tracks])
    P = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float64).reshape(-1, 12) for p, _ in tracks]))
    x = np.ascontiguousarray(np.concatenate([np.asarray(q, np.float64).reshape(-1, 2) for _, q in tracks]))
    if len(P) != len(x):
        raise ValueError("one image point per projection matrix")
    _lib.check(_lib.lib().psfm_triangulate_tracks(_lib.dptr(P), _lib.dptr(x), ptr.ctypes.data_as(C.POINTER(C.c_int32)), n, _lib.dptr(xyz)),
               "psfm_triangulate_tracks")
    return xyz
