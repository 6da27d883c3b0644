"""TEST INFRASTRUCTURE (never imported by the product): numpy restatement of the two RANSAC-free initialisation
steps of SURVEY.md 8(f) row f-4.

* optimize_relative_position_with_known_rotation — sfm/gmapper/src/global/known_rotation_util.cc:55-196
  (CreateConstraintMatrix :55-79, the IRLS loop :116-176, MajorityOfPointsInFrontOfCameras :85-101 /
  sign choice :181-189).  CheckCheirality / TriangulatePoint / CalculateDepth are COLMAP bd84ad6
  (misc/doc/colmap.md:31, not vendored): base/pose.cc CheckCheirality (two-view DLT of every correspondence
  with P1 = [I|0], P2 = [R|t]; a point counts when both depths lie in (eps, 1000 |R' t|)),
  base/triangulation.cc TriangulatePoint (right singular vector of the 4 x 4 DLT matrix), base/projection.cc
  CalculateDepth (third row of P times X, times the norm of P's third column).
* triangulate_multi_view_point — COLMAP base/triangulation.cc TriangulateMultiViewPoint, the estimator the
  reference's IncrementalTriangulator::Create reaches through EstimateTriangulation
  (sfm/incremental_triangulator.cc:463-548): A = sum (P - x x' P)' (P - x x' P) over the views with x the
  normalised homogeneous ray, the point is the eigenvector of the smallest eigenvalue.
Parity unpinned (no reference vectors exist, COLMAP cannot be built here): pinned by the known-answer tests of
tests/test_oracle_init.py only.
"""
import numpy as np


def quat_to_rot(q):
    """COLMAP QuaternionToRotationMatrix (w, x, y, z), normalised."""
    w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def constraint_matrix(p1, p2, q1, q2):
    """known_rotation_util.cc:55-79: column i = R2 ((R1' f1) x (R2' f2))."""
    R1, R2 = quat_to_rot(q1), quat_to_rot(q2)
    f1 = np.c_[p1, np.ones(len(p1))] @ R1          # rows: (R1' f1)'
    f2 = np.c_[p2, np.ones(len(p2))] @ R2
    return (np.cross(f1, f2) @ R2.T).T               # 3 x N


def triangulate_point(P1, P2, x1, x2):
    A = np.stack([x1[0] * P1[2] - P1[0], x1[1] * P1[2] - P1[1], x2[0] * P2[2] - P2[0], x2[1] * P2[2] - P2[1]])
    v = np.linalg.svd(A)[2][-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return v[:3] / v[3]


def count_in_front(p1, p2, R, t):
    """COLMAP CheckCheirality: number of correspondences triangulated in front of both cameras."""
    P1 = np.c_[np.eye(3), np.zeros(3)]
    P2 = np.c_[R, t]
    eps = np.finfo(float).eps
    max_depth = 1000.0 * np.linalg.norm(R.T @ t)
    n = 0
    for a, b in zip(p1, p2):
        X = triangulate_point(P1, P2, a, b)
        d1 = X[2]
        if d1 > eps and d1 < max_depth:
            d2 = (R[2] @ X + t[2]) * np.linalg.norm(R[:, 2])
            if d2 > eps and d2 < max_depth:
                n += 1
    return n


def optimize_relative_position_with_known_rotation(p1, p2, q1, q2, return_iterations=False):
    """known_rotation_util.cc:107-196.  p1, p2: [N][2] normalised image points."""
    p1, p2 = np.asarray(p1, float), np.asarray(p2, float)
    C = constraint_matrix(p1, p2, q1, q2)
    w = np.ones(C.shape[1])
    cost, inner, pos, its = 0.0, 0, np.zeros(3), 0
    for _ in range(100):
        if inner >= 10:
            break
        its += 1
        w = np.where(w < 1e-7, 1e-7, w)
        lhs = (C / w) @ C.T
        new = np.linalg.svd(lhs)[0][:, -1]
        w = np.abs(new @ C)
        new_cost = w.sum()
        delta = max(abs(cost - new_cost), 1.0 - new @ new)
        inner = inner + 1 if delta <= 1e-5 else 0
        cost, pos = new_cost, new
    R1, R2 = quat_to_rot(q1), quat_to_rot(q2)
    if not count_in_front(p1, p2, R2 @ R1.T, pos) > len(p1) // 2:
        pos = -pos
    return (pos, its) if return_iterations else pos


def triangulate_multi_view_point(proj, xy):
    """proj: [V][3][4], xy: [V][2] normalised image points."""
    A = np.zeros((4, 4))
    for P, x in zip(np.asarray(proj, float), np.asarray(xy, float)):
        r = np.array([x[0], x[1], 1.0])
        r /= np.linalg.norm(r)
        term = P - np.outer(r, r) @ P
        A += term.T @ term
    v = np.linalg.eigh(A)[1][:, 0]
    return v[:3] / v[3]
