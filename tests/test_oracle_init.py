"""Known-answer tests pinning oracle/init_oracle.py: exact data -> exact answers."""
import numpy as np

from oracle import init_oracle as io


def _scene(n, seed, noise=0.0, outliers=0):
    rng = np.random.default_rng(seed)
    q1 = rng.standard_normal(4); q1 /= np.linalg.norm(q1)
    q2 = q1 + 0.15 * rng.standard_normal(4); q2 /= np.linalg.norm(q2)
    R1, R2 = io.quat_to_rot(q1), io.quat_to_rot(q2)
    c1, c2 = rng.standard_normal(3), rng.standard_normal(3)
    mid = 0.5 * (c1 + c2)
    X = mid + R1.T @ np.array([0, 0, 6.0]) + rng.uniform(-1.5, 1.5, (n, 3))
    def project(R, c):
        p = (X - c) @ R.T
        return p[:, :2] / p[:, 2:3]
    p1, p2 = project(R1, c1), project(R2, c2)
    p2 = p2 + noise * rng.standard_normal(p2.shape)
    if outliers:
        p2[:outliers] += rng.uniform(-0.3, 0.3, (outliers, 2))
    return p1, p2, q1, q2, R1, R2, c1, c2, X


def test_quaternion_convention():
    R = io.quat_to_rot([np.cos(0.2), 0, 0, np.sin(0.2)])           # rotation about z by 0.4 rad
    assert np.allclose(R, [[np.cos(0.4), -np.sin(0.4), 0], [np.sin(0.4), np.cos(0.4), 0], [0, 0, 1]])


def test_exact_correspondences_give_the_true_direction():
    p1, p2, q1, q2, R1, R2, c1, c2, _ = _scene(200, 1)
    t = io.optimize_relative_position_with_known_rotation(p1, p2, q1, q2)
    # t is the relative translation of camera 2 w.r.t. camera 1 (P2 = [R2 R1' | t]) up to scale:
    # x2 ~ R2 R1' x1 + t  with  t = R2 (c1 - c2) / |c1 - c2|
    truth = R2 @ (c1 - c2)
    truth /= np.linalg.norm(truth)
    assert abs(np.linalg.norm(t) - 1.0) < 1e-12
    assert np.abs(t - truth).max() < 1e-9


def test_irls_is_robust_to_outliers_and_chooses_the_sign_by_cheirality():
    p1, p2, q1, q2, R1, R2, c1, c2, _ = _scene(400, 2, noise=1e-4, outliers=60)
    t, its = io.optimize_relative_position_with_known_rotation(p1, p2, q1, q2, return_iterations=True)
    truth = R2 @ (c1 - c2)
    truth /= np.linalg.norm(truth)
    assert t @ truth > 0.9995
    assert 10 <= its <= 100


def test_multi_view_dlt_recovers_exact_points():
    rng = np.random.default_rng(3)
    X = rng.uniform(-1, 1, 3) + np.array([0, 0, 5.0])
    proj, xy = [], []
    for _ in range(7):
        q = np.r_[1.0, 0.1 * rng.standard_normal(3)]
        R = io.quat_to_rot(q)
        t = 0.5 * rng.standard_normal(3)
        p = R @ X + t
        proj.append(np.c_[R, t]); xy.append(p[:2] / p[2])
    assert np.abs(io.triangulate_multi_view_point(proj, xy) - X).max() < 1e-10
    assert np.abs(io.triangulate_multi_view_point(proj[:2], xy[:2]) - X).max() < 1e-10


def test_two_view_dlt_and_cheirality_count():
    p1, p2, q1, q2, R1, R2, c1, c2, X = _scene(50, 4)
    R = R2 @ R1.T
    t = R2 @ (c1 - c2)
    P1, P2 = np.c_[np.eye(3), np.zeros(3)], np.c_[R, t]
    Xc1 = (X - c1) @ R1.T
    for k in range(5):
        assert np.abs(io.triangulate_point(P1, P2, p1[k], p2[k]) - Xc1[k]).max() < 1e-9
    assert io.count_in_front(p1, p2, R, t) == 50
    assert io.count_in_front(p1, p2, R, -t) == 0
