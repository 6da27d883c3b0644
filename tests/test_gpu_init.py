"""SURVEY.md 8(f) row f-4 on the device (csrc/init_geometry.cu) against oracle/init_oracle.py, through the C ABI."""
import numpy as np
import pytest

from oracle import init_oracle as io
from particlesfm_b200 import init_geometry as ig
from test_oracle_init import _scene

pytestmark = pytest.mark.gpu


def test_known_rotation_translation_matches_the_oracle(gpu):
    pairs, ref, ref_its = [], [], []
    for seed, (n, noise, outl) in enumerate([(200, 0.0, 0), (400, 1e-4, 60), (33, 1e-3, 3), (1500, 5e-4, 300), (8, 0.0, 0), (129, 2e-3, 0)]):
        p1, p2, q1, q2, *_ = _scene(n, 10 + seed, noise=noise, outliers=outl)
        pairs.append((p1, p2, q1, q2))
        t, its = io.optimize_relative_position_with_known_rotation(p1, p2, q1, q2, return_iterations=True)
        ref.append(t); ref_its.append(its)
    t_gpu, its = ig.batch_optimize_relative_position_with_known_rotation(pairs, return_iterations=True)
    for k in range(len(pairs)):
        assert abs(np.linalg.norm(t_gpu[k]) - 1.0) < 1e-12
        # tolerance: exact data -> 1e-9.  With noise / outliers the IRLS weights 1 / max(|t' c|, 1e-7) amplify rounding:
        # the ORACLE ITSELF moves by up to 7e-6 when the correspondences are merely summed in another order
        # (measured on the 1500-point case, same iteration count), and the reference stops on a cost change of 1e-5.
        tol = 1e-9 if k in (0, 4) else 5e-5
        assert np.abs(t_gpu[k] - ref[k]).max() < tol, (k, t_gpu[k], ref[k])
        assert abs(int(its[k]) - ref_its[k]) <= 1
    # one pair through the scalar entry point, and the sign rule on a mirrored problem
    p1, p2, q1, q2, R1, R2, c1, c2, _ = _scene(100, 3)
    t = ig.optimize_relative_position_with_known_rotation(p1, p2, q1, q2)
    truth = R2 @ (c1 - c2)
    truth /= np.linalg.norm(truth)
    assert np.abs(t - truth).max() < 1e-9


def test_known_rotation_edge_cases(gpu):
    assert ig.batch_optimize_relative_position_with_known_rotation([]).shape == (0, 3)
    p1, p2, q1, q2, *_ = _scene(50, 5)
    t = ig.batch_optimize_relative_position_with_known_rotation([(p1[:0], p2[:0], q1, q2), (p1, p2, q1, q2)])
    assert np.array_equal(t[0], np.zeros(3))                      # no correspondences: nothing to estimate
    assert np.abs(t[1] - io.optimize_relative_position_with_known_rotation(p1, p2, q1, q2)).max() < 1e-9
    with pytest.raises(ValueError):
        ig.batch_optimize_relative_position_with_known_rotation([(p1, p2[:-1], q1, q2)])


def test_multi_view_triangulation_matches_the_oracle(gpu):
    rng = np.random.default_rng(7)
    tracks, ref = [], []
    for k in range(300):
        v = int(rng.integers(2, 40))
        X = rng.uniform(-1, 1, 3) + np.array([0, 0, 5.0])
        proj, xy = [], []
        for _ in range(v):
            R = io.quat_to_rot(np.r_[1.0, 0.1 * rng.standard_normal(3)])
            t = 0.5 * rng.standard_normal(3)
            p = R @ X + t
            proj.append(np.c_[R, t]); xy.append(p[:2] / p[2] + (1e-3 * rng.standard_normal(2) if k % 2 else 0.0))
        tracks.append((np.array(proj), np.array(xy)))
        ref.append(io.triangulate_multi_view_point(proj, xy))
    X_gpu = ig.triangulate_multi_view_points(tracks)
    assert np.abs(X_gpu - np.array(ref)).max() < 1e-9
    assert ig.triangulate_multi_view_points([]).shape == (0, 3)
