"""Known-answer tests that pin the HP1 oracle (the reference ships none — SURVEY.md §4):
bilinear interpolator, residual/Jacobian definition, closed-form minimiser, scipy optimum,
thread- and reduction-order invariance."""
import numpy as np
import pytest
import scipy.optimize

import oracle
from particlesfm_b200 import synthetic as syn


def test_bilinear_exact_on_affine_field():
    h, w = 9, 13
    rr, cc = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    flow = np.stack([0.5 + 0.25 * cc - 0.125 * rr, -1.0 + 0.5 * cc + 2.0 * rr], -1).astype(np.float32)
    for r, c in [(2.25, 3.5), (0.0, 0.0), (7.999, 11.5), (4.0, 6.0)]:
        f, dr, dc = oracle.bilinear(flow, r, c)
        assert np.allclose(f, [0.5 + 0.25 * c - 0.125 * r, -1.0 + 0.5 * c + 2.0 * r], atol=1e-12)
        assert np.allclose(dr, [-0.125, 2.0]) and np.allclose(dc, [0.25, 0.5])


def test_bilinear_clamps_like_grid2d():
    # ceres::Grid2D clamps indices; outside the map the value freezes and derivatives vanish
    h, w = 4, 5
    flow = np.random.default_rng(0).normal(size=(h, w, 2)).astype(np.float32)
    f, dr, dc = oracle.bilinear(flow, -2.3, -7.1)
    assert np.array_equal(f, flow[0, 0].astype(np.float64)) and not dr.any() and not dc.any()
    f, dr, dc = oracle.bilinear(flow, h + 3.5, w + 0.25)
    assert np.array_equal(f, flow[h - 1, w - 1].astype(np.float64)) and not dr.any() and not dc.any()
    # last row: row+1 clamps to row -> d/dr = 0 but d/dc is the row's slope
    f, dr, dc = oracle.bilinear(flow, h - 1.0, 1.5)
    assert np.allclose(dr, 0) and np.allclose(dc, flow[h - 1, 2].astype(float) - flow[h - 1, 1])


def test_residual_definition_and_jacobian():
    uv12, r1, r2, sc, f12 = syn.make_traj_inputs(200, 40, 60, seed=3)
    uv12 = uv12 + 0.137          # stay off the integer lattice (piecewise-bilinear kinks)
    r, J = oracle.traj_evaluate(uv12, r1, r2, sc, f12)
    fl = syn.bilinear_zeros(f12, uv12[:, :2])
    exp = np.concatenate([uv12[:, :2] - r1, (uv12[:, 2:] - r2) * sc, (uv12[:, 2:] - uv12[:, :2]) - fl], 1)
    assert np.allclose(r, exp, atol=1e-10)        # path_consistency_cost.h:49-57
    eps = 1e-6
    for k in range(4):
        d = np.zeros(4); d[k] = eps
        rp, _ = oracle.traj_evaluate(uv12 + d, r1, r2, sc, f12)
        rm, _ = oracle.traj_evaluate(uv12 - d, r1, r2, sc, f12)
        assert np.allclose((rp - rm) / (2 * eps), J[:, :, k], atol=1e-6)


def test_constant_flow_closed_form():
    h, w, n = 40, 60, 300
    rng = np.random.default_rng(0)
    c = np.array([1.5, -0.75])
    flow = np.tile(c.astype(np.float32), (h, w, 1))
    x1 = rng.uniform(5, 30, (n, 2)); x2 = x1 + c + rng.normal(0, 0.3, (n, 2))
    a = x1 + rng.normal(0, 0.2, (n, 2)); b = x2 + rng.normal(0, 0.2, (n, 2))
    out, s = oracle.traj_optimize(np.concatenate([x1, x2], 1), a, b, np.ones((n, 1)), flow)
    assert np.abs(out[:, :2] - (2 * (a - c) + (b + c)) / 3).max() < 1e-6
    assert np.abs(out[:, 2:] - ((a - c) + 2 * (b + c)) / 3).max() < 1e-6
    assert s.num_iterations <= 3          # a linear problem: one Gauss-Newton step + the stopping test
    # weight 0 on the flow02 term: x1 = a, x2 = a + c exactly
    out, _ = oracle.traj_optimize(np.concatenate([x1, x2], 1), a, b, np.zeros((n, 1)), flow)
    assert np.abs(out[:, :2] - a).max() < 1e-6 and np.abs(out[:, 2:] - (a + c)).max() < 1e-6


def test_matches_scipy_optimum():
    uv12, r1, r2, sc, f12 = syn.make_traj_inputs(40, 48, 64, seed=5)
    o = oracle.traj_default_options()
    o.function_tolerance = 1e-14; o.parameter_tolerance = 1e-14; o.gradient_tolerance = 1e-14
    out, s = oracle.traj_optimize(uv12, r1, r2, sc, f12, o)

    def fun(x):
        return oracle.traj_evaluate(x.reshape(-1, 4), r1, r2, sc, f12)[0].ravel()
    ref = scipy.optimize.least_squares(fun, uv12.ravel(), method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14)
    assert abs(2 * s.final_cost - 2 * ref.cost) <= 1e-8 * max(ref.cost, 1e-12) + 1e-12
    assert np.abs(out.ravel() - ref.x).max() < 1e-4


def test_thread_and_order_invariance():
    uv12, r1, r2, sc, f12 = syn.make_traj_inputs(3000, 96, 128, seed=7)
    a, sa = oracle.traj_optimize(uv12, r1, r2, sc, f12, num_threads=1)
    b, sb = oracle.traj_optimize(uv12, r1, r2, sc, f12, num_threads=4)
    assert np.array_equal(a, b) and sa.num_iterations == sb.num_iterations   # canonical sum
    c, sc_ = oracle.traj_optimize(uv12, r1, r2, sc, f12, reduction_mode=1)    # plain sums
    assert sc_.num_iterations == sa.num_iterations
    assert np.abs(a - c).max() < 1e-9     # decision margins: summation order does not flip anything here


def test_cost_decreases_and_inputs_untouched():
    uv12, r1, r2, sc, f12 = syn.make_traj_inputs(500, 64, 64, seed=9)
    keep = uv12.copy()
    out, s = oracle.traj_optimize(uv12, r1, r2, sc, f12)
    assert s.final_cost <= s.initial_cost and np.array_equal(uv12, keep)
    r, _ = oracle.traj_evaluate(out, r1, r2, sc, f12)
    assert abs(0.5 * (r ** 2).sum() - s.final_cost) <= 1e-9 * s.final_cost
    out0, _ = oracle.traj_optimize(np.zeros((0, 4)), np.zeros((0, 2)), np.zeros((0, 2)), np.zeros(0), f12)
    assert out0.shape == (0, 4)
