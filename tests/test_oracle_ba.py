"""Known-answer tests that pin the HP2 oracle: closed-form and finite-difference Jacobians
(incl. the quaternion manifold), loss corrector, dense normal-equation solve vs the Schur
path, PCG == Cholesky, zero-noise recovery, gauge, scipy optimum."""
import numpy as np
import pytest
import scipy.optimize

import oracle
from particlesfm_b200 import _abi, synthetic as syn


def _free_problem(F=6, P=60, L=4, seed=1):
    prob, truth = syn.make_ba_problem(F, P, L, seed=seed)
    return prob, truth


def _qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def test_jacobians_match_closed_form():
    """Oracle: ambient 2x4 Jacobian of UnitQuaternionRotatePoint times the 4x3 plus
    Jacobian (Ceres' route).  Here: the closed form dr/dp * (-2 [R X]x) the CUDA kernel uses."""
    prob, _ = _free_problem()
    prob.pose_constant[:] = 0; prob.tvec_constant_mask[:] = 0
    o = oracle.ba_global_options(True, True); o.refine_principal_point = 1
    jc, jp, jk = oracle.ba_jacobians(prob, o)
    q = prob.qvec / np.linalg.norm(prob.qvec, axis=1, keepdims=True)
    R = syn.qvec_to_rotmat(q)[prob.obs_image]
    w = np.einsum("mij,mj->mi", R, prob.xyz[prob.obs_point]); p = w + prob.tvec[prob.obs_image]
    f, cx, cy = prob.cam_params[0]
    iz = 1 / p[:, 2]; u = p[:, 0] * iz; v = p[:, 1] * iz
    e = np.stack([f * u + cx, f * v + cy], -1) - prob.obs_xy
    sq = (1 + (e ** 2).sum(1)) ** -0.25                   # sqrt(rho'), SoftL1(1)
    a00, a02, a12 = f * iz, -f * u * iz, -f * v * iz
    Jr0 = np.stack([2 * a02 * w[:, 1], 2 * (a00 * w[:, 2] - a02 * w[:, 0]), -2 * a00 * w[:, 1]], -1)
    Jr1 = np.stack([2 * (a12 * w[:, 1] - a00 * w[:, 2]), -2 * a12 * w[:, 0], 2 * a00 * w[:, 0]], -1)
    z = 0 * a00
    JC = sq[:, None, None] * np.stack([np.concatenate([Jr0, np.stack([a00, z, a02], -1)], -1),
                                       np.concatenate([Jr1, np.stack([z, a00, a12], -1)], -1)], 1)
    JP = sq[:, None, None] * np.stack([a00[:, None] * R[:, 0] + a02[:, None] * R[:, 2],
                                       a00[:, None] * R[:, 1] + a12[:, None] * R[:, 2]], 1)
    JK = sq[:, None, None] * np.stack([np.stack([u, 1 + z, z], -1), np.stack([v, z, 1 + z], -1)], 1)
    for a, b in ((JC, jc), (JP, jp), (JK, jk)):
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()


@pytest.mark.parametrize("loss", [_abi.LOSS_TRIVIAL, _abi.LOSS_SOFT_L1, _abi.LOSS_CAUCHY])
def test_gradient_matches_finite_differences(loss):
    """d(1/2 sum rho(|r|^2))/dx == J~' r~ with the Corrector's first-order scaling."""
    prob, _ = _free_problem()
    prob.pose_constant[:] = 0; prob.tvec_constant_mask[:] = 0
    o = oracle.ba_global_options(True, True); o.refine_principal_point = 1
    o.loss_function_type = loss
    _, _, gc, gp = oracle.ba_evaluate(prob, o)
    cost = lambda p: oracle.ba_evaluate(p, o)[0]
    eps = 1e-6
    F = prob.num_images
    qn = prob.qvec / np.linalg.norm(prob.qvec, axis=1, keepdims=True)
    for k in range(3):
        p1, p2 = prob.copy(), prob.copy(); p1.xyz[7, k] += eps; p2.xyz[7, k] -= eps
        assert np.isclose((cost(p1) - cost(p2)) / (2 * eps), gp[7, k], rtol=1e-5, atol=1e-6)
        p1, p2 = prob.copy(), prob.copy(); p1.tvec[2, k] += eps; p2.tvec[2, k] -= eps
        assert np.isclose((cost(p1) - cost(p2)) / (2 * eps), gc[6 * 2 + 3 + k], rtol=1e-5, atol=1e-5)
        p1, p2 = prob.copy(), prob.copy(); p1.cam_params[0, k] += eps; p2.cam_params[0, k] -= eps
        assert np.isclose((cost(p1) - cost(p2)) / (2 * eps), gc[6 * F + k], rtol=1e-5, atol=1e-5)
        d = np.zeros(3); d[k] = eps          # QuaternionParameterization::Plus: q+ = exp(d) * q
        qd = np.array([np.cos(eps), *(np.sin(eps) / eps * d)])
        p1, p2 = prob.copy(), prob.copy()
        p1.qvec[3] = _qmul(qd, qn[3]); qd[1:] *= -1; p2.qvec[3] = _qmul(qd, qn[3])
        assert np.isclose((cost(p1) - cost(p2)) / (2 * eps), gc[6 * 3 + k], rtol=1e-5, atol=1e-5)


def _dense_step(prob, o, radius):
    """(J'J + D^2) y = J'r solved densely in numpy with Ceres' Jacobi scaling and LM
    diagonal; returns the scaled-space step in the oracle's slot layout."""
    jc, jp, jk = oracle.ba_jacobians(prob, o)
    _, r, _, _ = oracle.ba_evaluate(prob, o)
    F, P, M = prob.num_images, prob.num_points, prob.num_observations
    NS = 6 * F + 3
    J = np.zeros((2 * M, NS + 3 * P))
    for i in range(M):
        im, pt = prob.obs_image[i], prob.obs_point[i]
        J[2 * i:2 * i + 2, 6 * im:6 * im + 6] = jc[i]
        J[2 * i:2 * i + 2, 6 * F:6 * F + 3] = jk[i]
        J[2 * i:2 * i + 2, NS + 3 * pt:NS + 3 * pt + 3] = jp[i]
    act = np.abs(J).sum(0) > 0
    Ja = J[:, act]
    scale = 1.0 / (1.0 + np.sqrt((Ja ** 2).sum(0)))
    Js = Ja * scale
    diag = np.clip((Js ** 2).sum(0), 1e-6, 1e32)
    y = np.linalg.solve(Js.T @ Js + np.diag(diag / radius), Js.T @ r.ravel())
    full = np.zeros(NS + 3 * P)
    full[act] = -y
    return full[:NS], full[NS:].reshape(P, 3)


@pytest.mark.parametrize("rot,focal", [(False, False), (True, True)])
def test_schur_step_equals_dense_normal_equations(rot, focal):
    prob, _ = syn.make_ba_problem(5, 40, 4, seed=2)
    o = oracle.ba_global_options(rot, focal)
    for radius in (1e4, 3.0):
        sc0, sp0 = _dense_step(prob, o, radius)
        sc1, sp1, _ = oracle.ba_linear_step(prob, o, radius, _abi.SOLVER_EXACT_SCHUR)
        assert np.abs(sc1 - sc0).max() <= 1e-8 * np.abs(sc0).max()
        assert np.abs(sp1 - sp0).max() <= 1e-8 * np.abs(sp0).max()


def test_tight_pcg_equals_cholesky_step():
    prob, _ = syn.make_ba_problem(8, 300, 5, seed=3)
    o = oracle.ba_global_options(True, True)
    sc0, sp0, _ = oracle.ba_linear_step(prob, o, 1e4, _abi.SOLVER_EXACT_SCHUR)
    o.eta = 1e-30; o.max_linear_solver_iterations = 2000
    sc1, sp1, it = oracle.ba_linear_step(prob, o, 1e4, _abi.SOLVER_ITERATIVE_SCHUR)
    assert np.abs(sc1 - sc0).max() <= 1e-8 * np.abs(sc0).max()
    assert np.abs(sp1 - sp0).max() <= 1e-8 * np.abs(sp0).max()
    o.eta = 0.1
    _, _, it_loose = oracle.ba_linear_step(prob, o, 1e4, _abi.SOLVER_ITERATIVE_SCHUR)
    assert it_loose < it


def test_zero_noise_recovers_truth_and_gauge_is_untouched():
    prob, truth = syn.make_ba_problem(10, 400, 6, seed=4, noise_px=0.0)
    start = prob.copy()
    o = oracle.ba_global_options(True, False); o.linear_solver = _abi.SOLVER_EXACT_SCHUR
    o.function_tolerance = 1e-14; o.gradient_tolerance = 1e-12; o.parameter_tolerance = 1e-14
    o.max_num_iterations = 60
    s = oracle.ba_solve(prob, o)
    assert s.final_cost < 1e-6 * s.initial_cost
    assert syn.umeyama_ate(syn.camera_centres(prob.qvec, prob.tvec), truth["centres"]) < 1e-5
    q0 = start.qvec[0] / np.linalg.norm(start.qvec[0])
    assert np.array_equal(prob.qvec[0], q0) and np.array_equal(prob.tvec[0], start.tvec[0])
    assert prob.tvec[1, 0] == start.tvec[1, 0]
    assert np.array_equal(prob.cam_params, start.cam_params)      # intrinsics constant in this pass


def test_exact_and_iterative_reach_the_same_minimum():
    prob, truth = syn.make_ba_problem(12, 800, 6, seed=5)
    o = oracle.ba_global_options(True, True)
    pa, pb = prob.copy(), prob.copy()
    o.linear_solver = _abi.SOLVER_EXACT_SCHUR
    sa = oracle.ba_solve(pa, o)
    o.linear_solver = _abi.SOLVER_ITERATIVE_SCHUR
    sb = oracle.ba_solve(pb, o)
    assert sa.final_cost < 0.1 * sa.initial_cost
    assert abs(sa.final_cost - sb.final_cost) <= 1e-3 * sa.final_cost
    assert sb.num_linear_iterations > sb.num_iterations


def test_matches_scipy_optimum_trivial_loss():
    prob, _ = syn.make_ba_problem(4, 25, 3, seed=6)
    o = oracle.ba_global_options(False, False)
    o.loss_function_type = _abi.LOSS_TRIVIAL
    o.function_tolerance = 1e-15; o.gradient_tolerance = 1e-12; o.parameter_tolerance = 1e-15
    o.max_num_iterations = 200; o.linear_solver = _abi.SOLVER_EXACT_SCHUR
    p = prob.copy()
    s = oracle.ba_solve(p, o)
    # free parameters of pass A: tvec of images 1.. (minus tvec[1][0]) and all points
    tm = np.ones((prob.num_images, 3), bool); tm[0] = False; tm[1, 0] = False

    def fun(x):
        q = prob.copy()
        q.tvec[tm] = x[:tm.sum()]
        q.xyz[:] = x[tm.sum():].reshape(-1, 3)
        return oracle.ba_evaluate(q, o)[1].ravel()
    x0 = np.concatenate([prob.tvec[tm], prob.xyz.ravel()])
    ref = scipy.optimize.least_squares(fun, x0, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-12)
    assert abs(s.final_cost - ref.cost) <= 1e-7 * ref.cost


def test_block_sparse_schur_matches_dense_assembly(monkeypatch):
    """The SPARSE_SCHUR restatement (oracle/ba_schur_blocks.h: block-sparse S, band + arrow
    Cholesky) against the general dense assembly + dense Cholesky, on a banded (video) and on a
    dense co-visibility pattern, pass A and pass B."""
    for (F, P, L, rng) in ((40, 1200, 5, None), (14, 500, 6, (2, 14))):
        prob, _ = syn.make_ba_problem(F, P, L, seed=21, track_len_range=rng)
        for rot, focal in ((False, False), (True, True)):
            o = oracle.ba_global_options(refine_rotation=rot, refine_focal_length=focal)
            o.linear_solver = _abi.SOLVER_EXACT_SCHUR
            monkeypatch.delenv("PSFM_ORACLE_DENSE_SCHUR", raising=False)
            sc0, sp0, _ = oracle.ba_linear_step(prob, o, 1e3, _abi.SOLVER_EXACT_SCHUR)
            monkeypatch.setenv("PSFM_ORACLE_DENSE_SCHUR", "1")
            sc1, sp1, _ = oracle.ba_linear_step(prob, o, 1e3, _abi.SOLVER_EXACT_SCHUR)
            monkeypatch.delenv("PSFM_ORACLE_DENSE_SCHUR", raising=False)
            assert np.abs(sc0 - sc1).max() <= 1e-9 * np.abs(sc1).max()
            assert np.abs(sp0 - sp1).max() <= 1e-9 * np.abs(sp1).max()
