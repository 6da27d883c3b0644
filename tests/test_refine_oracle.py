"""Known-answer tests that pin oracle/refine_oracle.py (the numpy restatement of the filters,
Normalize and the refinement loop around the global BA) to hand-computed cases taken from the
reference's code paths (base/reconstruction.cc:373-468, 697-729, 1321-1434)."""
import numpy as np

from oracle import refine_oracle as ro
from particlesfm_b200 import _abi, synthetic as syn


def _prob(centres, X, obs, f=100.0):
    """identity rotations, camera centres given: t = -c; obs = list of (image, point)."""
    F = len(centres)
    q = np.tile([1.0, 0, 0, 0], (F, 1))
    t = -np.asarray(centres, np.float64)
    oi = np.array([o[0] for o in obs], np.int32)
    op = np.array([o[1] for o in obs], np.int32)
    X = np.asarray(X, np.float64)
    p = X[op] + t[oi]
    xy = np.stack([f * p[:, 0] / p[:, 2], f * p[:, 1] / p[:, 2]], 1)
    return _abi.BAProblem(q, t, X, np.array([[f, 0.0, 0.0]]), oi, op, xy, np.zeros(F, np.int32))


def test_negative_depth_closed_form_counts():
    # cameras on the x axis looking along +z; point 0 in front of everyone, point 1 BEHIND cameras 0, 1
    # (track length 3 -> 2 negatives -> whole point deleted, 2 calls), point 2 behind camera 2 only
    # (track length 3 -> one observation deleted, 1 call), point 3 track length 2 with one negative
    # (-> point deleted, 1 call)
    cen = [[0, 0, 0], [1, 0, 0], [2, 0, 5.0]]
    X = [[0.5, 0, 10], [0.5, 0, -3.0], [0.5, 0.2, 4.0], [0.2, 0.1, 4.5]]
    obs = [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1), (2, 1), (0, 2), (1, 2), (2, 2), (0, 3), (2, 3)]
    prob = _prob(cen, X, obs)
    alive = np.ones(len(obs), bool)
    a2, n = ro.filter_negative_depth(prob, alive)
    # point 1: z = -3 for cams 0,1 (behind), cam 2 at z=5: -3-5 <0 too -> 3 negatives, L=3 -> calls = min(3, 2) = 2
    # point 2: z=4 ok for cams 0,1; cam 2: 4-5 = -1 -> 1 negative, L=3 -> obs deleted, 1 call
    # point 3: cam 0 ok, cam 2: 4.5-5 <0 -> L=2 -> point deleted, 1 call
    assert n == 2 + 1 + 1
    assert a2.tolist() == [True, True, True, False, False, False, True, True, False, False, False]


def test_depth_exactly_zero_is_negative():
    prob = _prob([[0, 0, 0], [0, 0, 1.0], [1, 0, 0]], [[0.1, 0.1, 1.0]], [(0, 0), (1, 0), (2, 0)])
    prob.obs_xy[:] = 0.0    # projection undefined for camera 1 (depth 0): value irrelevant
    a2, n = ro.filter_negative_depth(prob, np.ones(3, bool))
    assert n == 1 and a2.tolist() == [True, False, True]      # HasPointPositiveDepth: depth >= eps


def test_reprojection_filter_rules():
    cen = [[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]]
    X = [[0.5, 0, 10], [1.5, 0.3, 8], [0.2, 0.2, 6]]
    obs = [(0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (1, 1), (2, 1), (0, 2), (1, 2)]
    prob = _prob(cen, X, obs)
    prob.obs_xy[1] += [5.0, 0.0]           # point 0: one outlier of 4 -> observation deleted
    prob.obs_xy[4] += [0.0, 4.5]           # point 1: two outliers of 3 -> d >= L - 1 -> point deleted (3 counted)
    prob.obs_xy[5] += [4.01, 0.0]
    prob.obs_xy[7] += [3.0, 0.0]           # point 2: error 3 px < 4: kept, mean error (3 + 0) / 2
    a2, n, err = ro.filter_large_reprojection_error(prob, np.ones(9, bool), 4.0)
    assert n == 1 + 3
    assert a2.tolist() == [True, False, True, True, False, False, False, True, True]
    assert abs(err[0]) < 1e-9 and np.isnan(err[1]) and abs(err[2] - 1.5) < 1e-9
    # exactly at the threshold is kept (strict >)
    prob.obs_xy[2] += [4.0, 0.0]
    a3, n3, _ = ro.filter_large_reprojection_error(prob, a2.copy(), 4.0)
    assert a3[2] and n3 == 0


def test_triangulation_angle_values_and_filter():
    c1, c2 = np.array([0.0, 0, 0]), np.array([2.0, 0, 0])
    assert abs(ro.triangulation_angle(c1, c2, np.array([1.0, 0, 1.0])) - np.pi / 2) < 1e-12
    # obtuse -> pi - angle
    ang = ro.triangulation_angle(c1, c2, np.array([1.0, 0, 0.1]))
    assert abs(ang - (np.pi - 2 * np.arctan(10.0))) < 1e-12
    assert ro.triangulation_angle(c1, c1, c1) == 0.0
    # point far away: angle ~ 2/1000 rad = 0.11 deg < 1.5 -> deleted; near point kept
    prob = _prob([[0, 0, 0], [2, 0, 0]], [[1, 0, 1000.0], [1, 0, 20.0]], [(0, 0), (1, 0), (0, 1), (1, 1)])
    a2, n = ro.filter_small_triangulation_angle(prob, np.ones(4, bool), 1.5)
    assert n == 1 and a2.tolist() == [False, False, True, True]


def test_normalize_small_and_percentiles():
    # n <= 3: P0 = 0, P1 = n - 1; the three coordinates are sorted INDEPENDENTLY
    cen = np.array([[0.0, 5.0, 1.0], [4.0, 1.0, 2.0], [1.0, 2.0, 9.0]])
    prob = _prob(cen, [[1.0, 1.0, 1.0]], [(0, 0), (1, 0), (2, 0)])
    X0 = prob.xyz.copy()
    mean, scale = ro.normalize(prob)
    assert np.allclose(mean, [5.0 / 3, 8.0 / 3, 4.0])
    assert abs(scale - 10.0 / np.linalg.norm([4.0, 4.0, 8.0])) < 1e-15
    assert np.allclose(prob.xyz, (X0 - mean) * scale)
    assert np.allclose(ro.projection_centers(prob), (cen - mean) * scale)
    # n > 3: 11 cameras -> P0 = 1, P1 = 9
    rng = np.random.default_rng(0)
    cen = rng.normal(size=(11, 3)) * [10, 1, 3]
    prob = _prob(cen, [[0.0, 0, 50.0]], [(i, 0) for i in range(11)])
    mean, scale = ro.normalize(prob)
    cs = np.sort(cen.astype(np.float32), axis=0).astype(np.float64)
    assert np.allclose(mean, cs[1:10].mean(0), atol=1e-12)
    assert abs(scale - 10.0 / np.linalg.norm(cs[9] - cs[1])) < 1e-12
    # float keys: two centres that differ only below float32 resolution tie; extent from the float values
    cen = np.array([[0.0, 0, 0], [1.0 + 1e-9, 0, 0], [1.0, 0, 0], [3.0, 0, 0]])
    prob = _prob(cen, [[0.0, 0, 5.0]], [(i, 0) for i in range(4)])
    mean, scale = ro.normalize(prob)       # P0 = int(0.3) = 0, P1 = int(2.7) = 2
    assert np.allclose(mean, [2.0 / 3, 0, 0]) and abs(scale - 10.0) < 1e-12
    # degenerate extent -> scale 1
    prob = _prob([[1.0, 1, 1], [1.0, 1, 1]], [[0.0, 0, 5.0]], [(0, 0), (1, 0)])
    mean, scale = ro.normalize(prob)
    assert scale == 1.0 and np.allclose(mean, [1.0, 1, 1])


def test_refinement_loop_runs_and_stops():
    import oracle
    prob, truth = syn.make_ba_problem(12, 400, 6, seed=31)
    rng = np.random.default_rng(5)
    bad = rng.choice(prob.num_observations, 40, replace=False)
    prob.obs_xy[bad] += rng.normal(size=(40, 2)) * 30.0          # gross outliers: filtered after round 1
    o = oracle.ba_global_options(refine_rotation=True, refine_focal_length=True)
    o.linear_solver = _abi.SOLVER_EXACT_SCHUR
    alive, report, err = ro.iterative_global_refinement(prob, np.ones(prob.num_observations, bool), o, oracle.ba_solve)
    assert 2 <= len(report) <= 5
    assert report[0]["changed_observations"] >= 30 and report[-1]["changed"] < 5e-4
    assert not alive[bad].all()
    cen = ro.projection_centers(prob)
    assert syn.umeyama_ate(cen, truth["centres"]) < 0.05
