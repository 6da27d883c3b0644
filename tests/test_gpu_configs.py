"""BASELINE.json configs 2-5 (SURVEY.md 8d stand-ins, bench.py CONFIGS): per config one
scaled-down PARITY test against the oracle (same generator, fewer points / frames) and one
full-size PROPERTY test (termination, cost, gauge, accuracy against the generator's truth).
Config 2 has tracks up to 50 frames over 50 images (a DENSE reduced system, ~470 pair entries
per point: the opposite regime from the banded target workload), config 3 random drop-outs,
config 4 300 frames at 1 px noise, config 5 M = 2.4e7 on one GPU."""
import numpy as np
import pytest

import oracle
from particlesfm_b200 import _abi, ba, synthetic as syn

pytestmark = pytest.mark.gpu

SMALL = {
    "2": dict(num_images=50, num_points=1500, track_len=12, track_len_range=(3, 50), seed=1),
    "3": dict(num_images=80, num_points=3000, track_len=12, dynamic_fraction=0.3, seed=2),
    "4": dict(num_images=60, num_points=1200, track_len=15, noise_px=1.0, seed=3),
    "5": dict(num_images=70, num_points=4000, track_len=12, seed=4),
}
FULL = {
    "2": dict(num_images=50, num_points=300_000, track_len=12, track_len_range=(3, 50), seed=1),
    "3": dict(num_images=80, num_points=800_000, track_len=12, dynamic_fraction=0.3, seed=2),
    "4": dict(num_images=300, num_points=200_000, track_len=15, noise_px=1.0, seed=3),
    "5": dict(num_images=500, num_points=2_000_000, track_len=12, seed=4),
}


def _opts(rot=True, focal=True):
    o = oracle.ba_global_options(refine_rotation=rot, refine_focal_length=focal)
    o.linear_solver = _abi.SOLVER_AUTO
    return o


@pytest.mark.parametrize("cfg", sorted(SMALL))
def test_config_scaled_down_parity(gpu, cfg):
    prob, _ = syn.make_ba_problem(**SMALL[cfg])
    for rot, focal in ((False, False), (True, True)):       # pass A, pass B
        p0, p1 = prob.copy(), prob.copy()
        s0 = oracle.ba_solve(p0, _opts(rot, focal))
        s1 = ba.solve_problem(p1, _opts(rot, focal))
        assert s1.linear_solver_used == _abi.SOLVER_EXACT_SCHUR
        assert s1.termination == s0.termination and s1.num_iterations == s0.num_iterations
        assert abs(s1.final_cost - s0.final_cost) <= 1e-5 * s0.final_cost
        for a, b in ((p1.qvec, p0.qvec), (p1.tvec, p0.tvec), (p1.xyz, p0.xyz), (p1.cam_params, p0.cam_params)):
            assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()


@pytest.mark.parametrize("cfg", sorted(FULL))
def test_config_full_size_properties(gpu, cfg):
    prob, truth = syn.make_ba_problem(**FULL[cfg])
    p = prob.copy()
    s = ba.solve_problem(p, _opts())
    assert s.linear_solver_used == _abi.SOLVER_EXACT_SCHUR        # the reference rule: <= 1000 images
    assert s.num_residuals_reduced == 2 * prob.num_observations
    assert s.termination in (0, 1, 2)
    assert s.final_cost < 0.1 * s.initial_cost
    sigma = FULL[cfg].get("noise_px", 0.5)
    per_obs = s.final_cost / prob.num_observations            # SoftL1 of ~N(0, sigma^2) residual pairs
    assert 0.3 * sigma * sigma < per_obs < 1.1 * sigma * sigma
    assert np.array_equal(p.tvec[0], prob.tvec[0]) and p.tvec[1, 0] == prob.tvec[1, 0]
    ate = syn.umeyama_ate(syn.camera_centres(p.qvec, p.tvec), truth["centres"])
    ate0 = syn.umeyama_ate(syn.camera_centres(prob.qvec, prob.tvec), truth["centres"])
    assert ate < 0.2 * ate0 and ate < 5e-3
