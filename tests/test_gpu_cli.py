"""Boundary B3: `python -m particlesfm_b200.ba_cli --input_path M --output_path M'` — a COLMAP
model directory in, the two IterativeGlobalRefinement passes of the reference's controller on the
device, a COLMAP model directory out (readable by the reference's own read_write_model layout)."""
import os
import sys

import numpy as np
import pytest

from particlesfm_b200 import ba, ba_cli, colmap_io, synthetic as syn

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_ba_host import reconstruction_from_problem   # noqa: E402


def test_cli_refines_a_model_directory(gpu, tmp_path):
    prob, truth = syn.make_ba_problem(24, 2500, 7, seed=81)
    order = np.argsort(prob.obs_image, kind="stable")
    prob.obs_image, prob.obs_point, prob.obs_xy = prob.obs_image[order], prob.obs_point[order], prob.obs_xy[order]
    rng = np.random.default_rng(3)
    bad = rng.choice(prob.num_observations, 50, replace=False)
    prob.obs_xy[bad] += rng.normal(size=(50, 2)) * 40.0
    rec = reconstruction_from_problem(prob)
    src, dst = str(tmp_path / "in"), str(tmp_path / "out")
    colmap_io.write_model(rec, src)
    n_in = rec.ComputeNumObservations()
    reports = ba_cli.run(ba_cli.build_parser().parse_args(["--input_path", src, "--output_path", dst, "--quiet"]))
    assert len(reports) == 2 and all(r.num_rounds >= 1 for r in reports)
    out = colmap_io.read_model(dst)
    assert sorted(out.images) == sorted(rec.images)
    n_out = out.ComputeNumObservations()
    assert n_in - 500 < n_out <= n_in - 30            # the gross outliers (and what hung on them) are gone, little else
    assert reports[-1].final_num_observations == n_out
    # images.bin and points3D.bin agree (every track element points at a Point2D of that point)
    for pid, p in out.points3D.items():
        assert len(p.image_ids) >= 2
        for iid, j in zip(p.image_ids, p.point2D_idxs):
            assert out.images[int(iid)].point3D_ids[int(j)] == pid
    ids = sorted(out.images)
    q = np.stack([out.images[i].qvec for i in ids]); t = np.stack([out.images[i].tvec for i in ids])
    assert syn.umeyama_ate(syn.camera_centres(q, t), truth["centres"]) < 5e-3
    errs = np.array([p.error for p in out.points3D.values()])
    assert 0.2 < np.median(errs) < 1.5                # Point3D::Error = mean reprojection error, ~0.6 px at 0.5 px noise
    # the same call through the Python mirror gives the same model
    rec2 = colmap_io.read_model(src)
    for force in (False, True):
        ba.iterative_global_refinement(rec2, force)
    for i in ids:      # not bit-identical: per-image sums use fp64 atomics (summation order varies run to run)
        assert np.abs(rec2.images[i].tvec - out.images[i].tvec).max() <= 1e-7 * max(1.0, np.abs(out.images[i].tvec).max())
