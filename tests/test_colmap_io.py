"""COLMAP binary model I/O at the HP2 boundary: particlesfm_b200.colmap_io against files written
by the reference's own module (tests/golden/colmap_model/, made by make_colmap_golden.py from
sfm/colmap_utils/read_write_model.py:447-456).  Reading gives the generating values exactly;
writing the same model reproduces the reference's bytes."""
import os
import sys

import numpy as np

from particlesfm_b200 import ba, colmap_io

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "colmap_model")
sys.path.insert(0, os.path.join(HERE, "golden"))
from colmap_model_def import model_values   # noqa: E402


def test_read_reference_written_model():
    cams, imgs, pts = model_values()
    rec = colmap_io.read_model(GOLD)
    assert list(rec.cameras) == list(cams) and list(rec.images) == list(imgs) and list(rec.points3D) == list(pts)
    for i, c in cams.items():
        r = rec.cameras[i]
        assert (r.camera_id, r.model_id, r.width, r.height) == (i, c["model_id"], c["width"], c["height"])
        assert np.array_equal(r.params, c["params"])
    for i, m in imgs.items():
        r = rec.images[i]
        assert r.image_id == i and r.camera_id == m["camera_id"] and r.name == m["name"]
        assert np.array_equal(r.qvec, m["qvec"]) and np.array_equal(r.tvec, m["tvec"])
        assert np.array_equal(r.xys, m["xys"]) and np.array_equal(r.point3D_ids, m["point3D_ids"])
        assert r.point3D_ids.dtype == np.int64
    for i, p in pts.items():
        r = rec.points3D[i]
        assert np.array_equal(r.xyz, p["xyz"]) and np.array_equal(r.rgb, p["rgb"]) and r.error == p["error"]
        assert np.array_equal(r.image_ids, p["image_ids"]) and np.array_equal(r.point2D_idxs, p["point2D_idxs"])


def test_write_is_byte_identical_to_reference(tmp_path):
    rec = colmap_io.read_model(GOLD)
    colmap_io.write_model(rec, str(tmp_path))
    for f in ("cameras.bin", "images.bin", "points3D.bin"):
        assert open(os.path.join(str(tmp_path), f), "rb").read() == open(os.path.join(GOLD, f), "rb").read(), f


def test_model_from_disk_feeds_the_bundle_adjuster(tmp_path):
    # disk -> Reconstruction -> flattened psfm_ba_problem (what BundleAdjuster::SetUp enumerates)
    rec = colmap_io.read_model(GOLD)
    for c in rec.cameras.values():            # the pipeline's only model; the fixture's second camera is PINHOLE
        c.model_id, c.params = 0, np.asarray(c.params[:3])
    cfg = ba.BundleAdjustmentConfig()
    for i in rec.RegImageIds():
        cfg.AddImage(i)
    prob, maps = ba.flatten(rec, cfg)
    n_obs = sum(int((im.point3D_ids >= 0).sum()) for im in rec.images.values())
    assert prob.num_observations == n_obs and prob.num_images == len(rec.images)
    assert prob.num_cameras == 2 and prob.obs_image.max() < prob.num_images
    # scatter back unchanged and round-trip through the writer
    ba.scatter(prob, maps, rec)
    colmap_io.write_model(rec, str(tmp_path))
    again = colmap_io.read_model(str(tmp_path))
    assert all(np.array_equal(again.points3D[k].xyz, rec.points3D[k].xyz) for k in rec.points3D)
