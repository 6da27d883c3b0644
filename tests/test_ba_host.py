"""Host mirror of the reference's Reconstruction methods the global BA brackets itself with
(particlesfm_b200/ba.py: Normalize, FilterObservationsWithNegativeDepth, DeleteObservation,
flatten / scatter / apply_observation_mask) against oracle/refine_oracle.py, whose closed forms are
pinned by hand-computed cases in tests/test_refine_oracle.py.  No GPU needed."""
import numpy as np
import pytest

from oracle import refine_oracle as ro
from particlesfm_b200 import ba, synthetic as syn


def reconstruction_from_problem(prob, image_id0=1, point_id0=10):
    """A Reconstruction holding the same geometry as a flattened BAProblem (ids offset, tracks filled)."""
    F, P = prob.num_images, prob.num_points
    cams = {1: ba.Camera(1, ba.SIMPLE_PINHOLE, 1024, 436, prob.cam_params[0].copy())}
    images, points = {}, {}
    per_img = [np.nonzero(prob.obs_image == i)[0] for i in range(F)]
    tracks = {p: [] for p in range(P)}
    for i in range(F):
        sel = per_img[i]
        images[image_id0 + i] = ba.Image(image_id0 + i, prob.qvec[i].copy(), prob.tvec[i].copy(), 1, f"f{i:04d}.png",
                                         prob.obs_xy[sel].copy(), (prob.obs_point[sel] + point_id0).astype(np.int64))
        for j, m in enumerate(sel):
            tracks[int(prob.obs_point[m])].append((image_id0 + i, j))
    for p in range(P):
        t = tracks[p]
        if not t:
            continue
        points[point_id0 + p] = ba.Point3D(point_id0 + p, prob.xyz[p].copy(), np.zeros(3, np.uint8), 0.0,
                                           np.array([a for a, _ in t], np.int32), np.array([b for _, b in t], np.int32))
    return ba.Reconstruction(cams, images, points)


def _alive_from(rec, prob, image_id0=1):
    """alive mask over prob's observations, read back from the container."""
    alive = np.zeros(prob.num_observations, bool)
    per_img_count = {}
    for m in range(prob.num_observations):
        i = int(prob.obs_image[m])
        j = per_img_count.get(i, 0)
        per_img_count[i] = j + 1
        alive[m] = rec.images[image_id0 + i].point3D_ids[j] >= 0
    return alive


def _sorted_by_image(prob):
    order = np.argsort(prob.obs_image, kind="stable")
    from particlesfm_b200._abi import BAProblem
    return BAProblem(prob.qvec, prob.tvec, prob.xyz, prob.cam_params, prob.obs_image[order], prob.obs_point[order],
                     prob.obs_xy[order], prob.image_camera, prob.pose_constant, prob.tvec_constant_mask, prob.camera_constant)


def test_negative_depth_filter_and_track_bookkeeping():
    prob, _ = syn.make_ba_problem(9, 300, 4, seed=61, track_len_range=(2, 6))
    prob = _sorted_by_image(prob)
    rng = np.random.default_rng(1)
    cen = syn.camera_centres(prob.qvec, prob.tvec)
    for p in rng.choice(300, 25, replace=False):
        i = prob.obs_image[np.nonzero(prob.obs_point == p)[0][0]]
        prob.xyz[p] = 1.5 * cen[i]
    rec = reconstruction_from_problem(prob)
    n = rec.FilterObservationsWithNegativeDepth()
    a_ref, n_ref = ro.filter_negative_depth(prob, np.ones(prob.num_observations, bool))
    assert n == n_ref and n > 0
    assert np.array_equal(_alive_from(rec, prob), a_ref)
    # the track elements go with the observations — points3D and images agree
    for pid, p in rec.points3D.items():
        assert len(p.image_ids) >= 2
        for iid, j in zip(p.image_ids, p.point2D_idxs):
            assert rec.images[int(iid)].point3D_ids[int(j)] == pid
    assert rec.ComputeNumObservations() == sum(len(p.image_ids) for p in rec.points3D.values()) == int(a_ref.sum())


@pytest.mark.parametrize("F", [2, 3, 11, 40])
def test_normalize_matches_oracle(F):
    prob, _ = syn.make_ba_problem(F, 120, 2, seed=70 + F)
    prob = _sorted_by_image(prob)
    rec = reconstruction_from_problem(prob)
    ref = prob.copy()
    mean, scale = rec.Normalize()
    m0, s0 = ro.normalize(ref)
    assert np.allclose(mean, m0, rtol=0, atol=1e-12) and abs(scale - s0) <= 1e-13 * s0
    for i in range(F):
        assert np.abs(rec.images[1 + i].tvec - ref.tvec[i]).max() <= 1e-12 * max(1.0, np.abs(ref.tvec).max())
    for p, pt in rec.points3D.items():
        assert np.abs(pt.xyz - ref.xyz[p - 10]).max() <= 1e-12 * max(1.0, np.abs(ref.xyz).max())
    one = ba.Reconstruction({}, {1: rec.images[1]}, {})
    assert one.Normalize() is None           # < 2 registered images: no-op (reconstruction.cc:382-385)


def test_flatten_uses_registration_order_and_mask_roundtrip():
    prob, _ = syn.make_ba_problem(6, 80, 3, seed=77)
    prob = _sorted_by_image(prob)
    rec = reconstruction_from_problem(prob)
    rec.reg_image_ids = [4, 2, 6, 1, 3, 5]          # registration order != id order
    cfg = ba.BundleAdjustmentConfig()
    for i in rec.RegImageIds():
        cfg.AddImage(i)
    cfg.SetConstantPose(rec.RegImageIds()[0])
    cfg.SetConstantTvec(rec.RegImageIds()[1], [0])
    flat, maps = ba.flatten(rec, cfg)
    assert maps["image_ids"] == [4, 2, 6, 1, 3, 5]
    assert flat.pose_constant.tolist() == [1, 0, 0, 0, 0, 0] and flat.tvec_constant_mask.tolist() == [0, 1, 0, 0, 0, 0]
    assert flat.num_observations == prob.num_observations
    # kill all observations of one point and one observation of another: container follows the mask
    alive = np.ones(flat.num_observations, bool)
    p_dead, p_cut = 3, 5
    alive[flat.obs_point == p_dead] = False
    cut = np.nonzero(flat.obs_point == p_cut)[0][0]
    alive[cut] = False
    n_cut_before = len(rec.points3D[maps["pt_ids"][p_cut]].image_ids)
    ba.apply_observation_mask(flat, maps, rec, alive, np.full(flat.num_points, 0.25))
    assert maps["pt_ids"][p_dead] not in rec.points3D
    pc = rec.points3D[maps["pt_ids"][p_cut]]
    assert len(pc.image_ids) == n_cut_before - 1 and pc.error == 0.25
    assert rec.ComputeNumObservations() == int(alive.sum())
