"""BASELINE.json config 1 — "20-frame synthetic static scene, path consistency + global BA on the
CPU path (plumbing, no GPU)": flows of a static slanted plane seen by a translating camera ->
batched tracker with the path-consistency optimiser (oracle) -> trajectories -> hand-off
(handoff.tracks_to_observations) -> global BA (oracle) from a perturbed start.  Everything
between the flow fields and the refined poses runs through this repo's host code; the two
solvers are the CPU oracle (test infrastructure), exactly as SURVEY.md §8(d) defines config 1."""
import numpy as np

import oracle
from particlesfm_b200 import _abi, handoff, synthetic as syn, tracker

F, H, W, FOCAL = 12, 64, 96, 120.0
K = np.array([[FOCAL, 0, W / 2], [0, FOCAL, H / 2], [0, 0, 1.0]])
PLANE_N, PLANE_D = np.array([0.15, -0.1, 1.0]), 6.0            # n . X = d  (slanted, in front of all cameras)


def _poses():
    q, t = [], []
    for i in range(F):
        ang = 0.004 * i
        R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        C = np.array([0.05 * i, 0.01 * i, 0.0])
        q.append(syn.rotmat_to_qvec(R)); t.append(-R @ C)
    return np.array(q), np.array(t)


def _homography(Ra, ta, Rb, tb):
    """pixel in a -> pixel in b for points on the world plane n.X = d."""
    R = Rb @ Ra.T
    t = tb - R @ ta
    na = Ra @ PLANE_N                       # plane in camera a: na . Xa = d + na . ta
    da = PLANE_D + na @ ta
    return K @ (R + np.outer(t, na) / da) @ np.linalg.inv(K)


def _flow(Hm):
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    p = Hm @ np.stack([xx.ravel(), yy.ravel(), np.ones(H * W)])
    return np.stack([p[0] / p[2] - xx.ravel(), p[1] / p[2] - yy.ravel()], 1).reshape(H, W, 2).astype(np.float32)


def _oracle_opt(uv12, r1, r2, sc, fmap, n, w, h):
    return oracle.traj_optimize(uv12, r1, r2, sc, fmap)[0]


def test_flows_to_refined_poses():
    q, t = _poses()
    R = syn.qvec_to_rotmat(q)
    Hm = lambda a, b: _homography(R[a], t[a], R[b], t[b])
    fw = [_flow(Hm(i, i + 1)) for i in range(F - 1)]
    fb = [_flow(Hm(i + 1, i)) for i in range(F - 1)]
    f2 = [_flow(Hm(i, i + 2)) for i in range(F - 2)]
    b2 = [_flow(Hm(i + 2, i)) for i in range(F - 2)]
    # ---- HP1 stage: tracker + path-consistency optimiser
    tracks = tracker.main_connect_point_trajectories(fw, fb, f2, b2, 4, 1.0, 3, optimize_fn=_oracle_opt)
    assert len(tracks) > 150
    lens = np.array([len(v["frame_ids"]) for v in tracks.values()])
    assert lens.min() >= 3 and lens.max() == F                 # static scene: seeds of frame 0 survive to the end
    # ---- hand-off: one 3D point per trajectory
    img, pt, xy, keys = handoff.tracks_to_observations(tracks, remove_dynamic=True, min_length=3)
    assert np.unique(pt).shape[0] == len(keys) == len(tracks)
    # truth of every point: back-project its first sample onto the plane
    first = np.array([np.nonzero(pt == p)[0][0] for p in range(len(keys))])
    ray_c = np.linalg.inv(K) @ np.vstack([xy[first].T, np.ones(len(keys))])
    Xw = []
    for k, i in enumerate(first):
        Ri, ti = R[img[i]], t[img[i]]
        d_w, C = Ri.T @ ray_c[:, k], -Ri.T @ ti
        lam = (PLANE_D - PLANE_N @ C) / (PLANE_N @ d_w)
        Xw.append(C + lam * d_w)
    Xw = np.array(Xw)
    # the tracked samples of a point reproject onto its plane point within tracking accuracy
    proj = np.einsum("nij,nj->ni", R[img], Xw[pt]) + t[img]
    uv = (K @ proj.T).T
    err = np.linalg.norm(uv[:, :2] / uv[:, 2:3] - xy, axis=1)
    assert np.median(err) < 0.05 and err.max() < 1.0
    # ---- HP2 stage from a perturbed start
    rng = np.random.default_rng(4)
    C0 = syn.camera_centres(q, t)
    q0, t0 = q.copy(), t.copy()
    for i in range(2, F):
        dq = np.concatenate([[1.0], rng.normal(0, 0.002, 3)])
        Rn = syn.qvec_to_rotmat(dq / np.linalg.norm(dq)) @ R[i]
        q0[i] = syn.rotmat_to_qvec(Rn)
        t0[i] = -Rn @ (C0[i] + rng.normal(0, 0.01, 3))
    X0 = Xw + rng.normal(0, 0.03, Xw.shape)
    pose_const = np.zeros(F, np.uint8); pose_const[0] = 1                      # gauge as the reference sets it
    tmask = np.zeros(F, np.uint8); tmask[1] = 1
    prob = _abi.BAProblem(q0, t0, X0, np.array([[FOCAL, W / 2, H / 2]]), img, pt, xy, np.zeros(F, np.int32),
                          pose_const, tmask, np.zeros(1, np.uint8))
    o = oracle.ba_global_options(refine_rotation=True, refine_focal_length=False)
    ate0 = syn.umeyama_ate(syn.camera_centres(prob.qvec, prob.tvec), C0)
    s = oracle.ba_solve(prob, o)
    ate1 = syn.umeyama_ate(syn.camera_centres(prob.qvec, prob.tvec), C0)
    assert s.termination in (0, 1, 2) and s.final_cost < 0.02 * s.initial_cost
    assert ate1 < 0.25 * ate0 and ate1 < 2e-3
