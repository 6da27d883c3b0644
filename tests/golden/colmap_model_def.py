"""Deterministic small COLMAP model used by make_colmap_golden.py (written with the reference's
module) and by tests/test_colmap_io.py (compared with what particlesfm_b200.colmap_io reads)."""
import numpy as np


def model_values(seed=33):
    rng = np.random.default_rng(seed)
    cameras = {1: dict(model_id=0, width=1024, height=436, params=np.array([900.0, 512.0, 218.0])),
               4: dict(model_id=1, width=640, height=480, params=np.array([500.5, 501.25, 320.0, 240.0]))}
    n_pts = 40
    pids = np.sort(rng.choice(np.arange(1, 500), n_pts, replace=False)).astype(np.int64)
    images = {}
    tracks = {int(p): [] for p in pids}
    for k, iid in enumerate([3, 1, 7, 8, 12]):
        m = [0, 17, 30, 5, 23][k]
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        xys = rng.uniform(0, 400, (m, 2))
        ids = np.full(m, -1, dtype=np.int64)
        has = rng.random(m) < 0.7
        ids[has] = rng.choice(pids, int(has.sum()), replace=False) if has.sum() <= n_pts else -1
        for j in np.nonzero(ids >= 0)[0]:
            tracks[int(ids[j])].append((iid, int(j)))
        images[iid] = dict(qvec=q, tvec=rng.normal(size=3), camera_id=1 if k % 2 == 0 else 4,
                           name=f"frames/{iid:05d}.png", xys=xys, point3D_ids=ids)
    points = {}
    for p in pids:
        t = tracks[int(p)]
        points[int(p)] = dict(xyz=rng.normal(size=3) * 5, rgb=rng.integers(0, 256, 3).astype(np.uint8),
                              error=float(rng.uniform(0, 2)),
                              image_ids=np.array([a for a, _ in t], dtype=np.int32),
                              point2D_idxs=np.array([b for _, b in t], dtype=np.int32))
    return cameras, images, points
