"""Seeded synthetic workloads of the shapes named in BASELINE.json / SURVEY.md §8(d).

No dataset ships with the repo (no network): flows, trajectories and bundle-adjustment
problems are generated from geometry.  Used by tests/, bench.py and __graft_entry__.
"""
import numpy as np

from ._abi import BAProblem


# ----------------------------------------------------------------------------- rotations

def qvec_to_rotmat(q):
    """COLMAP convention, q = (w, x, y, z), vectorised over leading dims."""
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - w * z)
    R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y)
    R[..., 2, 1] = 2 * (y * z + w * x)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rotmat_to_qvec(R):
    R = np.asarray(R, dtype=np.float64)
    flat = R.reshape(-1, 3, 3)
    out = np.empty((flat.shape[0], 4))
    for i, m in enumerate(flat):
        K = np.array([
            [m[0, 0] - m[1, 1] - m[2, 2], 0, 0, 0],
            [m[1, 0] + m[0, 1], m[1, 1] - m[0, 0] - m[2, 2], 0, 0],
            [m[2, 0] + m[0, 2], m[2, 1] + m[1, 2], m[2, 2] - m[0, 0] - m[1, 1], 0],
            [m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1], m[0, 0] + m[1, 1] + m[2, 2]]]) / 3.0
        vals, vecs = np.linalg.eigh(K)
        q = vecs[[3, 0, 1, 2], np.argmax(vals)]
        if q[0] < 0:
            q = -q
        out[i] = q
    return out.reshape(R.shape[:-2] + (4,))


def axis_angle_to_rotmat(v):
    v = np.asarray(v, dtype=np.float64)
    th = np.linalg.norm(v, axis=-1, keepdims=True)
    k = np.divide(v, th, out=np.zeros_like(v), where=th > 0)
    K = np.zeros(v.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = np.sin(th)[..., None], np.cos(th)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


# ----------------------------------------------------------------------------- HP2

def make_ba_problem(num_images, num_points, track_len=12, seed=0, noise_px=0.5, focal=900.0,
                    cx=512.0, cy=218.0, center_noise=0.02, rot_noise_deg=0.5, point_noise=0.05,
                    track_len_range=None, dynamic_fraction=0.0, shuffle=True):
    """Global-BA stand-in (SURVEY.md §8d, config 5 generator): cameras on a 2-turn helix
    looking at the scene centroid, points uniform in a box in front of every camera, each
    point seen in a contiguous window of `track_len` frames (or U{lo..hi} when
    `track_len_range`), SIMPLE_PINHOLE shared by all images, observations = projection +
    N(0, noise_px^2) rounded to f32 (keypoints are stored f32, colmap_utils/database.py:185).
    Start = truth perturbed; gauge as the reference fixes it (sfm/global_mapper.cc:431-435):
    image 0 pose constant, image 1 tvec[0] constant.
    Returns (BAProblem, truth dict)."""
    rng = np.random.default_rng(seed)
    F, P = int(num_images), int(num_points)
    ang = np.linspace(0.0, 4.0 * np.pi, F)
    centres = np.stack([8.0 * np.cos(ang), 8.0 * np.sin(ang), np.linspace(-2.0, 2.0, F)], -1)
    zax = -centres / np.linalg.norm(centres, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    xax = np.cross(up, zax)
    xax /= np.linalg.norm(xax, axis=1, keepdims=True)
    yax = np.cross(zax, xax)
    R_true = np.stack([xax, yax, zax], 1)                       # world -> camera
    t_true = -np.einsum("fij,fj->fi", R_true, centres)
    q_true = rotmat_to_qvec(R_true)
    X_true = rng.uniform(-2.5, 2.5, size=(P, 3))

    if track_len_range is None:
        lens = np.full(P, min(int(track_len), F), dtype=np.int64)
    else:
        lo, hi = track_len_range
        lens = rng.integers(lo, min(hi, F) + 1, size=P)
    starts = (rng.random(P) * (F - lens + 1)).astype(np.int64)
    M = int(lens.sum())
    obs_point = np.repeat(np.arange(P, dtype=np.int64), lens)
    first = np.cumsum(lens) - lens
    obs_image = np.arange(M, dtype=np.int64) - np.repeat(first, lens) + np.repeat(starts, lens)
    if dynamic_fraction > 0:          # observations flagged dynamic are dropped before BA
        keep = rng.random(M) >= dynamic_fraction
        obs_point, obs_image = obs_point[keep], obs_image[keep]
        M = obs_point.shape[0]
    Xc = np.einsum("mij,mj->mi", R_true[obs_image], X_true[obs_point]) + t_true[obs_image]
    uv = Xc[:, :2] / Xc[:, 2:3]
    xy = focal * uv + np.array([cx, cy])
    xy = (xy + rng.normal(0.0, noise_px, size=xy.shape)).astype(np.float32).astype(np.float64)
    if shuffle:                        # the reference adds residuals image by image, not by point
        perm = rng.permutation(M)
        obs_point, obs_image, xy = obs_point[perm], obs_image[perm], xy[perm]

    # perturbed start
    c0 = centres + rng.normal(0.0, center_noise, size=centres.shape)
    dR = axis_angle_to_rotmat(rng.normal(size=(F, 3)) * np.deg2rad(rot_noise_deg) / np.sqrt(3.0))
    R0 = np.einsum("fij,fjk->fik", dR, R_true)
    R0[0], c0[0] = R_true[0], centres[0]
    q0 = rotmat_to_qvec(R0)
    t0 = -np.einsum("fij,fj->fi", R0, c0)
    X0 = X_true + rng.normal(0.0, point_noise, size=X_true.shape)
    pose_constant = np.zeros(F, np.uint8)
    pose_constant[0] = 1
    tmask = np.zeros(F, np.uint8)
    if F > 1:
        tmask[1] = 1
    prob = BAProblem(q0, t0, X0, np.array([[focal, cx, cy]]), obs_image.astype(np.int32),
                     obs_point.astype(np.int32), xy, np.zeros(F, np.int32), pose_constant, tmask,
                     np.zeros(1, np.uint8))
    truth = dict(qvec=q_true, tvec=t_true, xyz=X_true, centres=centres)
    return prob, truth


def camera_centres(qvec, tvec):
    """Image::ProjectionCenter() = -R(q)^T t."""
    R = qvec_to_rotmat(qvec)
    return -np.einsum("fji,fj->fi", R, np.asarray(tvec))


def umeyama_ate(est_centres, gt_centres):
    """ATE as evaluation_evo/eval_sintel.py:57-107 computes it through evo
    (align=True, correct_scale=True): Sim(3) Umeyama alignment of the estimated camera
    centres to ground truth, RMSE of the translation residuals."""
    x, y = np.asarray(est_centres, float), np.asarray(gt_centres, float)
    mx, my = x.mean(0), y.mean(0)
    xc, yc = x - mx, y - my
    cov = yc.T @ xc / x.shape[0]
    U, D, Vt = np.linalg.svd(cov)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    var_x = (xc ** 2).sum() / x.shape[0]
    s = np.trace(np.diag(D) @ S) / var_x
    t = my - s * R @ mx
    err = y - (s * (R @ x.T).T + t)
    return float(np.sqrt((err ** 2).sum(1).mean()))


# ----------------------------------------------------------------------------- HP1

def bilinear_zeros(img, xy):
    """Bilinear sample of img [H,W,C] at xy [N,2] (x=col, y=row), zeros outside — the
    semantics of torch grid_sample(align_corners=True, padding_mode='zeros') that
    point_trajectory/trajectory.py:25-37 uses (here in float64; the bit-faithful float32
    version lives in tracker.py)."""
    img = np.asarray(img)
    H, W = img.shape[:2]
    x, y = xy[:, 0], xy[:, 1]
    x0, y0 = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
    fx, fy = x - x0, y - y0
    out = np.zeros((xy.shape[0],) + img.shape[2:], dtype=np.float64)
    for dy, wy in ((0, 1 - fy), (1, fy)):
        for dx, wx in ((0, 1 - fx), (1, fx)):
            xi, yi = x0 + dx, y0 + dy
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            v = img[np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)].astype(np.float64)
            wgt = (wx * wy * ok)
            out += v * (wgt[:, None] if v.ndim == 2 else wgt)
    return out


def smooth_flow(height, width, rng, amplitude=6.0, waves=8):
    """Sum of `waves` low-frequency sinusoids, |flow| <= amplitude px (SURVEY.md §8d cfg 2)."""
    yy, xx = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64),
                         indexing="ij")
    flow = np.zeros((height, width, 2))
    for _ in range(waves):
        kx, ky = rng.uniform(-3, 3, 2) * 2 * np.pi / np.array([width, height])
        ph = rng.uniform(0, 2 * np.pi, 2)
        a = rng.uniform(-1, 1, 2) * amplitude / waves
        flow[..., 0] += a[0] * np.sin(kx * xx + ky * yy + ph[0])
        flow[..., 1] += a[1] * np.sin(kx * xx + ky * yy + ph[1])
    return flow


def make_flow_triplet(height, width, seed=0, amplitude=6.0, noise=0.1):
    """flow01, flow12, flow02 (f32 [H,W,2]) with flow02 = flow01 + flow12∘(x+flow01) + noise,
    and an occlusion map for the stride-2 pair (bool [H,W], ~8% set)."""
    rng = np.random.default_rng(seed)
    f01 = smooth_flow(height, width, rng, amplitude)
    f12 = smooth_flow(height, width, rng, amplitude)
    yy, xx = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64),
                         indexing="ij")
    p1 = np.stack([xx + f01[..., 0], yy + f01[..., 1]], -1).reshape(-1, 2)
    f12_at = bilinear_zeros(f12, p1).reshape(height, width, 2)
    f02 = f01 + f12_at + rng.normal(0, noise, size=f01.shape)
    f01n = f01 + rng.normal(0, noise, size=f01.shape)
    f12n = f12 + rng.normal(0, noise, size=f12.shape)
    occ02 = smooth_flow(height, width, rng, 1.0, 4)[..., 0] > 0.45
    return f01n.astype(np.float32), f12n.astype(np.float32), f02.astype(np.float32), occ02


def make_traj_inputs(num, height, width, seed=0, amplitude=6.0, upper_flow=20.0):
    """The argument tuple IncrementalTrajectorySet.optimize_buffer hands to
    particlesfm.optimize_location (point_trajectory/trajectory.py:161-186) for `num`
    trajectories: uv12 [N,4], ref1 [N,2], ref2 [N,2], scale [N,1], flow12 [H,W,2] f32."""
    f01, f12, f02, occ02 = make_flow_triplet(height, width, seed, amplitude)
    rng = np.random.default_rng(seed + 1000003)
    x0 = np.stack([rng.uniform(8, width - 9, num), rng.uniform(8, height - 9, num)], -1)
    fl01 = bilinear_zeros(f01, x0)
    x1 = x0 + fl01
    x2 = x1 + bilinear_zeros(f12, x1)
    fl02 = bilinear_zeros(f02, x0)
    occ = bilinear_zeros(occ02.astype(np.float64)[..., None], x0)
    scale = (1.0 - occ) * (np.linalg.norm(fl02, axis=-1, keepdims=True) < upper_flow)
    # torch grid_sample returns float32; the reference then mixes into float64 arrays
    fl01 = fl01.astype(np.float32).astype(np.float64)
    fl02 = fl02.astype(np.float32).astype(np.float64)
    scale = scale.astype(np.float32).astype(np.float64)
    uv12 = np.concatenate([x1, x2], -1)
    return uv12, x0 + fl01, x0 + fl02, scale, f12
