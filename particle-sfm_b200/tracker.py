"""Batched (SoA) mirror of the reference's sequential point tracker around HP1.

Reference:
  track_optimize                      point_trajectory/track_optimize.py:24-54
  IncrementalTrajectorySet            point_trajectory/trajectory.py:98-194
      new_traj_all :117, get_cur_pos :122, extend_all :129, clear_active :154,
      optimize_buffer :161-194
  step_forward                        point_trajectory/trajectory.py:45-62
  grid_sample                         point_trajectory/trajectory.py:25-37
  flow_check / get_occ_mask           point_trajectory/utils.py:85-105

The reference keeps one Python/pybind `Trajectory` object per particle and walks them in
Python loops (≈5 pybind calls per trajectory per frame).  Here the live particles are
struct-of-arrays: per time step one [n, 2] array of locations plus index arrays that link
a particle to its slot in the two previous time steps (the 3-frame FIFO buffer of the
reference is exactly "the last three time steps").  The per-frame work is vectorised and
the optimiser is called once per frame on the whole batch, like the reference does.

Semantics that decide the INTEGER TRACK CONNECTIVITY are kept bit-for-bit:
  * flows / occlusion maps are sampled with torch.nn.functional.grid_sample on the CPU in
    float32 with the reference's normalisation (x /= (W-1)/2, -= 1; align_corners=True),
  * a particle survives iff 0 < x < W-1, 0 < y < H-1 and the sampled occlusion <= 0.1,
  * re-seeding: occupancy at (int(y), int(x)), Euclidean distance transform > ratio on the
    ratio-strided grid,
  * trajectory ids are the positions in the reference's `full_trajs` list (retire order).
"""
import ctypes as _C

import numpy as np


# ----------------------------------------------------------------------------- device ops (csrc/tracker.cu)

def _f32p(a):
    return a.ctypes.data_as(_C.POINTER(_C.c_float))


def _u8p(a):
    return a.ctypes.data_as(_C.POINTER(_C.c_uint8))


def grid_sample_device(map_hwc, xy):
    """grid_sample (trajectory.py:25-37) on the GPU, bit for bit what torch's CPU kernel returns.
    map_hwc: [H, W, C] float32 (C = 1 or 2) or [H, W]; xy [N, 2] float64 -> [N, C] float32."""
    from . import _lib
    m = np.ascontiguousarray(map_hwc, np.float32)
    if m.ndim == 2:
        m = m[:, :, None]
    h, w, c = m.shape
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    out = np.empty((xy.shape[0], c), np.float32)
    _lib.check(_lib.lib().psfm_grid_sample(_f32p(m), h, w, c, _lib.dptr(xy), xy.shape[0], _f32p(out)), "psfm_grid_sample")
    return out


def flow_check_device(flows, flows_b, thres):
    """flow_check (point_trajectory/utils.py:58-105) on the GPU -> (error_maps, occ_maps)."""
    from . import _lib
    error_maps, occ_maps = [], []
    for f, f_b in zip(flows, flows_b):
        f = np.ascontiguousarray(f, np.float32)
        f_b = np.ascontiguousarray(f_b, np.float32)
        h, w = f.shape[:2]
        err = np.empty((h, w), np.float32)
        occ = np.empty((h, w), np.uint8)
        _lib.check(_lib.lib().psfm_flow_check(_f32p(f), _f32p(f_b), h, w, float(thres), _f32p(err), _u8p(occ)), "psfm_flow_check")
        error_maps.append(err)
        occ_maps.append(occ.astype(bool))
    return error_maps, occ_maps


def tracker_step_device(flow, occ, cur_xy, sample_ratio):
    """step_forward + the re-seeding mask of extend_all -> (next_xy, flags, reseed_mask or None)."""
    from . import _lib
    flow = np.ascontiguousarray(flow, np.float32)
    occ8 = np.ascontiguousarray(occ, np.uint8)
    h, w = flow.shape[:2]
    cur = np.ascontiguousarray(cur_xy, np.float64).reshape(-1, 2)
    n = cur.shape[0]
    nxt = np.empty((n, 2), np.float64)
    flags = np.zeros(n, np.uint8)
    gh, gw = -(-h // sample_ratio), -(-w // sample_ratio)
    mask = np.zeros((gh, gw), np.uint8)
    _lib.check(_lib.lib().psfm_tracker_step(_f32p(flow), _u8p(occ8), h, w, _lib.dptr(cur), n, sample_ratio, _lib.dptr(nxt),
                                            _u8p(flags), _u8p(mask)), "psfm_tracker_step")
    return nxt, flags, (mask.astype(bool) if flags.any() else None)


def buffer_inputs_device(flow01, flow02, occ02, x0, upper_flow=20.0):
    """optimize_buffer's ref1, ref2, scale (trajectory.py:171-183) on the GPU."""
    from . import _lib
    f1 = np.ascontiguousarray(flow01, np.float32)
    f2 = np.ascontiguousarray(flow02, np.float32)
    o2 = np.ascontiguousarray(occ02, np.uint8)
    h, w = f1.shape[:2]
    x0 = np.ascontiguousarray(x0, np.float64).reshape(-1, 2)
    n = x0.shape[0]
    ref1, ref2, scale = np.empty((n, 2)), np.empty((n, 2)), np.empty((n, 1))
    _lib.check(_lib.lib().psfm_tracker_buffer_inputs(_f32p(f1), _f32p(f2), _u8p(o2), h, w, _lib.dptr(x0), n, float(upper_flow),
                                                     _lib.dptr(ref1), _lib.dptr(ref2), _lib.dptr(scale)), "psfm_tracker_buffer_inputs")
    return ref1, ref2, scale


def grid_sample(data, xy):
    """data: torch tensor [C, H, W]; xy: [N, 2] float64 → [N, C] float32 numpy.
    Same arithmetic as point_trajectory/trajectory.py:25-37."""
    import torch
    data = data.unsqueeze(0)
    g = torch.from_numpy(np.ascontiguousarray(xy)).float().to(data.device)
    g = g.unsqueeze(0).unsqueeze(0)
    H, W = data.shape[2], data.shape[3]
    g[:, :, :, 0] /= ((W - 1) / 2)
    g[:, :, :, 1] /= ((H - 1) / 2)
    g -= 1
    out = torch.nn.functional.grid_sample(data, g, align_corners=True)
    return out.squeeze(0).squeeze(1).permute(1, 0).cpu().numpy()


def flow_check(flows, flows_b, thres):
    """Forward/backward consistency → (error_maps, occ_maps); point_trajectory/utils.py:58-105."""
    import torch
    import torch.nn.functional as F
    error_maps, occ_maps = [], []
    for f, f_b in zip(flows, flows_b):
        f_t = torch.from_numpy(f).permute(2, 0, 1).unsqueeze(0).float()
        b_t = torch.from_numpy(f_b).permute(2, 0, 1).unsqueeze(0).float()
        B, _, H, W = f_t.shape
        hh, ww = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
        coord = torch.stack([ww, hh], 0).unsqueeze(0)
        grid = coord + f_t
        oob = ((grid[:, 0] < 0) + (grid[:, 0] > W - 1) + (grid[:, 1] < 0) + (grid[:, 1] > H - 1)).float()
        grid = grid.clone()
        grid[:, 0] /= (W - 1) / 2
        grid[:, 1] /= (H - 1) / 2
        grid -= 1
        warp = F.grid_sample(b_t, grid.permute(0, 2, 3, 1), align_corners=True)
        err = torch.norm(warp + f_t, dim=1)
        mask = torch.clamp((err > thres) + oob, 0, 1) > 0
        error_maps.append(err.squeeze().numpy())
        occ_maps.append(mask.squeeze().numpy())
    return error_maps, occ_maps


class BatchedTrajectorySet:
    """SoA replacement of IncrementalTrajectorySet (buffer_size = 3)."""

    def __init__(self, total_length, img_h, img_w, sample_ratio, optimize_fn, device=False):
        self.total_length, self.h, self.w, self.ratio = total_length, img_h, img_w, sample_ratio
        self.optimize_fn = optimize_fn
        self.device = device            # True: sampling / survival / re-seeding on the GPU (csrc/tracker.cu)
        x, y = np.arange(0, img_w), np.arange(0, img_h)
        xx, yy = np.meshgrid(x, y)
        self.all_candidates = np.stack([xx, yy], -1)[::sample_ratio, ::sample_ratio, :]
        self.sample_candidates = np.reshape(np.copy(self.all_candidates), (-1, 2))
        # per time step: ids [n] and locations [n, 2] of the particles alive at that time
        self.ids_at, self.xy_at = {}, {}
        # live particles (in the reference's active_trajs order)
        self.act_id = np.zeros(0, np.int64)
        self.act_len = np.zeros(0, np.int64)            # observations so far
        self.act_i0 = np.zeros(0, np.int64)             # slot in the current time step's arrays
        self.act_im1 = np.zeros(0, np.int64)            # slot one step back (-1: none)
        self.act_im2 = np.zeros(0, np.int64)            # slot two steps back (-1: none)
        self.next_id = 0
        self.cur_time = None
        self.retired = []                               # arrays of ids in retire order
        self.start_time = {}

    # -- new_traj_all (trajectory.py:117-120)
    def new_traj_all(self, time, points):
        n = points.shape[0]
        ids = np.arange(self.next_id, self.next_id + n, dtype=np.int64)
        self.next_id += n
        pts = points.astype(np.float64)
        if time in self.ids_at:
            base = self.ids_at[time].shape[0]
            self.ids_at[time] = np.concatenate([self.ids_at[time], ids])
            self.xy_at[time] = np.concatenate([self.xy_at[time], pts])
        else:
            base = 0
            self.ids_at[time], self.xy_at[time] = ids, pts
        self.act_id = np.concatenate([self.act_id, ids])
        self.act_len = np.concatenate([self.act_len, np.ones(n, np.int64)])
        self.act_i0 = np.concatenate([self.act_i0, base + np.arange(n, dtype=np.int64)])
        self.act_im1 = np.concatenate([self.act_im1, np.full(n, -1, np.int64)])
        self.act_im2 = np.concatenate([self.act_im2, np.full(n, -1, np.int64)])
        self.cur_time = time

    # -- get_cur_pos (trajectory.py:122-127)
    def get_cur_pos(self):
        return self.xy_at[self.cur_time][self.act_i0]

    # -- extend_all (trajectory.py:129-152)
    def extend_all(self, next_xys, next_time, flags, reseed_mask=None):
        assert len(next_xys) == self.act_id.shape[0] == len(flags)
        keep = np.asarray(flags) != 0
        self.retired.append(self.act_id[~keep])
        nx = next_xys[keep]
        n = int(keep.sum())
        self.ids_at[next_time], self.xy_at[next_time] = self.act_id[keep], nx.astype(np.float64)
        self.act_id = self.act_id[keep]
        self.act_len = self.act_len[keep] + 1
        self.act_im2 = self.act_im1[keep]
        self.act_im1 = self.act_i0[keep]
        self.act_i0 = np.arange(n, dtype=np.int64)
        self.cur_time = next_time
        if reseed_mask is not None:          # computed on the device with the step (exact: squared integer distances)
            sample_map = reseed_mask
        else:
            import scipy.ndimage
            occupied = np.zeros((self.h, self.w, 1))
            occupied[nx[:, 1].astype(np.int64), nx[:, 0].astype(np.int64)] = 1      # int() truncation
            dist = scipy.ndimage.distance_transform_edt(1.0 - occupied)
            sample_map = (dist > self.ratio)[::self.ratio, ::self.ratio, 0]
        self.sample_candidates = np.copy(self.all_candidates[sample_map])

    def clear_active(self):
        self.retired.append(self.act_id)
        self.act_id = np.zeros(0, np.int64)

    # -- optimize_buffer (trajectory.py:161-194)
    def optimize_buffer(self, flow01_map, flow12_map, flow02_map, occ02_map, next_time, upper_flow=20.0):
        import torch
        sel = np.nonzero(self.act_len >= 3)[0]          # len(buffer_xys) == 3
        t2, t1, t0 = next_time, next_time - 1, next_time - 2
        i2, i1, i0 = self.act_i0[sel], self.act_im1[sel], self.act_im2[sel]
        x0 = self.xy_at[t0][i0]
        x1 = self.xy_at[t1][i1]
        x2 = self.xy_at[t2][i2]
        if sel.shape[0] == 0:
            raise ValueError("need at least one array to stack")     # np.stack([]) in the reference
        h, w = flow01_map.shape[0], flow01_map.shape[1]
        if self.device:
            ref1, ref2, scale = buffer_inputs_device(flow01_map, flow02_map, occ02_map, x0, upper_flow)
        else:
            flow01 = grid_sample(torch.from_numpy(flow01_map).permute(2, 0, 1).float(), x0)
            flow02 = grid_sample(torch.from_numpy(flow02_map).permute(2, 0, 1).float(), x0)
            occ02 = grid_sample(torch.from_numpy(occ02_map).unsqueeze(0).float(), x0)
            scale = (1.0 - occ02) * (np.linalg.norm(flow02, axis=-1, keepdims=True) < upper_flow)
            ref1 = x0 + flow01
            ref2 = x0 + flow02
        uv12 = np.concatenate([x1, x2], axis=1)
        new = self.optimize_fn(uv12, ref1, ref2, scale, flow12_map, uv12.shape[0], w, h)
        new = np.asarray(new, np.float64).reshape(-1, 2, 2)
        self.xy_at[t1][i1] = new[:, 0]
        self.xy_at[t2][i2] = new[:, 1]

    # -- results, in the reference's `full_trajs` order
    def full_trajs(self, traj_min_len=0):
        order = np.concatenate(self.retired) if self.retired else np.zeros(0, np.int64)
        rank = np.empty(self.next_id, np.int64)
        rank[order] = np.arange(order.shape[0])
        times = sorted(self.ids_at)
        ids = np.concatenate([self.ids_at[t] for t in times])
        tt = np.concatenate([np.full(self.ids_at[t].shape[0], t, np.int64) for t in times])
        xy = np.concatenate([self.xy_at[t] for t in times])
        key = np.lexsort((tt, rank[ids]))
        ids, tt, xy = rank[ids][key], tt[key], xy[key]
        bounds = np.flatnonzero(np.diff(ids)) + 1
        starts = np.concatenate([[0], bounds])
        ends = np.concatenate([bounds, [ids.shape[0]]])
        out = {}
        for s, e in zip(starts, ends):
            if e - s >= traj_min_len:
                out[int(ids[s])] = {"frame_ids": tt[s:e].tolist(), "locations": [xy[k].copy() for k in range(s, e)],
                                    "labels": [False] * int(e - s)}
        return out


def _default_optimize():
    from . import traj
    return traj.optimize_location


def track_optimize(flows, flows_f2, occ_maps, occ_maps_s2, sample_ratio, optimize_fn=None, traj_min_len=0, device=False):
    """Sequentially track and optimise point trajectories (track_optimize.py:24-54).
    Returns {traj_id: {"frame_ids", "locations", "labels"}} with the reference's ids
    (= positions in its `full_trajs` list); `traj_min_len` applies the filter of
    main_connect_point_trajectories.py:57-60.  device=True: the float32 sampling, the survival test and
    the re-seeding run on the GPU (same bits, csrc/tracker.cu) instead of torch-CPU / scipy."""
    import torch
    optimize_fn = optimize_fn or _default_optimize()
    n_flows = len(flows)
    h, w = flows[0].shape[:2]
    trajs = BatchedTrajectorySet(n_flows + 1, h, w, sample_ratio, optimize_fn, device=device)
    for frame_id in range(n_flows):
        trajs.new_traj_all(frame_id, trajs.sample_candidates)
        cur_xys = trajs.get_cur_pos()
        if device:
            next_xys, flags, mask = tracker_step_device(flows[frame_id], occ_maps[frame_id], cur_xys, sample_ratio)
            trajs.extend_all(next_xys, frame_id + 1, flags, mask)
            if frame_id + 1 >= 2:
                trajs.optimize_buffer(flows[frame_id - 1], flows[frame_id], flows_f2[frame_id - 1],
                                      occ_maps_s2[frame_id - 1], frame_id + 1)
            continue
        flow_sample = grid_sample(torch.from_numpy(flows[frame_id]).permute(2, 0, 1).float(), cur_xys)
        # step_forward (trajectory.py:45-62)
        occ = grid_sample(torch.from_numpy(occ_maps[frame_id]).unsqueeze(0).float(), cur_xys) > 0.1
        next_xys = cur_xys + flow_sample
        valid = (next_xys[:, 0] > 0) * (next_xys[:, 0] < w - 1) * (next_xys[:, 1] > 0) * (next_xys[:, 1] < h - 1)
        flags = valid * (1.0 - np.squeeze(occ, axis=-1))
        trajs.extend_all(next_xys, frame_id + 1, flags)
        if frame_id + 1 >= 2:
            trajs.optimize_buffer(flows[frame_id - 1], flows[frame_id], flows_f2[frame_id - 1],
                                  occ_maps_s2[frame_id - 1], frame_id + 1)
    trajs.clear_active()
    return trajs.full_trajs(traj_min_len)


def main_connect_point_trajectories(flows_f, flows_b, flows_f2, flows_b2, sample_ratio=2, flow_check_thres=1.0,
                                    traj_min_len=3, optimize_fn=None, device=False):
    """In-memory equivalent of main_connect_point_trajectories.py:27-62 with
    skip_path_consistency=False: returns the dict a `particlesfm.TrajectorySet` is built
    from (and np.save'd as track.npy)."""
    fc = flow_check_device if device else flow_check
    _, occ = fc(flows_f, flows_b, flow_check_thres)
    _, occ2 = fc(flows_f2, flows_b2, flow_check_thres)
    return track_optimize(flows_f, flows_f2, occ, occ2, sample_ratio, optimize_fn, traj_min_len, device=device)
