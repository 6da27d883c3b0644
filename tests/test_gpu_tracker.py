"""f-1: the tracker stage around HP1 on the device (csrc/tracker.cu) — BIT FOR BIT against torch's CPU
grid_sample / scipy's distance transform (what the reference's Python tracker runs:
point_trajectory/trajectory.py:25-62,117-194, utils.py:58-105) and against the golden track set made from
the reference (tests/golden/tracker_small.npz)."""
import os
import sys

import numpy as np
import pytest

from particlesfm_b200 import tracker

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_tracker import _load, _compare   # noqa: E402


@pytest.mark.parametrize("h,w", [(37, 53), (436, 1024), (480, 854)])
def test_grid_sample_bits(gpu, h, w):
    import torch
    rng = np.random.default_rng(h)
    data = (rng.normal(size=(h, w, 2)) * 5).astype(np.float32)
    xy = np.stack([rng.uniform(-4, w + 3, 60000), rng.uniform(-4, h + 3, 60000)], 1)
    xy[:100] = np.round(xy[:100])                      # exactly on pixel centres
    xy[100:200, 0] = w - 1                             # on the last column / row
    xy[200:300, 1] = h - 1
    ref = tracker.grid_sample(torch.from_numpy(data).permute(2, 0, 1).float(), xy)
    out = tracker.grid_sample_device(data, xy)
    assert np.array_equal(out, ref)
    occ = rng.random((h, w)) > 0.7
    ref1 = tracker.grid_sample(torch.from_numpy(occ).unsqueeze(0).float(), xy)
    out1 = tracker.grid_sample_device(occ.astype(np.float32), xy)
    assert np.array_equal(out1, ref1)


def test_flow_check_bits(gpu):
    g, fw, fb, f2, b2, occ, occ2 = _load()
    e_ref, o_ref = tracker.flow_check(fw, fb, 1.0)
    e_dev, o_dev = tracker.flow_check_device(fw, fb, 1.0)
    for a, b, c, d, e in zip(e_ref, e_dev, o_ref, o_dev, occ):
        assert np.array_equal(a, b) and np.array_equal(c, d) and np.array_equal(d, e)
    rng = np.random.default_rng(3)
    f = (rng.normal(size=(218, 512, 2)) * 4).astype(np.float32)
    b = (-f + rng.normal(size=f.shape) * 0.8).astype(np.float32)
    e_ref, o_ref = tracker.flow_check([f], [b], 1.0)
    e_dev, o_dev = tracker.flow_check_device([f], [b], 1.0)
    assert np.array_equal(e_ref[0], e_dev[0]) and np.array_equal(o_ref[0], o_dev[0])


@pytest.mark.parametrize("ratio", [1, 2, 3])
def test_step_and_reseed_bits(gpu, ratio):
    import scipy.ndimage
    import torch
    rng = np.random.default_rng(10 + ratio)
    h, w, n = 120, 200, 9000
    flow = (rng.normal(size=(h, w, 2)) * 3).astype(np.float32)
    occ = rng.random((h, w)) > 0.85
    cur = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], 1)
    nxt, flags, mask = tracker.tracker_step_device(flow, occ, cur, ratio)
    fs = tracker.grid_sample(torch.from_numpy(flow).permute(2, 0, 1).float(), cur)
    oc = tracker.grid_sample(torch.from_numpy(occ).unsqueeze(0).float(), cur) > 0.1
    ref_next = cur + fs
    valid = (ref_next[:, 0] > 0) * (ref_next[:, 0] < w - 1) * (ref_next[:, 1] > 0) * (ref_next[:, 1] < h - 1)
    ref_flags = valid * (1.0 - np.squeeze(oc, axis=-1))
    assert np.array_equal(nxt, ref_next) and np.array_equal(flags != 0, ref_flags != 0)
    keep = ref_flags != 0
    occupied = np.zeros((h, w, 1))
    occupied[ref_next[keep, 1].astype(np.int64), ref_next[keep, 0].astype(np.int64)] = 1
    dist = scipy.ndimage.distance_transform_edt(1.0 - occupied)
    assert np.array_equal(mask, (dist > ratio)[::ratio, ::ratio, 0])


def test_buffer_inputs_bits(gpu):
    import torch
    rng = np.random.default_rng(21)
    h, w, n = 90, 160, 7000
    f01 = (rng.normal(size=(h, w, 2)) * 3).astype(np.float32)
    f02 = (rng.normal(size=(h, w, 2)) * 12).astype(np.float32)       # some |flow02| above the 20 px bound
    occ = rng.random((h, w)) > 0.8
    x0 = np.stack([rng.uniform(1, w - 2, n), rng.uniform(1, h - 2, n)], 1)
    ref1, ref2, scale = tracker.buffer_inputs_device(f01, f02, occ, x0)
    a = tracker.grid_sample(torch.from_numpy(f01).permute(2, 0, 1).float(), x0)
    b = tracker.grid_sample(torch.from_numpy(f02).permute(2, 0, 1).float(), x0)
    o = tracker.grid_sample(torch.from_numpy(occ).unsqueeze(0).float(), x0)
    sc = (1.0 - o) * (np.linalg.norm(b, axis=-1, keepdims=True) < 20.0)
    assert np.array_equal(ref1, x0 + a) and np.array_equal(ref2, x0 + b)
    assert np.array_equal(scale, sc.astype(np.float64)) and 0 < (scale == 0).sum() < n


def test_device_tracker_reproduces_the_reference_track_set(gpu):
    """The whole stage — flow check, sampling, survival, re-seeding on the GPU, HP1 through the pybind11
    module — gives the reference's integer connectivity and the same doubles."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "particle-sfm_b200"))
    import particlesfm
    g, fw, fb, f2, b2, occ, occ2 = _load()
    res = tracker.track_optimize(fw, f2, occ, occ2, 2, optimize_fn=particlesfm.optimize_location, device=True)
    _compare(res, g)
    res3 = tracker.main_connect_point_trajectories(fw, fb, f2, b2, 2, 1.0, 3, optimize_fn=particlesfm.optimize_location, device=True)
    assert sorted(res3) == [int(i) for i, l in zip(g["ids"], g["lens"]) if l >= 3]
