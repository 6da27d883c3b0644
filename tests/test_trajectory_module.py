"""The host-side classes of the `particlesfm` module (csrc/trajectory_base.h, csrc/bindings.cc) against the behaviour of
the reference's point_trajectory/optimize/src/trajectory_base.cpp: buffer semantics (:52-74), the four RuntimeError
messages (:78, :88, :111, :130), keyword-only buffer_size / labels (bindings.cc:34-36 of the reference), as_dict / dict constructors (:37-50, :96-108), pickling (bindings.cc:44-75 of the
reference), the inverted index and sample_inside_window (:113-186).  No GPU needed: only optimize_location touches
the CUDA library."""
import os
import pickle
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "particle-sfm_b200"))
particlesfm = pytest.importorskip("particlesfm")


class _RefTrajectory:
    """Python restatement of the reference class (trajectory_base.cpp:21-93), the checker of this file."""

    def __init__(self, time, xy, buffer_size):
        self.times, self.xys, self.labels, self.buffer, self.bs = [], [], [], [], buffer_size
        self.extend(time, xy)

    def extend(self, time, xy):
        self.times.append(time); self.labels.append(False)
        if self.bs == 0:
            self.xys.append(list(xy)); return
        self.buffer.append(list(xy))
        if len(self.buffer) > self.bs:
            self.xys.append(self.buffer.pop(0))

    def clear_buffer(self):
        self.xys += self.buffer; self.buffer = []

    def length(self):
        return len(self.xys) + len(self.buffer)

    def tail(self):
        return self.buffer[-1] if self.buffer else self.xys[-1]


@pytest.mark.parametrize("buffer_size", [0, 1, 3])
def test_buffer_semantics_follow_the_reference(buffer_size):
    rng = np.random.default_rng(buffer_size)
    t = particlesfm.Trajectory(5, np.array([1.5, 2.5]), buffer_size=buffer_size)
    r = _RefTrajectory(5, [1.5, 2.5], buffer_size)
    for k in range(9):
        xy = rng.uniform(0, 100, 2)
        t.extend(6 + k, xy); r.extend(6 + k, xy)
        assert t.length() == r.length()
        assert np.allclose(t.get_tail_location(), r.tail(), rtol=0, atol=0)
        if buffer_size and k == 4:
            new = np.array([7.0, 8.0])
            t.set_buffer_xy(buffer_size - 1, new); r.buffer[buffer_size - 1] = list(new)
    d = t.as_dict()
    assert d["frame_ids"] == r.times and d["labels"] == r.labels
    assert np.array_equal(np.array(d["locations"]).reshape(-1, 2), np.array(r.xys).reshape(-1, 2))     # the buffer is not part of locations
    t.clear_buffer(); r.clear_buffer()
    assert np.array_equal(np.array(t.as_dict()["locations"]), np.array(r.xys))
    assert t.length() == len(r.times)


def test_error_messages_are_the_reference_ones():
    t = particlesfm.Trajectory(0, np.array([0.0, 0.0]), buffer_size=2)
    with pytest.raises(RuntimeError, match="Error! Index out of bound for the buffer."):
        t.set_buffer_xy(1, np.array([1.0, 1.0]))                  # one element buffered: index 1 is out of bound (:78)
    empty = particlesfm.Trajectory([], [])
    with pytest.raises(RuntimeError, match="Error! The trajectory is empty!"):
        empty.get_tail_location()                                   # :88
    s = particlesfm.TrajectorySet()
    s.insert(3, t)
    with pytest.raises(RuntimeError, match="Error! The trajectory id already exists!"):
        s.insert(3, t)                                              # :111
    with pytest.raises(RuntimeError, match="Error! The inverted index maps have not been built!"):
        s.sample_inside_window([0, 1], 1, 10)                       # :130


def test_dict_and_pickle_round_trips():
    t = particlesfm.Trajectory([3, 4, 5], [np.array([1.0, 2.0]), np.array([3.0, 4.0]), np.array([5.0, 6.0])], labels=[False, True, False])
    d = t.as_dict()
    assert d["frame_ids"] == [3, 4, 5] and d["labels"] == [False, True, False]
    t2 = particlesfm.Trajectory(d)
    assert t2.as_dict()["frame_ids"] == d["frame_ids"] and np.array_equal(np.array(t2.as_dict()["locations"]), np.array(d["locations"]))
    t3 = pickle.loads(pickle.dumps(t))
    assert t3.as_dict()["labels"] == d["labels"] and np.array_equal(np.array(t3.as_dict()["locations"]), np.array(d["locations"]))
    # labels default to all-false (:29-33)
    assert particlesfm.Trajectory([1, 2], [np.zeros(2), np.ones(2)]).as_dict()["labels"] == [False, False]
    s = particlesfm.TrajectorySet({7: t, 9: t2})
    sd = s.as_dict()
    assert sorted(sd) == [7, 9] and sd[7]["frame_ids"] == [3, 4, 5]
    s2 = pickle.loads(pickle.dumps(s))
    assert sorted(s2.as_dict()) == [7, 9] and s2.as_dict()[9]["labels"] == d["labels"]
    s3 = particlesfm.TrajectorySet(sd)                              # dict-of-dicts constructor (:103-108)
    assert s3.as_dict()[7]["frame_ids"] == [3, 4, 5]


def test_sample_inside_window_masks_and_locations():
    s = particlesfm.TrajectorySet()
    s.insert(0, particlesfm.Trajectory([0, 1, 2, 3], [np.array([k, 10.0 + k]) for k in range(4)]))
    s.insert(1, particlesfm.Trajectory([2, 3], [np.array([20.0, 21.0]), np.array([22.0, 23.0])]))
    s.insert(2, particlesfm.Trajectory([9], [np.array([5.0, 5.0])]))
    s.build_invert_indexes()
    out = s.sample_inside_window([1, 2, 3, 7], 2, 100)              # frame 7 has no trajectory: mask false (:167-168)
    ids = list(out["traj_ids"])
    assert sorted(ids) == [0, 1]                                    # trajectory 2 is seen in < min_length frames of the window
    masks = np.asarray(out["masks"])
    xs, ys = (np.asarray(a) for a in out["locations"])
    i0, i1 = ids.index(0), ids.index(1)
    assert masks[i0].tolist() == [1, 1, 1, 0] and masks[i1].tolist() == [0, 1, 1, 0]
    assert xs[i0].tolist() == [1.0, 2.0, 3.0, 0.0] and ys[i0].tolist() == [11.0, 12.0, 13.0, 0.0]
    assert xs[i1].tolist() == [0.0, 20.0, 22.0, 0.0] and ys[i1].tolist() == [0.0, 21.0, 23.0, 0.0]
    few = s.sample_inside_window([0, 1, 2, 3], 1, 1)                # more candidates than max_num_tracks: shuffled, cut (:150-153)
    assert len(few["traj_ids"]) == 1 and np.asarray(few["masks"]).shape == (1, 4)
