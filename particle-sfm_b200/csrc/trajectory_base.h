// trajectory_base.h — host containers behind the `particlesfm` Python module.
//
// API parity target: the reference's Trajectory / TrajectorySet
// (point_trajectory/optimize/src/trajectory_base.h:35-76, semantics in
// trajectory_base.cpp:21-185, Python surface in bindings.cc:33-75).  Written from the
// behaviour, not from the source: locations are plain {x, y} pairs (no Eigen), labels
// are bytes, and the inverted index is a flat hash map.
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <map>
#include <random>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace psfm {

struct Vec2 {
  double x = 0.0, y = 0.0;
};

class Trajectory {
 public:
  Trajectory() = default;
  Trajectory(int time, Vec2 xy, int buffer_size) : buffer_size_(buffer_size) { extend(time, xy); }
  Trajectory(std::vector<int> times_in, std::vector<Vec2> xys_in, std::vector<bool> labels_in)
      : times(std::move(times_in)), xys(std::move(xys_in)) {
    if (labels_in.empty()) labels.assign(times.size(), false);
    else labels = std::move(labels_in);
  }

  // A new observation: enters the FIFO buffer (when one is configured) and pushes the
  // oldest buffered location into the settled list once the buffer is over capacity.
  void extend(int time, Vec2 xy) {
    times.push_back(time);
    labels.push_back(false);
    if (buffer_size_ <= 0) {
      xys.push_back(xy);
      return;
    }
    buffer_xys.push_back(xy);
    if ((int)buffer_xys.size() > buffer_size_) {
      xys.push_back(buffer_xys.front());
      buffer_xys.pop_front();
    }
  }
  void clear_buffer() {
    xys.insert(xys.end(), buffer_xys.begin(), buffer_xys.end());
    buffer_xys.clear();
  }
  void set_buffer_xy(int index, Vec2 xy) {
    if (index < 0 || index >= (int)buffer_xys.size())
      throw std::runtime_error("Error! Index out of bound for the buffer.");
    buffer_xys[(size_t)index] = xy;
  }
  void set_label(int index, bool label) { labels.at((size_t)index) = label; }
  void set_labels(const std::vector<bool>& l) { labels = l; }
  int length() const { return (int)(xys.size() + buffer_xys.size()); }
  Vec2 get_tail_location() const {
    if (length() == 0) throw std::runtime_error("Error! The trajectory is empty!");
    return buffer_xys.empty() ? xys.back() : buffer_xys.back();
  }
  int buffer_size() const { return buffer_size_; }

  std::vector<int> times;
  std::vector<bool> labels;
  std::vector<Vec2> xys;
  std::deque<Vec2> buffer_xys;

 private:
  int buffer_size_ = 0;
};

struct WindowSample {
  std::vector<int> traj_ids;
  int K = 0, L = 0;
  std::vector<double> loc_x, loc_y;   // [K*L] row-major
  std::vector<int> masks;             // [K*L]
};

class TrajectorySet {
 public:
  TrajectorySet() = default;
  explicit TrajectorySet(std::map<int, Trajectory> t) : trajs(std::move(t)) {}

  void insert(int traj_id, Trajectory traj) {
    if (!trajs.emplace(traj_id, std::move(traj)).second)
      throw std::runtime_error("Error! The trajectory id already exists!");
  }

  // frame -> (trajectory id -> position of that frame inside the trajectory)
  void build_invert_indexes() {
    for (const auto& kv : trajs) {
      const Trajectory& t = kv.second;
      const size_t n = std::min(t.times.size(), (size_t)t.length());
      for (size_t i = 0; i < n; ++i) frame_index_[t.times[i]].emplace(kv.first, i);
    }
  }

  // Trajectories seen in >= min_length of the given frames, as padded [K, L] arrays.
  WindowSample sample_inside_window(const std::vector<int>& frame_ids, int min_length, int max_num_tracks) const {
    if (frame_index_.empty()) throw std::runtime_error("Error! The inverted index maps have not been built!");
    std::map<int, int> hits;
    for (int f : frame_ids) {
      auto it = frame_index_.find(f);
      if (it == frame_index_.end()) continue;
      for (const auto& e : it->second) hits[e.first]++;
    }
    WindowSample out;
    for (const auto& h : hits)
      if (h.second >= min_length) out.traj_ids.push_back(h.first);
    if ((int)out.traj_ids.size() > max_num_tracks) {
      std::mt19937 rng(5489u);
      std::shuffle(out.traj_ids.begin(), out.traj_ids.end(), rng);
      out.traj_ids.resize((size_t)max_num_tracks);
    }
    out.K = (int)out.traj_ids.size();
    out.L = (int)frame_ids.size();
    out.loc_x.assign((size_t)out.K * out.L, 0.0);
    out.loc_y.assign((size_t)out.K * out.L, 0.0);
    out.masks.assign((size_t)out.K * out.L, 0);
    for (int j = 0; j < out.L; ++j) {
      auto it = frame_index_.find(frame_ids[(size_t)j]);
      if (it == frame_index_.end()) continue;
      for (int i = 0; i < out.K; ++i) {
        auto e = it->second.find(out.traj_ids[(size_t)i]);
        if (e == it->second.end()) continue;
        const Vec2& p = trajs.at(out.traj_ids[(size_t)i]).xys.at(e->second);
        out.masks[(size_t)i * out.L + j] = 1;
        out.loc_x[(size_t)i * out.L + j] = p.x;
        out.loc_y[(size_t)i * out.L + j] = p.y;
      }
    }
    return out;
  }

  std::map<int, Trajectory> trajs;

 private:
  std::unordered_map<int, std::map<int, size_t>> frame_index_;
};

}  // namespace psfm
