// TEST INFRASTRUCTURE — CPU oracle (see orc_math.h header).
//
// Branch-and-bound loop-closure matcher in the form this fork actually calls:
// FastCorrelativeScanMatcher3D::MatchWith3DofInitial — a single discretised scan at the given pose, search over the
// (x, y, z) translation window only. Restated from
//   SM/precomputation_grid_3d.h:25-36 (8-bit grid, ToProbability), precomputation_grid_3d.cc:27-81
//     (DivideByTwoRoundingTowardsNegativeInfinity, ConvertToPrecomputationGrid, PrecomputeGrid)
//   SM/fast_correlative_scan_matcher_3d.cc:57-77 (PrecomputationGridStack3D), :79-110 (DiscreteScan3D, Candidate3D),
//     :165-196 (MatchWith3DofInitial), :253-295 (DiscretizeScan), :351-384 (GenerateLowestResolutionCandidates),
//     :386-409 (ScoreCandidates), :423-430 (GetPoseFromCandidate), :432-495 (BranchAndBound)
//   SM/low_resolution_matcher.cc:24-36
// The 8-bit precomputation grids reuse the HybridGrid container (same three-level geometry; HybridGridBase<uint8>
// in the reference) and simply store values 0..255.
#pragma once
#include <algorithm>
#include <functional>
#include <limits>
#include <memory>
#include <vector>

#include "orc_grid.h"
#include "orc_math.h"
#include "orc_rotational.h"

namespace orc {

struct FcsmOptions {
  int branch_and_bound_depth = 8;
  int full_resolution_depth = 3;
  double min_rotational_score = 0.77;
  double min_low_resolution_score = 0.55;
  double linear_xy_search_window = 5.;
  double linear_z_search_window = 1.;
  double angular_search_window = 0.2617993877991494;  // unused by MatchWith3DofInitial
};

inline float precomp_to_probability(float value) { return kMinProbability + value * ((kMaxProbability - kMinProbability) / 255.f); }
inline int divide_by_two_rounding_down(int v) { return v >= 0 ? v / 2 : (v - 1) / 2; }

inline std::unique_ptr<HybridGrid> convert_to_precomputation_grid(const HybridGrid& grid) {
  auto result = std::make_unique<HybridGrid>(grid.resolution());
  grid.ForEachCell([&](const I3& c, uint16_t v) {
    const int cell_value =
        round_to_int((value_to_probability(v) - kMinProbability) * (255.f / (kMaxProbability - kMinProbability)));
    *result->mutable_value(c) = (uint16_t)cell_value;
  });
  return result;
}

inline std::unique_ptr<HybridGrid> precompute_grid(const HybridGrid& grid, bool half_resolution, const I3& shift) {
  auto result = std::make_unique<HybridGrid>(grid.resolution());
  grid.ForEachCell([&](const I3& c, uint16_t v) {
    for (int i = 0; i != 8; ++i) {
      I3 ci{c.x - shift.x * (i & 1), c.y - shift.y * ((i >> 1) & 1), c.z - shift.z * ((i >> 2) & 1)};
      if (half_resolution) ci = {divide_by_two_rounding_down(ci.x), divide_by_two_rounding_down(ci.y), divide_by_two_rounding_down(ci.z)};
      uint16_t* cell = result->mutable_value(ci);
      *cell = std::max(v, *cell);
    }
  });
  return result;
}

class PrecomputationGridStack {
 public:
  PrecomputationGridStack(const HybridGrid& grid, const FcsmOptions& o) {
    grids_.push_back(convert_to_precomputation_grid(grid));
    int last_width = 1;
    for (int depth = 1; depth != o.branch_and_bound_depth; ++depth) {
      const bool half_resolution = depth >= o.full_resolution_depth;
      const int next_width = 1 << depth;
      const int full_voxels = 1 << std::max(0, depth - o.full_resolution_depth);
      const int shift = (next_width - last_width + (full_voxels - 1)) / full_voxels;
      grids_.push_back(precompute_grid(*grids_.back(), half_resolution, I3{shift, shift, shift}));
      last_width = next_width;
    }
  }
  const HybridGrid& Get(int depth) const { return *grids_.at(depth); }
  int max_depth() const { return (int)grids_.size() - 1; }

 private:
  std::vector<std::unique_ptr<HybridGrid>> grids_;
};

struct FcsmResult {
  bool found = false;
  float score = 0.f;
  Rigid3d pose;
  float rotational_score = 0.f;
  float low_resolution_score = 0.f;
  I3 offset{0, 0, 0};
  int64_t leaves_scored = 0;  // work the branch and bound actually did (for the comparison with brute force)
  int scan_index = 0, num_scans = 1;  // which yaw step won, of how many (full Match only)
};

// Eigen::Quaternion::inverse(): conjugate divided by the squared norm (not just the conjugate)
template <typename T>
inline Quat<T> eigen_inverse(const Quat<T>& q) {
  const T n2 = qsquared_norm(q);
  return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
}

class FastCorrelativeScanMatcher {
 public:
  // Builds the precomputation stack once per submap, like the reference's constructor (cc:129-141); Match* is const and
  // keeps its per-call state in a local Search record, so one matcher serves concurrent callers (the reference's
  // thread-pool pattern, constraint_builder_3d.cc:189-197). `histograms_at_angles` feeds the rotational matcher
  // (HistogramsAtAnglesFromNodes, cc:114-127); the 3-DoF entry point does not use it.
  FastCorrelativeScanMatcher(const HybridGrid& hi, const HybridGrid* lo, const FcsmOptions& o,
                             const std::vector<std::pair<Histogram, float>>& histograms_at_angles = zero_histogram())
      : o_(o), resolution_(hi.resolution()), stack_(hi, o), lo_(lo), rotational_(histograms_at_angles) {}

  static std::vector<std::pair<Histogram, float>> zero_histogram() {  // the reference's own test fixture: Zero(10) at angle 0
    std::vector<std::pair<Histogram, float>> v;
    v.emplace_back(Histogram(10, 0.f), 0.f);
    return v;
  }

  FcsmResult MatchWith3DofInitial(const Rigid3d& pose_in_submap_guess, const float* hi_pts, int64_t n_hi,
                                  const float* lo_pts, int64_t n_lo, float min_score) const {
    Search s;
    s.wxy = round_to_int(o_.linear_xy_search_window / resolution_);
    s.wz = round_to_int(o_.linear_z_search_window / resolution_);
    s.lo_pts = lo_pts;
    s.n_lo = n_lo;
    s.scans.push_back(Discretize(s, hi_pts, n_hi, cast_f(pose_in_submap_guess), (float)(o_.min_rotational_score + 0.01)));
    return Finish(&s, min_score);
  }

  // Match (cc:145-162) -> MatchWithSearchParameters (:221-250): one discrete scan per yaw step inside the angular window
  // that passes the rotational score, then the same branch and bound over (scan, x, y, z).
  FcsmResult Match(const Rigid3d& global_node_pose, const Rigid3d& global_submap_pose, const float* hi_pts, int64_t n_hi,
                   const float* lo_pts, int64_t n_lo, const Histogram& scan_histogram, const Quatd& gravity_alignment,
                   float min_score) const {
    Search s;
    s.wxy = round_to_int(o_.linear_xy_search_window / resolution_);
    s.wz = round_to_int(o_.linear_z_search_window / resolution_);
    s.lo_pts = lo_pts;
    s.n_lo = n_lo;
    const Rigid3f node = cast_f(global_node_pose), submap = cast_f(global_submap_pose);
    // GenerateDiscreteScans (cc:296-350)
    float max_scan_range = 3.f * resolution_;
    for (int64_t i = 0; i < n_hi; ++i) max_scan_range = std::max(norm(V3f{hi_pts[3 * i], hi_pts[3 * i + 1], hi_pts[3 * i + 2]}), max_scan_range);
    const float kSafetyMargin = 1.f - 1e-2f;
    const float angular_step_size =
        kSafetyMargin * std::acos(1.f - (resolution_ * resolution_) / (2.f * (max_scan_range * max_scan_range)));
    const int angular_window_size = round_to_int(o_.angular_search_window / angular_step_size);  // double / float -> double
    std::vector<float> angles;
    for (int rz = -angular_window_size; rz <= angular_window_size; ++rz) angles.push_back(rz * angular_step_size);
    const Rigid3f node_to_submap = compose(inverse(submap), node);
    const Quatd ga_inv_d = eigen_inverse(gravity_alignment);  // .inverse().cast<float>()
    const Quatf ga_inv{(float)ga_inv_d.w, (float)ga_inv_d.x, (float)ga_inv_d.y, (float)ga_inv_d.z};
    const V3f dir = rotate(qmul(node_to_submap.q, ga_inv), V3f{1.f, 0.f, 0.f});  // GetYaw(rotation): atan2 of rotated UnitX
    const std::vector<float> scores = rotational_.Match(scan_histogram, std::atan2(dir.y, dir.x), angles);
    for (size_t i = 0; i != angles.size(); ++i) {
      if (scores[i] < o_.min_rotational_score) continue;
      const Quatf q = qmul(qmul(eigen_inverse(submap.q), angle_axis_to_quat(V3f{0.f, 0.f, angles[i]})), node.q);
      s.scans.push_back(Discretize(s, hi_pts, n_hi, Rigid3f{node_to_submap.t, q}, scores[i]));
    }
    if (s.scans.empty()) return FcsmResult();
    return Finish(&s, min_score);
  }

 private:
  struct Candidate {
    int scan_index = 0;
    I3 offset{0, 0, 0};
    float score = -std::numeric_limits<float>::infinity();
    float low_resolution_score = 0.f;
    bool operator>(const Candidate& other) const { return score > other.score; }
    bool operator<(const Candidate& other) const { return score < other.score; }
  };
  struct Scan {  // DiscreteScan3D
    Rigid3f pose;
    std::vector<std::vector<I3>> cells;
    float rotational_score = 0.f;
  };
  struct Search {  // the discrete scans + the window + counters of one call
    int wxy = 0, wz = 0;
    std::vector<Scan> scans;
    const float* lo_pts = nullptr;
    int64_t n_lo = 0, leaves = 0;
  };

  FcsmResult Finish(Search* s, float min_score) const {
    std::vector<Candidate> lowest = GenerateLowest(*s);
    Score(s, stack_.max_depth(), &lowest);
    const Candidate best = BranchAndBound(s, lowest, stack_.max_depth(), min_score);
    FcsmResult r;
    r.leaves_scored = s->leaves;
    if (best.score > min_score) {
      r.found = true;
      r.score = best.score;
      r.pose = cast_d(PoseFromCandidate(*s, best));
      r.rotational_score = s->scans[best.scan_index].rotational_score;
      r.low_resolution_score = best.low_resolution_score;
      r.offset = best.offset;
      r.scan_index = best.scan_index;
      r.num_scans = (int)s->scans.size();
    }
    return r;
  }

  Scan Discretize(const Search& s, const float* pts, int64_t n, const Rigid3f& pose, float rotational_score) const {
    Scan scan;
    scan.pose = pose;
    scan.rotational_score = rotational_score;
    std::vector<I3> full;
    for (int64_t i = 0; i < n; ++i)
      full.push_back(stack_.Get(0).GetCellIndex(apply(pose, V3f{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]})));
    const int full_depth = std::min(o_.full_resolution_depth, o_.branch_and_bound_depth);
    for (int i = 0; i != full_depth; ++i) scan.cells.push_back(full);
    const int low_depth = o_.branch_and_bound_depth - full_depth;
    const I3 start{-s.wxy, -s.wxy, -s.wz};
    for (int i = 0; i != low_depth; ++i) {
      const int e = i + 1;
      const I3 low_start{start.x >> e, start.y >> e, start.z >> e};
      scan.cells.emplace_back();
      for (const I3& c : full) {
        const I3 at_start{c.x + start.x, c.y + start.y, c.z + start.z};
        scan.cells.back().push_back(I3{(at_start.x >> e) - low_start.x, (at_start.y >> e) - low_start.y, (at_start.z >> e) - low_start.z});
      }
    }
    return scan;
  }

  std::vector<Candidate> GenerateLowest(const Search& s) const {
    const int step = 1 << stack_.max_depth();
    std::vector<Candidate> out;
    for (int scan_index = 0; scan_index != (int)s.scans.size(); ++scan_index)
      for (int z = -s.wz; z <= s.wz; z += step)
        for (int y = -s.wxy; y <= s.wxy; y += step)
          for (int x = -s.wxy; x <= s.wxy; x += step) {
            Candidate c;
            c.scan_index = scan_index;
            c.offset = {x, y, z};
            out.push_back(c);
          }
    return out;
  }

  void Score(Search* s, int depth, std::vector<Candidate>* candidates) const {
    const int e = std::max(0, depth - o_.full_resolution_depth + 1);
    const HybridGrid& grid = stack_.Get(depth);
    for (Candidate& c : *candidates) {
      int sum = 0;
      const I3 off{c.offset.x >> e, c.offset.y >> e, c.offset.z >> e};
      const std::vector<I3>& cells = s->scans[c.scan_index].cells[depth];
      for (const I3& cell : cells) sum += grid.value(I3{cell.x + off.x, cell.y + off.y, cell.z + off.z});
      c.score = precomp_to_probability(sum / (float)cells.size());
      if (depth == 0) ++s->leaves;
    }
    std::sort(candidates->begin(), candidates->end(), std::greater<Candidate>());
  }

  Rigid3f PoseFromCandidate(const Search& s, const Candidate& c) const {
    const Rigid3f translation{{resolution_ * (float)c.offset.x, resolution_ * (float)c.offset.y, resolution_ * (float)c.offset.z},
                              {1.f, 0.f, 0.f, 0.f}};
    return compose(translation, s.scans[c.scan_index].pose);
  }

  float LowResolutionScore(const Search& s, const Rigid3f& pose) const {
    float score = 0.f;
    for (int64_t i = 0; i < s.n_lo; ++i)
      score += lo_->GetProbability(lo_->GetCellIndex(apply(pose, V3f{s.lo_pts[3 * i], s.lo_pts[3 * i + 1], s.lo_pts[3 * i + 2]})));
    return score / (float)s.n_lo;  // float / size_t
  }

  Candidate BranchAndBound(Search* s, const std::vector<Candidate>& candidates, int depth, float min_score) const {
    if (depth == 0) {
      for (const Candidate& c : candidates) {
        if (c.score <= min_score) return Candidate();
        const float low = LowResolutionScore(*s, PoseFromCandidate(*s, c));
        if (low >= o_.min_low_resolution_score) {
          Candidate best = c;
          best.low_resolution_score = low;
          return best;
        }
      }
      return Candidate();
    }
    Candidate best;
    best.score = min_score;
    for (const Candidate& c : candidates) {
      if (c.score <= min_score) break;
      std::vector<Candidate> higher;
      const int half_width = 1 << (depth - 1);
      for (int z : {0, half_width}) {
        if (c.offset.z + z > s->wz) break;
        for (int y : {0, half_width}) {
          if (c.offset.y + y > s->wxy) break;
          for (int x : {0, half_width}) {
            if (c.offset.x + x > s->wxy) break;
            Candidate h;
            h.scan_index = c.scan_index;
            h.offset = {c.offset.x + x, c.offset.y + y, c.offset.z + z};
            higher.push_back(h);
          }
        }
      }
      Score(s, depth - 1, &higher);
      const Candidate sub = BranchAndBound(s, higher, depth - 1, best.score);
      if (best < sub) best = sub;  // std::max(a, b) keeps a unless a < b
    }
    return best;
  }

  FcsmOptions o_;
  float resolution_;
  PrecomputationGridStack stack_;
  const HybridGrid* lo_;
  RotationalScanMatcher rotational_;
};

}  // namespace orc
