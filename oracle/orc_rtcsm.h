// TEST INFRASTRUCTURE — CPU oracle (see orc_math.h header).
//
// Exhaustive 6-DoF window search on a probability grid, restated from
//   C/mapping/internal/3d/scan_matching/real_time_correlative_scan_matcher_3d.cc:34-53 (Match),
//   :55-95 (GenerateExhaustiveSearchTransforms), :97-113 (ScoreCandidate)
// Candidate linear index = (((((z+L)*(2L+1) + (y+L))*(2L+1) + (x+L))*(2A+1) + (rz+A))*(2A+1)
//                           + (ry+A))*(2A+1) + (rx+A), the reference's emplace order.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "orc_grid.h"
#include "orc_math.h"

namespace orc {

struct RtcsmOptions {
  double linear_search_window;
  double angular_search_window;
  double translation_delta_cost_weight;
  double rotation_delta_cost_weight;
};

struct RtcsmWindow {
  int linear;          // L
  int angular;         // A
  float angular_step;  // radians
  float max_scan_range;
};

inline RtcsmWindow rtcsm_window(const RtcsmOptions& opt, float resolution, const float* pts, int64_t n) {
  RtcsmWindow w;
  // double / float -> double -> lround
  w.linear = round_to_int(opt.linear_search_window / resolution);
  float max_scan_range = 3.f * resolution;
  for (int64_t i = 0; i < n; ++i) {
    const float range = norm(V3f{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
    max_scan_range = std::max(range, max_scan_range);
  }
  const float kSafetyMargin = 1.f - 1e-3f;
  w.angular_step =
      kSafetyMargin * std::acos(1.f - (resolution * resolution) / (2.f * (max_scan_range * max_scan_range)));
  w.angular = round_to_int(opt.angular_search_window / w.angular_step);
  w.max_scan_range = max_scan_range;
  return w;
}

// The relative transform of one candidate (translation offset, rotation) in reference order.
inline Rigid3f rtcsm_candidate_transform(const RtcsmWindow& w, float resolution, int x, int y, int z, int rx, int ry,
                                         int rz) {
  const V3f angle_axis{rx * w.angular_step, ry * w.angular_step, rz * w.angular_step};
  return {V3f{x * resolution, y * resolution, z * resolution}, angle_axis_to_quat(angle_axis)};
}

inline float rtcsm_score(const RtcsmOptions& opt, const HybridGrid& grid, const Rigid3f& candidate,
                         const Rigid3f& transform, const float* pts, int64_t n) {
  float score = 0.f;
  for (int64_t i = 0; i < n; ++i) {
    const V3f p = apply(candidate, V3f{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
    score += grid.GetProbability(grid.GetCellIndex(p));
  }
  score /= (float)n;
  const float angle = rotation_angle(transform.q);
  // float * double -> double; Pow2 in double; std::exp(double); float *= double narrows the product.
  const double a = norm(transform.t) * opt.translation_delta_cost_weight + angle * opt.rotation_delta_cost_weight;
  score = (float)((double)score * std::exp(-(a * a)));
  return score;
}

struct RtcsmResult {
  float score = -1.f;
  int64_t best_index = -1;
  Rigid3d pose;
  RtcsmWindow window;
};

// `all_scores` (optional) receives every candidate's score in candidate order.
inline RtcsmResult rtcsm_match(const RtcsmOptions& opt, const Rigid3d& initial, const float* pts, int64_t n,
                               const HybridGrid& grid, std::vector<float>* all_scores = nullptr) {
  RtcsmResult r;
  r.window = rtcsm_window(opt, grid.resolution(), pts, n);
  const int L = r.window.linear, A = r.window.angular;
  const Rigid3f initial_f = cast_f(initial);
  int64_t index = 0;
  for (int z = -L; z <= L; ++z)
    for (int y = -L; y <= L; ++y)
      for (int x = -L; x <= L; ++x)
        for (int rz = -A; rz <= A; ++rz)
          for (int ry = -A; ry <= A; ++ry)
            for (int rx = -A; rx <= A; ++rx, ++index) {
              const Rigid3f transform = rtcsm_candidate_transform(r.window, grid.resolution(), x, y, z, rx, ry, rz);
              const Rigid3f candidate = compose(initial_f, transform);
              const float score = rtcsm_score(opt, grid, candidate, transform, pts, n);
              if (all_scores) all_scores->push_back(score);
              if (score > r.score) {
                r.score = score;
                r.best_index = index;
                r.pose = cast_d(candidate);
              }
            }
  return r;
}

}  // namespace orc
