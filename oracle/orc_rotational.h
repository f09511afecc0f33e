// TEST INFRASTRUCTURE — CPU oracle (see orc_math.h header).
//
// RotationalScanMatcher (SM/rotational_scan_matcher.{h,cc}): histograms of the horizontal directions between
// neighbouring points, rotated and compared by normalised dot product. Restated from
//   rotational_scan_matcher.cc:31-52 (AddValueToHistogram), :54-92 (centroid, slice accumulation), :94-121 (SortSlice),
//   :123-141 (RotateHistogram), :143-155 (MatchHistograms), :159-170 (ComputeHistogram), :172-192 (constructor, Match).
// Only the full (rotational) form of the loop-closure matcher uses it; the fork's call site passes through
// MatchWith3DofInitial, which does not (constraint_builder_3d.cc:268-277).
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <utility>
#include <vector>

#include "orc_math.h"

namespace orc {

using Histogram = std::vector<float>;

inline float histogram_norm(const Histogram& h) {  // Eigen's VectorXf::norm(): sqrt of the sum of squares
  float s = 0.f;
  for (float v : h) s += v * v;
  return std::sqrt(s);
}

inline void add_value_to_histogram(float angle, float value, Histogram* histogram) {
  while (angle > (float)M_PI) angle -= (float)M_PI;
  while (angle < 0.f) angle += (float)M_PI;
  const float zero_to_one = angle / (float)M_PI;
  const int n = (int)histogram->size();
  const int bucket = std::min(std::max(round_to_int(n * zero_to_one - 0.5f), 0), n - 1);
  (*histogram)[bucket] += value;
}

inline V3f centroid_of(const std::vector<V3f>& slice) {
  V3f sum{0.f, 0.f, 0.f};
  for (const V3f& p : slice) sum = V3f{sum.x + p.x, sum.y + p.y, sum.z + p.z};
  const float n = (float)slice.size();
  return {sum.x / n, sum.y / n, sum.z / n};
}

inline void add_slice_to_histogram(const std::vector<V3f>& slice, Histogram* histogram) {
  if (slice.empty()) return;
  const float kMinDistance = 0.2f, kMaxDistance = 0.9f;
  const V3f centroid = centroid_of(slice);
  V3f last = slice.front();
  for (const V3f& point : slice) {
    const float dx = point.x - last.x, dy = point.y - last.y;
    const float cx = point.x - centroid.x, cy = point.y - centroid.y;
    const float distance = std::sqrt(dx * dx + dy * dy);
    const float direction_norm = std::sqrt(cx * cx + cy * cy);
    if (distance < kMinDistance || direction_norm < kMinDistance) continue;
    if (distance > kMaxDistance) {
      last = point;
      continue;
    }
    const float angle = std::atan2(dy, dx);
    const float dot = (dx / distance) * (cx / direction_norm) + (dy / distance) * (cy / direction_norm);
    add_value_to_histogram(angle, std::max(0.f, 1.f - std::abs(dot)), histogram);
  }
}

inline std::vector<V3f> sort_slice(const std::vector<V3f>& slice) {
  const float kMinDistance = 0.2f;
  const V3f centroid = centroid_of(slice);
  std::vector<std::pair<float, V3f>> by_angle;
  for (const V3f& p : slice) {
    const float dx = p.x - centroid.x, dy = p.y - centroid.y;
    if (std::sqrt(dx * dx + dy * dy) < kMinDistance) continue;
    by_angle.push_back({std::atan2(dy, dx), p});
  }
  std::sort(by_angle.begin(), by_angle.end(), [](const std::pair<float, V3f>& a, const std::pair<float, V3f>& b) { return a.first < b.first; });
  std::vector<V3f> out;
  for (const auto& a : by_angle) out.push_back(a.second);
  return out;
}

inline Histogram compute_histogram(const float* pts, int64_t n, int size) {
  const float kSliceHeight = 0.2f;
  Histogram histogram(size, 0.f);
  std::map<int, std::vector<V3f>> slices;
  for (int64_t i = 0; i < n; ++i) slices[round_to_int(pts[3 * i + 2] / kSliceHeight)].push_back(V3f{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
  for (const auto& s : slices) add_slice_to_histogram(sort_slice(s.second), &histogram);
  return histogram;
}

// Rotation by a fractional number of buckets, linear interpolation between the two neighbouring whole rotations.
inline Histogram rotate_histogram(const Histogram& histogram, float angle) {
  const int n = (int)histogram.size();
  const float rotate_by_buckets = (float)(-angle * n / M_PI);  // float * int / double -> double, narrowed
  int full_buckets = round_to_int(rotate_by_buckets - 0.5f);
  const float fraction = rotate_by_buckets - full_buckets;
  while (full_buckets < 0) full_buckets += n;
  Histogram out(n);
  for (int i = 0; i < n; ++i)
    out[i] = fraction * histogram[(i + 1 + full_buckets) % n] + (1.f - fraction) * histogram[(i + full_buckets) % n];
  return out;
}

inline float match_histograms(const Histogram& submap, const Histogram& scan) {
  const float normalization = histogram_norm(scan) * histogram_norm(submap);
  if (normalization < 1e-3f) return 1.f;
  float dot = 0.f;
  for (size_t i = 0; i < submap.size(); ++i) dot += submap[i] * scan[i];
  return dot / normalization;
}

class RotationalScanMatcher {
 public:
  explicit RotationalScanMatcher(const std::vector<std::pair<Histogram, float>>& histograms_at_angles)
      : histogram_(histograms_at_angles.at(0).first.size(), 0.f) {
    for (const auto& ha : histograms_at_angles) {
      const Histogram r = rotate_histogram(ha.first, ha.second);
      for (size_t i = 0; i < r.size(); ++i) histogram_[i] += r[i];
    }
  }
  std::vector<float> Match(const Histogram& histogram, float initial_angle, const std::vector<float>& angles) const {
    std::vector<float> result;
    for (float angle : angles) result.push_back(match_histograms(histogram_, rotate_histogram(histogram, initial_angle + angle)));
    return result;
  }

 private:
  Histogram histogram_;
};

}  // namespace orc
