// TEST INFRASTRUCTURE — CPU oracle (see orc_math.h header).
//
// First-point-per-voxel filter and its adaptive (bisection) variant, restated from
//   C/sensor/internal/voxel_filter.h:34-79, voxel_filter.cc:28-131,147-150
// Points are passed as float rows of `stride` floats (3 = PointCloud, 4 = TimedPointCloud,
// 8 = RangeMeasurement{Vector4f, size_t} viewed as floats); only the first three are read.
#pragma once
#include <cstdint>
#include <unordered_set>
#include <vector>

#include "orc_math.h"

namespace orc {

struct VoxelKey {
  int32_t x, y, z;
  bool operator==(const VoxelKey& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelKeyHash {
  size_t operator()(const VoxelKey& k) const {
    uint64_t h = (uint32_t)k.x * 0x9E3779B97F4A7C15ull;
    h ^= ((uint32_t)k.y + 0x7F4A7C15u) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= ((uint32_t)k.z + 0x165667B1u) * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};

// Stateful like the reference object: the voxel set persists across Filter calls.
class VoxelFilter {
 public:
  explicit VoxelFilter(float size) : resolution_(size) {}

  // Appends the input-order indices of the surviving points to `keep`.
  void Filter(const float* pts, int64_t n, int stride, std::vector<int64_t>* keep) {
    for (int64_t i = 0; i < n; ++i) {
      const float* p = pts + i * stride;
      const I3 c = cell_index(V3f{p[0], p[1], p[2]}, resolution_);
      if (set_.insert(VoxelKey{c.x, c.y, c.z}).second) keep->push_back(i);
    }
  }

 private:
  float resolution_;
  std::unordered_set<VoxelKey, VoxelKeyHash> set_;
};

struct AdaptiveVoxelFilterOptions {
  float max_length;
  float min_num_points;  // float in the proto (adaptive_voxel_filter_options.proto)
  float max_range;
};

// Returns indices into the ORIGINAL cloud. `passes` (optional) records every voxel edge tried,
// in order, so the device implementation can be checked pass by pass.
inline std::vector<int64_t> AdaptiveVoxelFilter(const AdaptiveVoxelFilterOptions& opt, const float* pts, int64_t n,
                                                int stride, std::vector<float>* passes = nullptr) {
  // FilterByMaxRange, voxel_filter.cc:28-38
  std::vector<int64_t> in_range;
  std::vector<float> cropped;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = pts + i * stride;
    if (norm(V3f{p[0], p[1], p[2]}) <= opt.max_range) {
      in_range.push_back(i);
      cropped.insert(cropped.end(), {p[0], p[1], p[2]});
    }
  }
  const int64_t m = (int64_t)in_range.size();
  auto run = [&](float edge) {
    if (passes) passes->push_back(edge);
    std::vector<int64_t> keep;
    VoxelFilter(edge).Filter(cropped.data(), m, 3, &keep);
    return keep;
  };
  auto to_original = [&](const std::vector<int64_t>& keep) {
    std::vector<int64_t> out;
    out.reserve(keep.size());
    for (int64_t k : keep) out.push_back(in_range[k]);
    return out;
  };
  // AdaptivelyVoxelFiltered, voxel_filter.cc:40-77. size() (integer) is compared with the float option.
  if ((float)m <= opt.min_num_points) return in_range;
  std::vector<int64_t> result = run(opt.max_length);
  if ((float)result.size() >= opt.min_num_points) return to_original(result);
  for (float high_length = opt.max_length; high_length > 1e-2f * opt.max_length; high_length /= 2.f) {
    float low_length = high_length / 2.f;
    result = run(low_length);
    if ((float)result.size() >= opt.min_num_points) {
      while ((high_length - low_length) / low_length > 1e-1f) {
        const float mid_length = (low_length + high_length) / 2.f;
        std::vector<int64_t> candidate = run(mid_length);
        if ((float)candidate.size() >= opt.min_num_points) {
          low_length = mid_length;
          result = std::move(candidate);
        } else {
          high_length = mid_length;
        }
      }
      return to_original(result);
    }
  }
  return to_original(result);
}

}  // namespace orc
