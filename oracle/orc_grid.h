// TEST INFRASTRUCTURE — CPU oracle (see orc_math.h header).
//
// Probability <-> uint16 value tables and the sparse three-level voxel grid, restated from
//   C/mapping/probability_values.h:30-102, probability_values.cc:27-94
//   C/mapping/3d/hybrid_grid.h:40-52 (flat index), :68-138 (8^3 leaf), :143-246 (8^3 node of
//   leaves), :251-412 (growable top level, bits 1..8), :416-547 (HybridGrid: cell index, centre,
//   SetProbability, ApplyLookupTable/FinishUpdate, iteration order used by ToProto)
// The storage here is index-vectors instead of owning pointers, but the lookup geometry
// (top cell = 64 voxels, node = 8 leaves per axis, leaf = 8 voxels per axis, z-major flat
// order, origin shift by grid_size/2, unsigned bounds test, growth by doubling) is the same,
// so value(), growth and iteration order are reproduced exactly.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#include "orc_math.h"

namespace orc {

constexpr float kMinProbability = 0.1f;
constexpr float kMaxProbability = 1.f - kMinProbability;
constexpr uint16_t kUpdateMarker = 1u << 15;

inline float clampf(float v, float lo, float hi) {
  if (v > hi) return hi;
  if (v < lo) return lo;
  return v;
}
inline float odds(float p) { return p / (1.f - p); }
inline float probability_from_odds(float o) { return o / (o + 1.f); }

// probability_values.h:30-41
inline uint16_t bounded_float_to_value(float v, float lo, float hi) {
  const int value = round_to_int((clampf(v, lo, hi) - lo) * (32766.f / (hi - lo))) + 1;
  return (uint16_t)value;
}
inline uint16_t probability_to_value(float p) { return bounded_float_to_value(p, kMinProbability, kMaxProbability); }

// probability_values.cc:27-51: 65536 entries, the upper half repeats the lower (marker bit ignored).
inline const std::vector<float>& value_to_probability_table() {
  static const std::vector<float> table = [] {
    std::vector<float> t;
    t.reserve(65536);
    for (int repeat = 0; repeat != 2; ++repeat) {
      for (int value = 0; value != 32768; ++value) {
        if (value == 0) {
          t.push_back(kMinProbability);
        } else {
          const float kScale = (kMaxProbability - kMinProbability) / 32766.f;
          t.push_back(value * kScale + (kMinProbability - kScale));
        }
      }
    }
    return t;
  }();
  return table;
}
inline float value_to_probability(uint16_t v) { return value_to_probability_table()[v]; }

// probability_values.cc:70-80
inline std::vector<uint16_t> lookup_table_to_apply_odds(float o) {
  std::vector<uint16_t> result;
  result.reserve(32768);
  result.push_back(probability_to_value(probability_from_odds(o)) + kUpdateMarker);
  for (int cell = 1; cell != 32768; ++cell) {
    result.push_back(
        probability_to_value(probability_from_odds(o * odds(value_to_probability_table()[cell]))) + kUpdateMarker);
  }
  return result;
}

class HybridGrid {
 public:
  static constexpr int kLeafBits = 3;   // FlatGrid<uint16, 3>
  static constexpr int kNodeBits = 3;   // NestedGrid<..., 3>
  static constexpr int kLeafSize = 1 << kLeafBits;               // 8
  static constexpr int kNodeSize = kLeafSize << kNodeBits;       // 64 voxels per top cell
  using Leaf = std::array<uint16_t, 512>;
  struct Node {
    std::array<int32_t, 512> leaf;  // index into leaves_, -1 = absent
    Node() { leaf.fill(-1); }
  };

  explicit HybridGrid(float resolution) : resolution_(resolution), bits_(1), top_(8, -1) {}

  float resolution() const { return resolution_; }
  int bits() const { return bits_; }
  int grid_size() const { return kNodeSize << bits_; }

  I3 GetCellIndex(const V3f& p) const { return cell_index(p, resolution_); }
  // hybrid_grid.h:446-448: index.cast<float>() * resolution
  V3f GetCenterOfCell(const I3& i) const { return {(float)i.x * resolution_, (float)i.y * resolution_, (float)i.z * resolution_}; }

  uint16_t value(const I3& index) const {
    const int half = grid_size() >> 1;
    const int sx = index.x + half, sy = index.y + half, sz = index.z + half;
    const unsigned gs = (unsigned)grid_size();
    if ((unsigned)sx >= gs || (unsigned)sy >= gs || (unsigned)sz >= gs) return 0;
    const int mx = sx / kNodeSize, my = sy / kNodeSize, mz = sz / kNodeSize;
    const int32_t n = top_[flat(mx, my, mz, bits_)];
    if (n < 0) return 0;
    const int ix = sx - mx * kNodeSize, iy = sy - my * kNodeSize, iz = sz - mz * kNodeSize;
    const int lx = ix / kLeafSize, ly = iy / kLeafSize, lz = iz / kLeafSize;
    const int32_t l = nodes_[n]->leaf[flat(lx, ly, lz, kNodeBits)];
    if (l < 0) return 0;
    return (*leaves_[l])[flat(ix - lx * kLeafSize, iy - ly * kLeafSize, iz - lz * kLeafSize, kLeafBits)];
  }

  uint16_t* mutable_value(const I3& index) {
    for (;;) {
      const int half = grid_size() >> 1;
      const int sx = index.x + half, sy = index.y + half, sz = index.z + half;
      const unsigned gs = (unsigned)grid_size();
      if ((unsigned)sx >= gs || (unsigned)sy >= gs || (unsigned)sz >= gs) {
        Grow();
        continue;
      }
      const int mx = sx / kNodeSize, my = sy / kNodeSize, mz = sz / kNodeSize;
      int32_t& n = top_[flat(mx, my, mz, bits_)];
      if (n < 0) {
        n = (int32_t)nodes_.size();
        nodes_.emplace_back(new Node());
      }
      const int ix = sx - mx * kNodeSize, iy = sy - my * kNodeSize, iz = sz - mz * kNodeSize;
      const int lx = ix / kLeafSize, ly = iy / kLeafSize, lz = iz / kLeafSize;
      int32_t& l = nodes_[n]->leaf[flat(lx, ly, lz, kNodeBits)];
      if (l < 0) {
        l = (int32_t)leaves_.size();
        leaves_.emplace_back(new Leaf());
        leaves_.back()->fill(0);
      }
      return &(*leaves_[l])[flat(ix - lx * kLeafSize, iy - ly * kLeafSize, iz - lz * kLeafSize, kLeafBits)];
    }
  }

  void SetProbability(const I3& i, float p) { *mutable_value(i) = probability_to_value(p); }
  float GetProbability(const I3& i) const { return value_to_probability(value(i)); }
  bool IsKnown(const I3& i) const { return value(i) != 0; }

  // hybrid_grid.h:494-520
  bool ApplyLookupTable(const I3& i, const std::vector<uint16_t>& table) {
    uint16_t* cell = mutable_value(i);
    if (*cell >= kUpdateMarker) return false;
    update_cells_.push_back(cell);
    *cell = table[*cell];
    return true;
  }
  void FinishUpdate() {
    while (!update_cells_.empty()) {
      *update_cells_.back() -= kUpdateMarker;
      update_cells_.pop_back();
    }
  }

  // Visits non-zero cells in the order of the reference's nested iterators
  // (top cells in flat order, then leaves in flat order, then voxels in flat order).
  template <typename F>
  void ForEachCell(F&& f) const {
    const int half_top = (1 << (bits_ - 1)) * kNodeSize;
    const int ntop = 1 << (3 * bits_);
    const int tmask = (1 << bits_) - 1;
    for (int t = 0; t < ntop; ++t) {
      if (top_[t] < 0) continue;
      const int tx = t & tmask, ty = (t >> bits_) & tmask, tz = (t >> bits_) >> bits_;
      const Node& node = *nodes_[top_[t]];
      for (int l = 0; l < 512; ++l) {
        if (node.leaf[l] < 0) continue;
        const int lx = l & 7, ly = (l >> 3) & 7, lz = l >> 6;
        const Leaf& leaf = *leaves_[node.leaf[l]];
        for (int c = 0; c < 512; ++c) {
          if (leaf[c] == 0) continue;
          const int cx = c & 7, cy = (c >> 3) & 7, cz = c >> 6;
          f(I3{tx * kNodeSize + lx * kLeafSize + cx - half_top, ty * kNodeSize + ly * kLeafSize + cy - half_top,
               tz * kNodeSize + lz * kLeafSize + cz - half_top},
            leaf[c]);
        }
      }
    }
  }

 private:
  static int flat(int x, int y, int z, int bits) { return (((z << bits) + y) << bits) + x; }

  // hybrid_grid.h:389-407: double every axis, old content moves to the centre.
  void Grow() {
    const int new_bits = bits_ + 1;
    if (new_bits > 8) throw std::runtime_error("HybridGrid: CHECK_LE(new_bits, 8) failed");
    std::vector<int32_t> grown((size_t)8 * top_.size(), -1);
    for (int z = 0; z != (1 << bits_); ++z)
      for (int y = 0; y != (1 << bits_); ++y)
        for (int x = 0; x != (1 << bits_); ++x) {
          const int o = 1 << (bits_ - 1);
          grown[flat(x + o, y + o, z + o, new_bits)] = top_[flat(x, y, z, bits_)];
        }
    top_ = std::move(grown);
    bits_ = new_bits;
  }

  const float resolution_;
  int bits_;
  std::vector<int32_t> top_;
  std::vector<std::unique_ptr<Node>> nodes_;
  std::vector<std::unique_ptr<Leaf>> leaves_;
  std::vector<uint16_t*> update_cells_;
};

}  // namespace orc
