// Header-only C++ mirror of the reference's matcher / filter classes over the dliom_b200 C-ABI.
//
// Same class names, argument meaning and call shapes as the reference interfaces they replace
// (C/ = src/cartographer/cartographer/, SM/ = C/mapping/internal/3d/scan_matching/):
//   sensor::VoxelFilter, sensor::AdaptiveVoxelFilter        C/sensor/internal/voxel_filter.h:34-79
//   scan_matching::RealTimeCorrelativeScanMatcher3D          SM/real_time_correlative_scan_matcher_3d.h:33-60
//   scan_matching::CeresScanMatcher3D                        SM/ceres_scan_matcher_3d.h:34-61
//   mapping::RangeDataSynchronizer                           C/mapping/internal/3d/range_data_synchronizer.h:30-68
//   mapping::LocalTrajectoryBuilder3D                        C/mapping/internal/3d/local_trajectory_builder_3d.h:81-113
// Eigen / protobuf types are replaced by the plain structs below (this image has neither); INTEGRATION.md shows
// the three-line adapters for Eigen::Vector3f / transform::Rigid3d / proto options in a real Cartographer tree.
// Errors: the reference CHECK-aborts; this shim throws dliom::Error carrying the C-ABI status and message.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <deque>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <memory>
#include <vector>

#include "../../include/dliom_b200.h"

namespace dliom {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

struct Rigid3d {  // transform::Rigid3d: translation + rotation quaternion (w, x, y, z)
  double t[3] = {0, 0, 0};
  double q[4] = {1, 0, 0, 0};
  void to7(double* p) const { for (int i = 0; i < 3; ++i) p[i] = t[i]; for (int i = 0; i < 4; ++i) p[3 + i] = q[i]; }
  static Rigid3d from7(const double* p) { Rigid3d r; for (int i = 0; i < 3; ++i) r.t[i] = p[i]; for (int i = 0; i < 4; ++i) r.q[i] = p[3 + i]; return r; }
};
using PointCloud = std::vector<std::array<float, 3>>;       // sensor::PointCloud
using TimedPointCloud = std::vector<std::array<float, 4>>;  // sensor::TimedPointCloud

class Context {  // one per host thread (re-entrancy contract of CeresScanMatcher3D::Match)
 public:
  explicit Context(int device = 0) {
    const int st = dl_context_create(device, &ctx_);
    if (st != DL_OK) throw Error(st, dl_last_error(nullptr));
  }
  ~Context() { dl_context_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  dl_context* get() const { return ctx_; }
  void check(int st) const { if (st != DL_OK) throw Error(st, dl_last_error(ctx_)); }
 private:
  dl_context* ctx_ = nullptr;
};

// Device mirror of a mapping::HybridGrid. Fill from `for (auto it : hybrid_grid)` (x, y, z, value), or from
// HybridGrid::ToProto()'s parallel arrays; call Update() with the touched cells after each InsertRangeData.
class DeviceHybridGrid {
 public:
  DeviceHybridGrid(Context* ctx, float resolution) : ctx_(ctx) { ctx->check(dl_grid_create(ctx->get(), resolution, &grid_)); }
  ~DeviceHybridGrid() { dl_grid_destroy(grid_); }
  DeviceHybridGrid(const DeviceHybridGrid&) = delete;
  DeviceHybridGrid& operator=(const DeviceHybridGrid&) = delete;
  void Update(const std::vector<int32_t>& x, const std::vector<int32_t>& y, const std::vector<int32_t>& z,
              const std::vector<uint16_t>& value) {
    ctx_->check(dl_grid_set_cells(grid_, (int64_t)x.size(), x.data(), y.data(), z.data(), value.data()));
    ctx_->check(dl_grid_sync(grid_));
  }
  float resolution() const { return dl_grid_resolution(grid_); }
  const dl_grid* get() const { return grid_; }
 private:
  Context* ctx_;
  dl_grid* grid_ = nullptr;
};

namespace sensor {

class VoxelFilter {  // voxel_filter.h:34-62 (hot-path use: a temporary per call, LTB:393-395, :479-484)
 public:
  VoxelFilter(Context* ctx, float size) : ctx_(ctx), resolution_(size) {}
  PointCloud Filter(const PointCloud& point_cloud) { return FilterRows(point_cloud); }
  TimedPointCloud Filter(const TimedPointCloud& timed_point_cloud) { return FilterRows(timed_point_cloud); }
 private:
  template <typename Cloud>
  Cloud FilterRows(const Cloud& in) {
    std::vector<int64_t> keep(in.size() ? in.size() : 1);
    int64_t n_keep = 0;
    ctx_->check(dl_voxel_filter(ctx_->get(), in.empty() ? nullptr : in[0].data(), (int64_t)in.size(),
                                (int)(sizeof(in[0]) / sizeof(float)), resolution_, keep.data(), &n_keep));
    Cloud out;
    out.reserve(n_keep);
    for (int64_t i = 0; i < n_keep; ++i) out.push_back(in[keep[i]]);
    return out;
  }
  Context* ctx_;
  float resolution_;
};

struct AdaptiveVoxelFilterOptions {  // proto::AdaptiveVoxelFilterOptions
  float max_length, min_num_points, max_range;
};
class AdaptiveVoxelFilter {  // voxel_filter.h:66-79
 public:
  AdaptiveVoxelFilter(Context* ctx, const AdaptiveVoxelFilterOptions& options) : ctx_(ctx), options_(options) {}
  PointCloud Filter(const PointCloud& point_cloud) const {
    std::vector<int64_t> keep(point_cloud.size() ? point_cloud.size() : 1);
    int64_t n_keep = 0;
    const dl_adaptive_voxel_filter_options o{options_.max_length, options_.min_num_points, options_.max_range};
    ctx_->check(dl_adaptive_voxel_filter(ctx_->get(), &o, point_cloud.empty() ? nullptr : point_cloud[0].data(),
                                         (int64_t)point_cloud.size(), 3, keep.data(), &n_keep, nullptr, nullptr));
    PointCloud out;
    out.reserve(n_keep);
    for (int64_t i = 0; i < n_keep; ++i) out.push_back(point_cloud[keep[i]]);
    return out;
  }
 private:
  Context* ctx_;
  AdaptiveVoxelFilterOptions options_;
};

}  // namespace sensor

namespace scan_matching {

struct RealTimeCorrelativeScanMatcherOptions {  // proto::RealTimeCorrelativeScanMatcherOptions
  double linear_search_window, angular_search_window, translation_delta_cost_weight, rotation_delta_cost_weight;
};
class RealTimeCorrelativeScanMatcher3D {  // real_time_correlative_scan_matcher_3d.h:33-60
 public:
  RealTimeCorrelativeScanMatcher3D(Context* ctx, const RealTimeCorrelativeScanMatcherOptions& options)
      : ctx_(ctx), options_(options) {}
  // Returns the best score; *pose_estimate receives the best candidate (CHECK_NOTNULL in the reference).
  float Match(const Rigid3d& initial_pose_estimate, const PointCloud& point_cloud, const DeviceHybridGrid& hybrid_grid,
              Rigid3d* pose_estimate) const {
    if (!pose_estimate) throw Error(DL_ERR_ARG, "pose_estimate is null");
    const dl_rtcsm_options o{options_.linear_search_window, options_.angular_search_window,
                             options_.translation_delta_cost_weight, options_.rotation_delta_cost_weight};
    double init[7], out[7];
    initial_pose_estimate.to7(init);
    float score = 0.f;
    ctx_->check(dl_rtcsm_match(ctx_->get(), &o, init, point_cloud.empty() ? nullptr : point_cloud[0].data(),
                               (int64_t)point_cloud.size(), hybrid_grid.get(), out, &score, nullptr, nullptr));
    *pose_estimate = Rigid3d::from7(out);
    return score;
  }
 private:
  Context* ctx_;
  RealTimeCorrelativeScanMatcherOptions options_;
};

struct CeresScanMatcherOptions3D {  // proto::CeresScanMatcherOptions3D + common::proto::CeresSolverOptions
  std::vector<double> occupied_space_weight;
  double translation_weight = 5., rotation_weight = 4e2;
  bool only_optimize_yaw = false;
  bool use_nonmonotonic_steps = false;
  int max_num_iterations = 12;
  int num_threads = 1;
};
struct SolverSummary {  // the fields of ceres::Solver::Summary the reference reads (LTB:543) + counters
  double initial_cost = 0, final_cost = 0;
  int num_iterations = 0, num_successful_steps = 0, num_unsuccessful_steps = 0, termination = 1;
};
using PointCloudAndHybridGridPointers = std::pair<const PointCloud*, const DeviceHybridGrid*>;

class CeresScanMatcher3D {  // ceres_scan_matcher_3d.h:41-61
 public:
  CeresScanMatcher3D(Context* ctx, const CeresScanMatcherOptions3D& options) : ctx_(ctx), options_(options) {}
  void Match(const std::array<double, 3>& target_translation, const Rigid3d& initial_pose_estimate,
             const std::vector<PointCloudAndHybridGridPointers>& point_clouds_and_hybrid_grids,
             Rigid3d* pose_estimate, SolverSummary* summary) const {
    dl_ceres_options o{};
    o.num_occupied_space_weights = (int32_t)options_.occupied_space_weight.size();
    for (size_t i = 0; i < options_.occupied_space_weight.size() && i < DL_MAX_PAIRS; ++i)
      o.occupied_space_weight[i] = options_.occupied_space_weight[i];
    o.translation_weight = options_.translation_weight;
    o.rotation_weight = options_.rotation_weight;
    o.only_optimize_yaw = options_.only_optimize_yaw;
    o.use_nonmonotonic_steps = options_.use_nonmonotonic_steps;
    o.max_num_iterations = options_.max_num_iterations;
    o.num_threads = options_.num_threads;
    std::vector<const float*> clouds;
    std::vector<int64_t> sizes;
    std::vector<const dl_grid*> grids;
    for (const auto& pg : point_clouds_and_hybrid_grids) {
      clouds.push_back(pg.first->empty() ? nullptr : (*pg.first)[0].data());
      sizes.push_back((int64_t)pg.first->size());
      grids.push_back(pg.second->get());
    }
    double init[7], out[7];
    initial_pose_estimate.to7(init);
    dl_solve_summary s{};
    ctx_->check(dl_ceres_match(ctx_->get(), &o, target_translation.data(), init, (int32_t)clouds.size(), clouds.data(),
                               sizes.data(), grids.data(), out, &s));
    *pose_estimate = Rigid3d::from7(out);
    if (summary)
      *summary = {s.initial_cost, s.final_cost, s.num_iterations, s.num_successful_steps, s.num_unsuccessful_steps,
                  s.termination};
  }
 private:
  Context* ctx_;
  CeresScanMatcherOptions3D options_;
};

struct FastCorrelativeScanMatcherOptions3D {  // proto::FastCorrelativeScanMatcherOptions3D (pose_graph.lua:49-57)
  int branch_and_bound_depth = 8, full_resolution_depth = 3;
  double min_rotational_score = 0.77, min_low_resolution_score = 0.55;
  double linear_xy_search_window = 5., linear_z_search_window = 1., angular_search_window = 0.2617993877991494;
  dl_fcsm_options c() const {
    return {branch_and_bound_depth, full_resolution_depth, min_rotational_score, min_low_resolution_score,
            linear_xy_search_window, linear_z_search_window, angular_search_window};
  }
};

// fast_correlative_scan_matcher_3d.h:59-160, the entry point this fork calls (MatchWith3DofInitial). The constructor takes
// the two grids like the reference's; no precomputation stack is built — the device scores the whole window.
class FastCorrelativeScanMatcher3D {
 public:
  struct Result {
    float score;
    Rigid3d pose_estimate;
    float rotational_score;
    float low_resolution_score;
  };
  FastCorrelativeScanMatcher3D(Context* ctx, const DeviceHybridGrid& hybrid_grid, const DeviceHybridGrid* low_resolution_hybrid_grid,
                               const FastCorrelativeScanMatcherOptions3D& options)
      : ctx_(ctx), hi_(&hybrid_grid), lo_(low_resolution_hybrid_grid), options_(options) {}
  // nullptr when no leaf above min_score passes the low-resolution gate, like the reference.
  std::unique_ptr<Result> MatchWith3DofInitial(const Rigid3d& pose_in_submap_guess, const PointCloud& high_resolution_point_cloud,
                                               const PointCloud& low_resolution_point_cloud, float min_score) const {
    const dl_fcsm_options o = options_.c();
    double guess[7];
    pose_in_submap_guess.to7(guess);
    dl_fcsm_result r{};
    ctx_->check(dl_fcsm_match_3dof(ctx_->get(), &o, guess,
                                   high_resolution_point_cloud.empty() ? nullptr : high_resolution_point_cloud[0].data(),
                                   (int64_t)high_resolution_point_cloud.size(),
                                   low_resolution_point_cloud.empty() ? nullptr : low_resolution_point_cloud[0].data(),
                                   (int64_t)low_resolution_point_cloud.size(), hi_->get(), lo_->get(), min_score, &r));
    if (!r.found) return nullptr;
    return std::unique_ptr<Result>(new Result{r.score, Rigid3d::from7(r.pose_estimate), r.rotational_score, r.low_resolution_score});
  }
  // fast_correlative_scan_matcher_3d.h:85-95: the full search (yaw steps x translation window). `submap_histogram` is what the
  // reference's RotationalScanMatcher accumulates from the submap's nodes; `scan_histogram` / `gravity_alignment` (w x y z) come
  // from the node's TrajectoryNode::Data.
  std::unique_ptr<Result> Match(const Rigid3d& global_node_pose, const Rigid3d& global_submap_pose,
                                const std::vector<float>& submap_histogram, const std::vector<float>& scan_histogram,
                                const std::array<double, 4>& gravity_alignment, const PointCloud& high_resolution_point_cloud,
                                const PointCloud& low_resolution_point_cloud, float min_score) const {
    if (submap_histogram.size() != scan_histogram.size() || submap_histogram.empty()) throw Error(DL_ERR_ARG, "histogram sizes differ");
    const dl_fcsm_options o = options_.c();
    double node[7], submap[7];
    global_node_pose.to7(node);
    global_submap_pose.to7(submap);
    dl_fcsm_result r{};
    ctx_->check(dl_fcsm_match(ctx_->get(), &o, submap_histogram.data(), scan_histogram.data(), (int32_t)scan_histogram.size(), node,
                              submap, gravity_alignment.data(),
                              high_resolution_point_cloud.empty() ? nullptr : high_resolution_point_cloud[0].data(),
                              (int64_t)high_resolution_point_cloud.size(),
                              low_resolution_point_cloud.empty() ? nullptr : low_resolution_point_cloud[0].data(),
                              (int64_t)low_resolution_point_cloud.size(), hi_->get(), lo_->get(), min_score, &r));
    if (!r.found) return nullptr;
    return std::unique_ptr<Result>(new Result{r.score, Rigid3d::from7(r.pose_estimate), r.rotational_score, r.low_resolution_score});
  }
 private:
  Context* ctx_;
  const DeviceHybridGrid* hi_;
  const DeviceHybridGrid* lo_;
  FastCorrelativeScanMatcherOptions3D options_;
};

}  // namespace scan_matching

namespace constraints {

struct ConstraintBuilderOptions {  // proto::ConstraintBuilderOptions, the fields ComputeConstraint reads (pose_graph.lua:17-73)
  double min_score = 0.55;
  double loop_closure_translation_weight = 1.1e4, loop_closure_rotation_weight = 1e5;
  scan_matching::FastCorrelativeScanMatcherOptions3D fast_correlative_scan_matcher_options_3d;
  scan_matching::CeresScanMatcherOptions3D ceres_scan_matcher_options_3d;
  ConstraintBuilderOptions() {
    ceres_scan_matcher_options_3d.occupied_space_weight = {5., 30.};
    ceres_scan_matcher_options_3d.translation_weight = 10.;
    ceres_scan_matcher_options_3d.rotation_weight = 1.;
    ceres_scan_matcher_options_3d.max_num_iterations = 10;
  }
};

struct Constraint {  // PoseGraphInterface::Constraint (INTER_SUBMAP)
  int submap_index, node_index;
  Rigid3d zbar_ij;
  double translation_weight, rotation_weight;
  float score, low_resolution_score;
};

// The compute half of ConstraintBuilder3D (constraint_builder_3d.cc:202-333): queue (node, submap) searches with
// MaybeAddConstraint, then Compute() runs all of them in one device batch and returns the constraints found
// (the reference schedules one thread-pool task per pair and collects them in RunWhenDoneCallback, :335-358).
class ConstraintBuilder3D {
 public:
  ConstraintBuilder3D(Context* ctx, const ConstraintBuilderOptions& options) : ctx_(ctx), options_(options) {}
  void MaybeAddConstraint(int submap_index, const DeviceHybridGrid* high_resolution_grid, const DeviceHybridGrid* low_resolution_grid,
                          int node_index, const PointCloud& high_resolution_point_cloud,
                          const PointCloud& low_resolution_point_cloud, const Rigid3d& node_pose_in_submap_guess) {
    ids_.push_back({submap_index, node_index});
    hi_grids_.push_back(high_resolution_grid->get());
    lo_grids_.push_back(low_resolution_grid->get());
    for (const auto& p : high_resolution_point_cloud) hi_.insert(hi_.end(), p.begin(), p.end());
    for (const auto& p : low_resolution_point_cloud) lo_.insert(lo_.end(), p.begin(), p.end());
    hi_off_.push_back((int64_t)hi_.size() / 3);
    lo_off_.push_back((int64_t)lo_.size() / 3);
    double g[7];
    node_pose_in_submap_guess.to7(g);
    guesses_.insert(guesses_.end(), g, g + 7);
  }
  int GetNumQueuedSearches() const { return (int)ids_.size(); }
  std::vector<Constraint> Compute() {
    dl_constraint_options o{};
    o.min_score = options_.min_score;
    o.loop_closure_translation_weight = options_.loop_closure_translation_weight;
    o.loop_closure_rotation_weight = options_.loop_closure_rotation_weight;
    o.fast_correlative_scan_matcher_3d = options_.fast_correlative_scan_matcher_options_3d.c();
    const auto& c = options_.ceres_scan_matcher_options_3d;
    o.ceres_scan_matcher_3d.num_occupied_space_weights = (int32_t)c.occupied_space_weight.size();
    for (size_t i = 0; i < c.occupied_space_weight.size() && i < DL_MAX_PAIRS; ++i)
      o.ceres_scan_matcher_3d.occupied_space_weight[i] = c.occupied_space_weight[i];
    o.ceres_scan_matcher_3d.translation_weight = c.translation_weight;
    o.ceres_scan_matcher_3d.rotation_weight = c.rotation_weight;
    o.ceres_scan_matcher_3d.only_optimize_yaw = c.only_optimize_yaw;
    o.ceres_scan_matcher_3d.use_nonmonotonic_steps = c.use_nonmonotonic_steps;
    o.ceres_scan_matcher_3d.max_num_iterations = c.max_num_iterations;
    o.ceres_scan_matcher_3d.num_threads = c.num_threads;
    std::vector<dl_constraint> raw(ids_.size());
    ctx_->check(dl_constraint_search_batch(ctx_->get(), &o, (int32_t)ids_.size(), guesses_.data(), hi_.data(), hi_off_.data(),
                                           lo_.data(), lo_off_.data(), hi_grids_.data(), lo_grids_.data(), raw.data()));
    std::vector<Constraint> out;
    for (size_t k = 0; k < raw.size(); ++k)
      if (raw[k].found)
        out.push_back({ids_[k][0], ids_[k][1], Rigid3d::from7(raw[k].pose), raw[k].translation_weight, raw[k].rotation_weight,
                       raw[k].score, raw[k].low_resolution_score});
    ids_.clear(); hi_grids_.clear(); lo_grids_.clear(); hi_.clear(); lo_.clear(); guesses_.clear();
    hi_off_.assign(1, 0); lo_off_.assign(1, 0);
    return out;
  }
 private:
  Context* ctx_;
  ConstraintBuilderOptions options_;
  std::vector<std::array<int, 2>> ids_;
  std::vector<const dl_grid*> hi_grids_, lo_grids_;
  std::vector<float> hi_, lo_;
  std::vector<int64_t> hi_off_{0}, lo_off_{0};
  std::vector<double> guesses_;
};

}  // namespace constraints

namespace sensor {
struct ImuData {  // sensor::ImuData (C/sensor/imu_data.h)
  double time;
  std::array<double, 3> linear_acceleration, angular_velocity;
};
struct TimedPointCloudData {  // sensor::TimedPointCloudData (C/sensor/timed_point_cloud_data.h:27-31)
  double time;  // acquisition time of the LAST point; ranges[i][3] <= 0 is relative to it
  std::array<float, 3> origin;
  TimedPointCloud ranges;
};
struct RangeMeasurement {  // TimedPointCloudOriginData::RangeMeasurement (timed_point_cloud_data.h:33-42), 32 bytes
  std::array<float, 4> point_time;
  uint64_t origin_index;
  uint64_t pad_ = 0;
};
static_assert(sizeof(RangeMeasurement) == 32, "RangeMeasurement layout");
struct TimedPointCloudOriginData {
  double time = 0;
  std::vector<std::array<float, 3>> origins;
  std::vector<RangeMeasurement> ranges;
};
struct RangeData {  // sensor::RangeData
  std::array<float, 3> origin;
  PointCloud returns, misses;
};
}  // namespace sensor

namespace mapping {

// range_data_synchronizer.{h,cc}: the FIRST expected sensor is the prior LiDAR; clouds of the others are queued and the part of
// the oldest queued cloud that overlaps the prior cloud's sweep [start, end] is merged into it, re-stamped relative to the prior
// cloud's end and sorted by time. Host bookkeeping before the path, kept on the host here too.
class RangeDataSynchronizer {
 public:
  explicit RangeDataSynchronizer(const std::vector<std::string>& expected_range_sensor_ids)
      : expected_sensor_ids_(expected_range_sensor_ids.begin(), expected_range_sensor_ids.end()),
        prior_sensor_id_(expected_range_sensor_ids.empty() ? std::string() : expected_range_sensor_ids.front()) {}

  sensor::TimedPointCloudOriginData AddRangeData(const std::string& sensor_id, const sensor::TimedPointCloudData& data, bool deskew) {
    if (!expected_sensor_ids_.count(sensor_id)) throw Error(DL_ERR_ARG, "unexpected range sensor id (CHECK_NE)");
    sensor::TimedPointCloudOriginData result;
    sensor::TimedPointCloudData cloud = data;
    if (deskew) StampRangeData(&cloud, 0.1);
    if (sensor_id != prior_sensor_id_) {
      secondary_.push_back(cloud);
      return result;
    }
    if (secondary_.empty() || cloud.ranges.empty()) return Single(cloud);
    const double end = cloud.time, start = end + cloud.ranges.front()[3];
    while (!secondary_.empty() && secondary_.front().time < start) secondary_.pop_front();  // pop old clouds
    if (secondary_.empty()) return Single(cloud);
    const sensor::TimedPointCloudData& sec = secondary_.front();
    if (sec.ranges.empty() || sec.time + sec.ranges.front()[3] > end) return Single(cloud);  // "secondary lidar may be too fast"
    int i_start = -1, i_end = -1;
    for (int i = 0; i < (int)sec.ranges.size(); ++i) {
      const double t = sec.time + sec.ranges[i][3];
      if (t >= start && t <= end && i_start == -1) i_start = i;
      if (i_start != -1 && t > end) {
        i_end = i - 1;
        break;
      }
    }
    if (i_start == -1) throw Error(DL_ERR_ARG, "no overlap between the two sweeps (CHECK)");
    if (i_end == -1) i_end = (int)sec.ranges.size() - 1;
    result.time = cloud.time;
    result.origins = {cloud.origin, sec.origin};
    result.ranges.reserve(cloud.ranges.size() + (size_t)(i_end - i_start + 1));
    for (size_t i = 0; i < cloud.ranges.size(); ++i) result.ranges.push_back({data.ranges[i], 0});  // the prior cloud keeps its own times
    for (int i = i_start; i <= i_end; ++i) {
      sensor::RangeMeasurement m{sec.ranges[i], 1};
      const double relative_t = sec.ranges[i][3];
      m.point_time[3] = (float)(relative_t + sec.time - end);
      result.ranges.push_back(m);
    }
    std::stable_sort(result.ranges.begin(), result.ranges.end(),
                     [](const sensor::RangeMeasurement& a, const sensor::RangeMeasurement& b) { return a.point_time[3] < b.point_time[3]; });
    return result;
  }

 private:
  static sensor::TimedPointCloudOriginData Single(const sensor::TimedPointCloudData& cloud) {  // ToTimedPointCloudOriginData
    sensor::TimedPointCloudOriginData r;
    r.time = cloud.time;
    r.origins = {cloud.origin};
    r.ranges.reserve(cloud.ranges.size());
    for (const auto& p : cloud.ranges) r.ranges.push_back({p, 0});
    return r;
  }
  static void StampRangeData(sensor::TimedPointCloudData* cloud, double scan_period) {  // :117-131, the "very naive" stamping
    const int n = (int)cloud->ranges.size();
    if (n < 2) return;
    const double duration = scan_period / (n - 1);
    for (int i = 0; i < n; ++i) cloud->ranges[i][3] = (float)(-scan_period + i * duration);
    cloud->ranges.back()[3] = 0.f;
  }
  const std::set<std::string> expected_sensor_ids_;
  const std::string prior_sensor_id_;
  std::deque<sensor::TimedPointCloudData> secondary_;
};

struct LocalTrajectoryBuilderOptions3D {  // proto::LocalTrajectoryBuilderOptions3D, the fields the path reads
  dl_ltb_options c{};
  bool enable_manual_deskew = false;      // eable_mannually_discrew: re-stamp the points uniformly over the scan period
  LocalTrajectoryBuilderOptions3D() {     // trajectory_builder_3d.lua defaults
    dl_frontend_options& f = c.frontend;
    f.min_range = 1.f; f.max_range = 60.f; f.voxel_filter_size = 0.15f;
    f.high_resolution_adaptive_voxel_filter = {2.f, 150.f, 15.f};
    f.low_resolution_adaptive_voxel_filter = {4.f, 200.f, 60.f};
    f.scan_period = 0.1;
    f.ceres_scan_matcher.num_occupied_space_weights = 2;
    f.ceres_scan_matcher.occupied_space_weight[0] = 1.; f.ceres_scan_matcher.occupied_space_weight[1] = 6.;
    f.ceres_scan_matcher.translation_weight = 5.; f.ceres_scan_matcher.rotation_weight = 4e2;
    f.ceres_scan_matcher.max_num_iterations = 12; f.ceres_scan_matcher.num_threads = 1;
    c.imu_noise = {3.99e-2, 1.56e-2, 6.4e-5, 3.6e-5};
    c.imu_weight = 1.; c.gravity = 9.8;
    c.high_resolution = 0.1f; c.low_resolution = 0.45f; c.num_range_data = 160; c.high_resolution_max_range = 20;
    c.range_data_inserter = {0.55, 0.49, 2, 0};
    c.motion_filter_max_time_seconds = 0.5; c.motion_filter_max_distance_meters = 0.1; c.motion_filter_max_angle_radians = 0.004;
    c.rotational_histogram_size = 120; c.frames_for_static_initialization = 7;
    c.two_stage = 0;  // 1 = the reference's chain: plain match, then the window update (dl_window_optimize_batch)
    c.ceres_pose_noise_t = 1e-2; c.ceres_pose_noise_r = 1e-2; c.prior_pose_noise = 1e-2; c.prior_velocity_noise = 1e4; c.prior_bias_noise = 1e-2;
  }
};

struct TrajectoryNodeData {  // mapping::TrajectoryNode::Data (C/mapping/trajectory_node.h)
  double time;
  std::array<double, 4> gravity_alignment;  // w x y z
  PointCloud high_resolution_point_cloud, low_resolution_point_cloud;
  std::vector<float> rotational_scan_matcher_histogram;
  Rigid3d local_pose;
};

class LocalTrajectoryBuilder3D {  // local_trajectory_builder_3d.h:81-113
 public:
  struct InsertionResult {
    std::shared_ptr<const TrajectoryNodeData> constant_data;
    std::vector<int> insertion_submaps;  // indices; the grids are reachable through submap()
  };
  struct MatchingResult {
    double time;
    Rigid3d local_pose;
    sensor::RangeData range_data_in_local;
    std::unique_ptr<const InsertionResult> insertion_result;  // nullptr if dropped by the motion filter
  };
  LocalTrajectoryBuilder3D(Context* ctx, const LocalTrajectoryBuilderOptions3D& options,
                           const std::vector<std::string>& expected_range_sensor_ids)
      : ctx_(ctx), options_(options), synchronizer_(expected_range_sensor_ids) {
    ctx->check(dl_ltb_create(ctx->get(), &options.c, &builder_));
  }
  ~LocalTrajectoryBuilder3D() { dl_ltb_destroy(builder_); }
  LocalTrajectoryBuilder3D(const LocalTrajectoryBuilder3D&) = delete;
  LocalTrajectoryBuilder3D& operator=(const LocalTrajectoryBuilder3D&) = delete;

  void AddImuData(const sensor::ImuData& imu_data) {
    ctx_->check(dl_ltb_add_imu_data(builder_, imu_data.time, imu_data.linear_acceleration.data(), imu_data.angular_velocity.data()));
  }
  // Returns nullptr while initialising, when no IMU arrived since the last scan, when the secondary LiDAR's cloud was only
  // queued, or when the scan was dropped (empty filtered clouds), like the reference.
  std::unique_ptr<MatchingResult> AddRangeData(const std::string& sensor_id, const sensor::TimedPointCloudData& unsynchronized_data) {
    const sensor::TimedPointCloudOriginData data = synchronizer_.AddRangeData(sensor_id, unsynchronized_data, options_.enable_manual_deskew);
    if (data.ranges.empty()) return nullptr;
    dl_matching_result r{};
    ctx_->check(dl_ltb_add_synchronized_range_data(builder_, data.time, data.ranges.data(), (int64_t)data.ranges.size(), 8,
                                                   data.origins[0].data(), (int32_t)data.origins.size(), &r));
    if (!r.has_result) return nullptr;
    std::unique_ptr<MatchingResult> out(new MatchingResult);
    out->time = r.time;
    out->local_pose = Rigid3d::from7(r.local_pose);
    out->range_data_in_local.origin = {r.origin_in_local[0], r.origin_in_local[1], r.origin_in_local[2]};
    out->range_data_in_local.returns = Cloud(0);
    out->range_data_in_local.misses = Cloud(1);
    if (r.inserted) {
      auto node = std::make_shared<TrajectoryNodeData>();
      node->time = r.time;
      node->gravity_alignment = {r.local_pose[3], r.local_pose[4], r.local_pose[5], r.local_pose[6]};
      node->high_resolution_point_cloud = Cloud(2);
      node->low_resolution_point_cloud = Cloud(3);
      node->rotational_scan_matcher_histogram.resize(options_.c.rotational_histogram_size);
      ctx_->check(dl_ltb_get_histogram(builder_, node->rotational_scan_matcher_histogram.data(), options_.c.rotational_histogram_size));
      node->local_pose = out->local_pose;
      std::unique_ptr<InsertionResult> ins(new InsertionResult);
      ins->constant_data = node;
      for (int k = 0; k < r.num_insertion_submaps; ++k) ins->insertion_submaps.push_back(r.insertion_submap_index[k]);
      out->insertion_result = std::move(ins);
    }
    return out;
  }
  void AddOdometryData(double /*time*/, const Rigid3d& /*pose*/) {}  // the fork never constructs its extrapolator (LTB:574-582)
  // Replaces the NDT initialisation when the caller knows the state (tests, re-localisation).
  void SetInitialState(const dl_nav_state& state) { ctx_->check(dl_ltb_set_initial_state(builder_, &state)); }
  int num_submaps() const { return dl_ltb_num_submaps(builder_); }
  dl_local_trajectory_builder* get() const { return builder_; }

 private:
  PointCloud Cloud(int which) const {
    int64_t n = 0;
    ctx_->check(dl_ltb_get_cloud(builder_, which, nullptr, 0, &n));
    PointCloud c((size_t)n);
    if (n) ctx_->check(dl_ltb_get_cloud(builder_, which, c[0].data(), n, &n));
    return c;
  }
  Context* ctx_;
  LocalTrajectoryBuilderOptions3D options_;
  RangeDataSynchronizer synchronizer_;
  dl_local_trajectory_builder* builder_ = nullptr;
};

}  // namespace mapping
}  // namespace dliom
