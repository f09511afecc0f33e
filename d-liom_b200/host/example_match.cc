// Reads like SM/ceres_scan_matcher_3d_test.cc:34-116 and SM/real_time_correlative_scan_matcher_3d_test.cc:36-117,
// but runs the B200 path through the C++ shim. Build: g++ -std=c++17 example_match.cc -L.. -ldliom_b200
#include <cmath>
#include <cstdio>

#include "dliom_b200.hpp"

int main() {
  using namespace dliom;
  try {
    Context ctx(0);
    const PointCloud cloud = {{-3.f, 2.f, 0.f}, {-4.f, 2.f, 0.f}, {-5.f, 2.f, 0.f}, {-6.f, 2.f, 0.f},
                              {-6.f, 3.f, 1.f}, {-6.f, 4.f, 2.f}, {-7.f, 3.f, 1.f}};
    // hybrid_grid_.SetProbability(GetCellIndex(expected_pose * point), 1.) with expected_pose = Translation(-1,0,0);
    // ProbabilityToValue(1.) clamps to 0.9 -> 32767.
    DeviceHybridGrid grid(&ctx, 1.f);
    std::vector<int32_t> x, y, z;
    std::vector<uint16_t> v;
    for (const auto& p : cloud) {
      x.push_back((int32_t)std::lround(p[0] - 1.f)); y.push_back((int32_t)std::lround(p[1])); z.push_back((int32_t)std::lround(p[2]));
      v.push_back(32767);
    }
    grid.Update(x, y, z, v);
    scan_matching::CeresScanMatcherOptions3D options;
    options.occupied_space_weight = {1.};
    options.translation_weight = 0.01;
    options.rotation_weight = 0.1;
    options.use_nonmonotonic_steps = true;
    options.max_num_iterations = 10;
    scan_matching::CeresScanMatcher3D matcher(&ctx, options);
    Rigid3d initial, pose;
    initial.t[0] = -0.9; initial.t[1] = -0.2; initial.t[2] = 0.2;
    scan_matching::SolverSummary summary;
    matcher.Match({initial.t[0], initial.t[1], initial.t[2]}, initial, {{&cloud, &grid}}, &pose, &summary);
    std::printf("pose %.6f %.6f %.6f  final_cost %.6g  iterations %d\n", pose.t[0], pose.t[1], pose.t[2], summary.final_cost,
                summary.num_iterations);
    const bool ok = summary.final_cost <= 1e-2 && std::fabs(pose.t[0] + 1.) < 3e-2 && std::fabs(pose.t[1]) < 3e-2;
    const PointCloud filtered = sensor::VoxelFilter(&ctx, 2.5f).Filter(cloud);
    std::printf("voxel filter kept %zu of %zu\n", filtered.size(), cloud.size());
    return ok ? 0 : 1;
  } catch (const Error& e) {
    std::fprintf(stderr, "dliom error %d: %s\n", e.status, e.what());
    return 2;
  }
}
