// Replays a recorded drive through the C++ mirror of mapping::LocalTrajectoryBuilder3D (dliom_b200.hpp), the way
// GlobalTrajectoryBuilder::AddSensorData feeds the reference's (global_trajectory_builder.cc:56-104).
// Input file (little endian), written by tests/test_gpu_ltb.py:
//   int32 num_events, then per event: int32 kind (0 = imu, 1 = range data of sensor A, 2 = of sensor B), double time,
//   imu: 3 doubles acc, 3 doubles gyr;  range: 3 floats origin, int32 n, n x 4 floats (x y z t).
// Output: one line per MatchingResult: time, pose (7), inserted, number of returns, number of submaps.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dliom_b200.hpp"

int main(int argc, char** argv) {
  using namespace dliom;
  if (argc < 2) return 3;
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  try {
    Context ctx(0);
    mapping::LocalTrajectoryBuilderOptions3D options;
    options.c.num_range_data = 5;
    options.c.motion_filter_max_time_seconds = 0.05;
    options.c.imu_weight = 0.7;
    if (argc > 2) options.c.frontend.voxel_filter_size = (float)std::atof(argv[2]);
    mapping::LocalTrajectoryBuilder3D builder(&ctx, options, {"lidar_a", "lidar_b"});
    dl_nav_state init{};
    int32_t num_events = 0;
    if (std::fread(&init, sizeof(init), 1, f) != 1 || std::fread(&num_events, 4, 1, f) != 1) return 3;
    builder.SetInitialState(init);
    for (int e = 0; e < num_events; ++e) {
      int32_t kind;
      double time;
      if (std::fread(&kind, 4, 1, f) != 1 || std::fread(&time, 8, 1, f) != 1) return 3;
      if (kind == 0) {
        sensor::ImuData imu{time, {}, {}};
        if (std::fread(imu.linear_acceleration.data(), 8, 3, f) != 3 || std::fread(imu.angular_velocity.data(), 8, 3, f) != 3) return 3;
        builder.AddImuData(imu);
        continue;
      }
      sensor::TimedPointCloudData cloud{time, {}, {}};
      int32_t n;
      if (std::fread(cloud.origin.data(), 4, 3, f) != 3 || std::fread(&n, 4, 1, f) != 1) return 3;
      cloud.ranges.resize(n);
      if (n && std::fread(cloud.ranges[0].data(), 16, n, f) != (size_t)n) return 3;
      const auto result = builder.AddRangeData(kind == 1 ? "lidar_a" : "lidar_b", cloud);
      if (!result) {
        std::printf("none %.6f\n", time);
        continue;
      }
      std::printf("result %.6f", result->time);
      for (int k = 0; k < 3; ++k) std::printf(" %.17g", result->local_pose.t[k]);
      for (int k = 0; k < 4; ++k) std::printf(" %.17g", result->local_pose.q[k]);
      std::printf(" %d %zu %d", result->insertion_result ? 1 : 0, result->range_data_in_local.returns.size(), builder.num_submaps());
      if (result->insertion_result)
        std::printf(" %zu %zu %zu", result->insertion_result->constant_data->high_resolution_point_cloud.size(),
                    result->insertion_result->constant_data->low_resolution_point_cloud.size(),
                    result->insertion_result->insertion_submaps.size());
      std::printf("\n");
    }
    std::fclose(f);
    return 0;
  } catch (const Error& e) {
    std::fprintf(stderr, "dliom error %d: %s\n", e.status, e.what());
    return 2;
  }
}
