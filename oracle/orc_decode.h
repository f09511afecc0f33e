// TEST INFRASTRUCTURE — CPU oracle (see orc_math.h header).
//
// Wire format -> TimedPointCloud in the tracking frame: the per-sensor-type loops of
// SensorBridge::HandlePointCloud2Message (cartographer_ros/sensor_bridge.cc:176-240; point structs sensor_bridge.h:54-95,
// NaN/Inf test :100-107) followed by HandleRangefinder's TransformTimedPointCloud (sensor_bridge.cc:286-300,
// cartographer/sensor/point_cloud.cc:35-46). pcl::fromROSMsg is a field-wise copy, so the loops are restated directly on
// the message bytes: a point is `point_step` bytes, x / y / z float32 at their field offsets, the time field by type.
// The reference has no test for this function: parity here is oracle <-> device only ("unpinned" row).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orc_math.h"

namespace orc {

enum TimeType { kTimeNone = 0, kTimeFloat32Seconds = 1, kTimeUint32Nanoseconds = 2, kTimeFloat64Seconds = 3 };

struct PointCloud2Layout {
  int point_step, offset_x, offset_y, offset_z, offset_time, time_type;
};

template <typename T>
inline T read_at(const uint8_t* p) {
  T v;
  std::memcpy(&v, p, sizeof(T));
  return v;
}

// rows_out: 4 floats per kept point (x y z in the tracking frame, t relative to the last point). Returns the number kept;
// *stamp_offset = seconds to add to the message stamp (velodyne / ouster: the stamp is the FIRST point's time).
inline int64_t decode_point_cloud2(const PointCloud2Layout& l, const uint8_t* data, int64_t n, const Rigid3d& sensor_to_tracking,
                                   float* rows_out, double* stamp_offset) {
  const Rigid3f T = cast_f(sensor_to_tracking);  // sensor_to_tracking->cast<float>()
  double rel_time_last = 0.;                     // the reference's local is a double in every branch
  if (n > 0) {
    const uint8_t* last = data + (n - 1) * l.point_step + l.offset_time;
    if (l.time_type == kTimeFloat32Seconds) rel_time_last = read_at<float>(last);
    if (l.time_type == kTimeUint32Nanoseconds) rel_time_last = (float)read_at<uint32_t>(last) * 1e-9f;  // uint32 * float -> float
    if (l.time_type == kTimeFloat64Seconds) rel_time_last = read_at<double>(last);
  }
  *stamp_offset = (l.time_type == kTimeFloat32Seconds || l.time_type == kTimeUint32Nanoseconds) ? rel_time_last : 0.;
  int64_t kept = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* p = data + i * l.point_step;
    const float x = read_at<float>(p + l.offset_x), y = read_at<float>(p + l.offset_y), z = read_at<float>(p + l.offset_z);
    if (std::isnan(x) || std::isnan(y) || std::isnan(z) || std::isinf(x) || std::isinf(y) || std::isinf(z)) continue;
    float t = 0.f;  // Eigen::Vector4f(x, y, z, <double expression>) narrows the 4th argument to float
    if (l.time_type == kTimeFloat32Seconds) t = (float)((double)read_at<float>(p + l.offset_time) - rel_time_last);
    if (l.time_type == kTimeUint32Nanoseconds) t = (float)((double)((float)read_at<uint32_t>(p + l.offset_time) * 1e-9f) - rel_time_last);
    if (l.time_type == kTimeFloat64Seconds) t = (float)(read_at<double>(p + l.offset_time) - rel_time_last);
    const V3f q = apply(T, V3f{x, y, z});
    float* o = rows_out + 4 * kept++;
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = t;
  }
  return kept;
}

}  // namespace orc
