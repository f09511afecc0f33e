// Scan ingest: deskew + transform + range gate between the two voxel-hash passes, and the small glue kernels of the
// batched front end (frame change back to tracking, cloud gathers, pose algebra, result records).
//
// Replaces the per-point loops of LocalTrajectoryBuilder3D::AddRangeData (LTB:426-472, InterpolatePose LTB:871-879,
// TransformRangeData LTB:485-487) and the pose bookkeeping of AddAccumulatedRangeData (LTB:502-505, :553-554).
// The reference builds a std::vector<Rigid3f> of N poses (double slerp per point) and then walks it; here one thread
// per surviving point computes its pose in registers (double slerp with the per-scan acos/sin hoisted to the host,
// composition with the previous state in double, narrowing to float exactly where the reference narrows) and
// applies it. Survivors are compacted in input order, because the voxel filter that follows keeps the FIRST point
// of every voxel. Algorithmic traffic: 16 B in + 12 B out per point.
#include <algorithm>

#include "dl_internal.cuh"
#include "dl_pipeline.cuh"

namespace dl {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ int block_exclusive_scan2(int value, int* total) {
  __shared__ int warp_sums[kBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = value;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int base = 0, sum = 0;
#pragma unroll
  for (int w = 0; w < kBlock / 32; ++w) {
    const int s = warp_sums[w];
    if (w < warp) base += s;
    sum += s;
  }
  *total = sum;
  return base + inc - value;
}

// Identity.slerp(s, rel.q), s * rel.t (Eigen QuaternionBase::slerp in double; theta / sin(theta) per scan).
__device__ __forceinline__ Rigidd interpolate_pose(double s, const ScanConstants& c) {
  double scale0, scale1;
  if (c.linear_slerp) {
    scale0 = 1.0 - s;
    scale1 = s;
  } else {
    scale0 = sin((1.0 - s) * c.theta) / c.sin_theta;
    scale1 = sin(s * c.theta) / c.sin_theta;
  }
  if (c.negative_dot) scale1 = -scale1;
  Rigidd out;
  out.q = {scale0 * 1.0 + scale1 * c.rel.q.w, scale0 * 0.0 + scale1 * c.rel.q.x, scale0 * 0.0 + scale1 * c.rel.q.y,
           scale0 * 0.0 + scale1 * c.rel.q.z};
  out.t = mul(s, c.rel.t);
  return out;
}

// Phase 1: classify every first-filter survivor (0 drop, 1 return, 2 miss), store its local-frame point, and count
// per 256-point tile.
__global__ void __launch_bounds__(kBlock) ingest_classify_kernel(IngestArgs a) {
  const int b = blockIdx.y;
  const int m = a.keep_counts[b];
  const int k = blockIdx.x * kBlock + threadIdx.x;
  const ScanConstants& sc = a.scans[b];
  const float* rows = a.ranges + (size_t)b * a.in_cap * 8;
  int cls = 0;
  if (k < m) {
    const int row = a.keep[(size_t)b * a.cap + k];
    const float* h = rows + (size_t)row * 8;
    // "no per-point time" switch (LTB:430-433) looks at the first survivor, which is always input row 0
    const bool no_deskew = (double)fabsf(rows[3]) < 1e-3;
    Rigidf pose;
    if (no_deskew) {
      pose = to_float(sc.cur);
    } else {
      const double s = (a.scan_period + (double)h[3]) / a.scan_period;
      pose = to_float(compose(sc.prev, interpolate_pose(s, sc)));
    }
    const unsigned long long origin_index = *(const unsigned long long*)(h + 4);
    const float* o = a.origins + 3 * origin_index;
    const Vec3f hit = apply(pose, Vec3f{h[0], h[1], h[2]});
    const Vec3f org = apply(pose, Vec3f{o[0], o[1], o[2]});
    const Vec3f delta = sub(hit, org);
    const float range = norm3(delta);
    Vec3f outp = hit;
    if (range >= a.min_range) {
      if (range <= a.max_range) {
        cls = 1;
      } else {
        cls = 2;
        outp = add(org, mul(a.max_range / range, delta));
      }
    }
    float* t = a.tmp_points + ((size_t)b * a.cap + k) * 3;
    t[0] = outp.x; t[1] = outp.y; t[2] = outp.z;
    a.cls[(size_t)b * a.cap + k] = (uint8_t)cls;
    if (k == m - 1) {  // hits_poses.back() (LTB:476)
      float* cp = a.current_pose + 7 * b;
      cp[0] = pose.t.x; cp[1] = pose.t.y; cp[2] = pose.t.z; cp[3] = pose.q.w; cp[4] = pose.q.x; cp[5] = pose.q.y; cp[6] = pose.q.z;
    }
  }
  int total_r, total_m;
  block_exclusive_scan2(cls == 1, &total_r);
  block_exclusive_scan2(cls == 2, &total_m);
  if (threadIdx.x == 0) {
    a.tile_counts[((size_t)b * a.tiles + blockIdx.x) * 2] = total_r;
    a.tile_counts[((size_t)b * a.tiles + blockIdx.x) * 2 + 1] = total_m;
  }
}

// Phase 2: order-preserving scatter into the returns / misses clouds.
__global__ void __launch_bounds__(kBlock) ingest_scatter_kernel(IngestArgs a) {
  const int b = blockIdx.y;
  const int m = a.keep_counts[b];
  __shared__ int base_r, base_m;
  int pr = 0, pm = 0;
  for (int t = threadIdx.x; t < (int)blockIdx.x; t += kBlock) {
    pr += a.tile_counts[((size_t)b * a.tiles + t) * 2];
    pm += a.tile_counts[((size_t)b * a.tiles + t) * 2 + 1];
  }
  int tr, tm;
  block_exclusive_scan2(pr, &tr);
  block_exclusive_scan2(pm, &tm);
  if (threadIdx.x == 0) {
    base_r = tr;
    base_m = tm;
  }
  __syncthreads();
  const int k = blockIdx.x * kBlock + threadIdx.x;
  const int cls = k < m ? a.cls[(size_t)b * a.cap + k] : 0;
  int total_r, total_m;
  const int off_r = block_exclusive_scan2(cls == 1, &total_r);
  const int off_m = block_exclusive_scan2(cls == 2, &total_m);
  if (cls) {
    const float* t = a.tmp_points + ((size_t)b * a.cap + k) * 3;
    float* dst = (cls == 1 ? a.returns_local + ((size_t)b * a.cap + base_r + off_r) * 3
                           : a.misses_local + ((size_t)b * a.cap + base_m + off_m) * 3);
    dst[0] = t[0]; dst[1] = t[1]; dst[2] = t[2];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    a.num_returns[b] = base_r + total_r;
    a.num_misses[b] = base_m + total_m;
  }
}

// out[b][j] = inverse(current_pose[b]) * in[b][keep[b][j]]   (TransformRangeData with current_pose.inverse(), LTB:485-487)
__global__ void __launch_bounds__(kBlock) gather_to_tracking_kernel(const float* __restrict__ in, int64_t cap,
                                                                    const int32_t* __restrict__ keep,
                                                                    const int32_t* __restrict__ keep_counts,
                                                                    const float* __restrict__ current_pose,
                                                                    float* __restrict__ out) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= keep_counts[b]) return;
  const float* cp = current_pose + 7 * b;
  const Rigidf back = inverse(Rigidf{{cp[0], cp[1], cp[2]}, {cp[3], cp[4], cp[5], cp[6]}});
  const float* p = in + ((size_t)b * cap + keep[(size_t)b * cap + j]) * 3;
  const Vec3f q = apply(back, Vec3f{p[0], p[1], p[2]});
  float* o = out + ((size_t)b * cap + j) * 3;
  o[0] = q.x; o[1] = q.y; o[2] = q.z;
}

// Plain gather of selected rows (adaptive filter survivors) into a dense cloud.
__global__ void __launch_bounds__(kBlock) gather_rows_kernel(const float* __restrict__ in, int64_t cap_in, int pairs_per_cloud,
                                                             const int32_t* __restrict__ keep, const int32_t* __restrict__ keep_counts,
                                                             int64_t cap_out, float* __restrict__ out) {
  const int pair = blockIdx.y;
  const int b = pair / pairs_per_cloud;
  const int count = keep_counts[pair];
  // grid-stride: the survivors are a few hundred rows, so a handful of CTAs per cloud replaces one per 256 rows of capacity
  for (int j = blockIdx.x * kBlock + threadIdx.x; j < count; j += gridDim.x * kBlock) {
    const float* p = in + ((size_t)b * cap_in + keep[(size_t)pair * cap_in + j]) * 3;
    float* o = out + ((size_t)pair * cap_out + j) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
  }
}

// initial_ceres_pose = submap.local_pose^-1 * pose_prediction, pose_prediction = current_pose.cast<double>() (LTB:476-487, :504-505)
__global__ void initial_pose_kernel(int batch, const float* __restrict__ current_pose, Rigidd submap_inverse,
                                    double* __restrict__ initial_pose, double* __restrict__ target_translation) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float* cp = current_pose + 7 * b;
  const Rigidd prediction = to_double(Rigidf{{cp[0], cp[1], cp[2]}, {cp[3], cp[4], cp[5], cp[6]}});
  const Rigidd init = compose(submap_inverse, prediction);
  pose_to7(init, initial_pose + 7 * b);
  target_translation[3 * b] = init.t.x;
  target_translation[3 * b + 1] = init.t.y;
  target_translation[3 * b + 2] = init.t.z;
}

__global__ void finalize_results_kernel(ResultArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch) return;
  dl_scan_result& r = a.results[b];
  const int n_hi = a.adaptive_counts[2 * b], n_lo = a.adaptive_counts[2 * b + 1];
  r.num_first_filter = a.first_counts[b];
  r.num_returns = a.return_counts[b];
  r.num_misses = a.miss_counts[b];
  r.num_high_resolution = n_hi;
  r.num_low_resolution = n_lo;
  r.num_cropped_high = a.adaptive_cropped[2 * b];
  r.num_cropped_low = a.adaptive_cropped[2 * b + 1];
  r.num_passes_high = a.adaptive_passes[2 * b];
  r.num_passes_low = a.adaptive_passes[2 * b + 1];
  r.rtcsm_score = a.rtcsm_scores ? a.rtcsm_scores[b] : 0.f;
  r.reserved = 0;
  // the reference drops the scan when any of the three clouds is empty (LTB:497-500, :510-513, :531-534)
  r.ok = (a.return_counts[b] > 0 && n_hi > 0 && n_lo > 0) ? 1 : 0;
  if (a.error_flag && a.error_flag[b]) r.ok = -1;  // a point fell outside +-2^20 voxels: results are not valid
  if (a.imu_ok && a.imu_ok[b] == 0) r.ok = -2;   // no IMU factor (no samples / covariance not positive definite): no solve ran
  const double* pose = a.fused ? a.fused[b].state : a.nls[b].pose;
  r.summary = a.fused ? a.fused[b].summary : a.nls[b].summary;
  for (int i = 0; i < 7; ++i) r.pose_observation_in_submap[i] = pose[i];
  const Rigidd est = compose(a.submap, pose_from7(pose));  // LTB:553-554
  pose_to7(est, r.pose_estimate_local);
  if (a.fused && a.states_out) {  // solver state (submap frame) -> dl_nav_state in the local frame
    dl_nav_state& o = a.states_out[b];
    const Vec3d v = rotate(a.submap.q, Vec3d{pose[7], pose[8], pose[9]});
    o.p[0] = est.t.x; o.p[1] = est.t.y; o.p[2] = est.t.z;
    o.q[0] = est.q.w; o.q[1] = est.q.x; o.q[2] = est.q.y; o.q[3] = est.q.z;
    o.v[0] = v.x; o.v[1] = v.y; o.v[2] = v.z;
    for (int k = 0; k < 3; ++k) { o.ba[k] = pose[10 + k]; o.bg[k] = pose[13 + k]; }
  }
}

}  // namespace

int launch_ingest(dl_context* ctx, const IngestArgs& a, int batch) {
  if (batch <= 0) return DL_OK;
  const dim3 grid(a.tiles, batch);
  ingest_classify_kernel<<<grid, kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "ingest_classify_kernel");
  ingest_scatter_kernel<<<grid, kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "ingest_scatter_kernel");
  return DL_OK;
}

int launch_gather_to_tracking(dl_context* ctx, const float* in, int64_t cap, const int32_t* keep,
                              const int32_t* keep_counts, const float* current_pose, float* out, int batch) {
  if (batch <= 0) return DL_OK;
  const dim3 grid((unsigned)((cap + kBlock - 1) / kBlock), batch);
  gather_to_tracking_kernel<<<grid, kBlock, 0, ctx->stream>>>(in, cap, keep, keep_counts, current_pose, out);
  DL_LAUNCH_CHECK(ctx, "gather_to_tracking_kernel");
  return DL_OK;
}

int launch_gather_rows(dl_context* ctx, const float* in, int64_t cap_in, int pairs_per_cloud, const int32_t* keep,
                       const int32_t* keep_counts, int64_t cap_out, float* out, int pairs) {
  if (pairs <= 0) return DL_OK;
  const dim3 grid((unsigned)std::min<int64_t>((cap_out + kBlock - 1) / kBlock, 8), pairs);
  gather_rows_kernel<<<grid, kBlock, 0, ctx->stream>>>(in, cap_in, pairs_per_cloud, keep, keep_counts, cap_out, out);
  DL_LAUNCH_CHECK(ctx, "gather_rows_kernel");
  return DL_OK;
}

int launch_initial_pose(dl_context* ctx, int batch, const float* current_pose, const Rigidd& submap_inverse,
                        double* initial_pose, double* target_translation) {
  if (batch <= 0) return DL_OK;
  initial_pose_kernel<<<(batch + 127) / 128, 128, 0, ctx->stream>>>(batch, current_pose, submap_inverse, initial_pose,
                                                                    target_translation);
  DL_LAUNCH_CHECK(ctx, "initial_pose_kernel");
  return DL_OK;
}

int launch_finalize_results(dl_context* ctx, const ResultArgs& a) {
  if (a.batch <= 0) return DL_OK;
  finalize_results_kernel<<<(a.batch + 127) / 128, 128, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "finalize_results_kernel");
  return DL_OK;
}

}  // namespace dl
