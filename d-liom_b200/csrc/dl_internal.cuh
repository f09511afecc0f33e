// Internal declarations shared by the translation units of libdliom_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dliom_b200.h"
#include "dl_math.cuh"

namespace dl {

constexpr int kNumSMs = 148;  // B200

// ------------------------------------------------------------------------------------------------ device grid
// Index-array form of HybridGrid's three levels (hybrid_grid.h:411-412): top cell = 64^3 voxels
// (DynamicGrid meta cell), node = 8^3 bricks (NestedGrid), brick = 8^3 uint16 voxels, z-major (FlatGrid).
// A brick is 1 KiB and 1 KiB-aligned in HBM, so a whole brick is one bulk-copy unit.
struct GridView {
  const int32_t* __restrict__ top;    // (1 << bits)^3 entries -> node index, -1 = absent
  const int32_t* __restrict__ nodes;  // num_nodes * 512 entries -> brick index, -1 = absent
  const uint16_t* __restrict__ bricks;  // num_bricks * 512 voxels
  float resolution;
  int bits;
};

// HybridGrid::value(): origin shift by grid_size/2, unsigned bounds test, three dependent reads.
__device__ __forceinline__ uint16_t grid_value(const GridView& g, int x, int y, int z) {
  const int gs = 64 << g.bits;
  const int half = gs >> 1;
  const unsigned sx = (unsigned)(x + half), sy = (unsigned)(y + half), sz = (unsigned)(z + half);
  if (sx >= (unsigned)gs || sy >= (unsigned)gs || sz >= (unsigned)gs) return 0;
  const int node = __ldg(g.top + ((((sz >> 6) << g.bits) + (sy >> 6)) << g.bits) + (sx >> 6));
  if (node < 0) return 0;
  const int brick = __ldg(g.nodes + (size_t)node * 512 + ((((sz >> 3) & 7) << 6) | (((sy >> 3) & 7) << 3) | ((sx >> 3) & 7)));
  if (brick < 0) return 0;
  return __ldg(g.bricks + (size_t)brick * 512 + (((sz & 7) << 6) | ((sy & 7) << 3) | (sx & 7)));
}

}  // namespace dl

// ------------------------------------------------------------------------------------------------ handles
struct dl_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // uploads of host scans, overlapped with the kernels of the previous sub-batch
  cudaStream_t aux_stream = nullptr;    // odd sub-batches of the front end (see frontend_run)
  cudaStream_t tail_stream = nullptr;   // high priority: the latency-bound back half (adaptive filter, LM solve) of every sub-batch
  cudaEvent_t staging_done = nullptr;   // the pinned staging block of the previous call has been consumed
  cudaEvent_t batch_done = nullptr;     // dl_frontend_submit: everything of the batch in flight, incl. the result download
  // adaptive voxel filter: how many (cloud, filter) pairs of the last probed launch needed the single-CTA search
  // ([0] pairs that fell through, [1] pairs probed); pinned host copy of a device counter, read without synchronising
  int32_t* h_adaptive_stats = nullptr;
  int32_t* d_adaptive_stats = nullptr;
  int64_t adaptive_calls = 0;
  uint8_t* d_fcsm_lut = nullptr;        // loop-closure search: cell value -> 8-bit precomputation value (dl_fcsm.cu), built on first use
  int in_flight = 0;                    // scans of the submitted, not yet collected batch
  bool in_flight_states = false;        // ... and whether it also stages the estimated IMU states
  size_t results_staging_offset = 0;    // where in h_pinned the in-flight batch's results land
  std::string error;
  int64_t launches = 0;
  // growable scratch arenas (device + pinned host), reused across calls
  void* d_scratch = nullptr;
  size_t d_scratch_bytes = 0;
  void* h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  // optional per-stage timing (events recorded on `stream`)
  bool profiling = false;
  struct Mark {
    int stage;
    cudaEvent_t begin, end;
  };
  std::vector<Mark> marks;
  std::vector<cudaEvent_t> event_pool;
  std::vector<std::string> stage_names;
  std::vector<double> stage_ms;
  std::vector<int64_t> stage_calls;
  int stage_id(const char* name);
  cudaEvent_t take_event();

  int fail(int status, const std::string& msg) {
    error = msg;
    return status;
  }
  int cuda_fail(cudaError_t e, const char* what) {
    error = std::string(what) + ": " + cudaGetErrorString(e);
    return DL_ERR_CUDA;
  }
  int reserve_device(size_t bytes);
  int reserve_pinned(size_t bytes);
  // Host wait for everything on `stream`. blocking_sync: the thread sleeps on a cudaEventBlockingSync event instead of
  // spinning in cudaStreamSynchronize — for background threads (loop closure, pose graph) on hosts with fewer CPUs than threads.
  bool blocking_sync = false;
  cudaEvent_t sync_event = nullptr;
  cudaError_t wait_stream() {
    if (!blocking_sync) return cudaStreamSynchronize(stream);
    if (!sync_event) {
      const cudaError_t e = cudaEventCreateWithFlags(&sync_event, cudaEventBlockingSync | cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    const cudaError_t e = cudaEventRecord(sync_event, stream);
    return e != cudaSuccess ? e : cudaEventSynchronize(sync_event);
  }
};

struct dl_grid {
  dl_context* ctx = nullptr;
  float resolution = 0.f;
  int bits = 1;
  // host mirror of the three levels
  std::vector<int32_t> top;                 // (1<<bits)^3
  std::vector<int32_t> nodes;               // num_nodes * 512
  std::vector<uint16_t> bricks;             // num_bricks * 512
  std::vector<uint8_t> brick_dirty;         // per brick
  bool structure_dirty = true;
  // device copies
  int32_t* d_top = nullptr;
  int32_t* d_nodes = nullptr;
  uint16_t* d_bricks = nullptr;
  size_t d_top_cap = 0, d_nodes_cap = 0, d_bricks_cap = 0;  // element capacities
  int32_t* d_counters = nullptr;  // [0] nodes in use, [1] bricks in use, [2] scratch (update-list length)
  bool mirror_stale = false;      // the device copy was modified by dl_grid_insert_range_data: the host mirror is behind
  uint64_t version = 1;           // bumped by every modification (cells set, sync, device insert)
  // loop-closure search index (dl_fcsm.cu), built on first use and rebuilt when `version` moved on
  uint8_t* d_m8 = nullptr;
  size_t d_m8_bytes = 0;
  uint64_t m8_version = 0;        // 0 = none
  int m8_org[3] = {0, 0, 0}, m8_dim[3] = {0, 0, 0};
  std::mutex index_mutex;
  dl::GridView view() const { return {d_top, d_nodes, d_bricks, resolution, bits}; }
};

#define DL_CUDA(ctx, call)                                  \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call); \
  } while (0)

#define DL_TRY_STATUS(expr)         \
  do {                              \
    const int st__ = (expr);        \
    if (st__ != DL_OK) return st__; \
  } while (0)

#define DL_LAUNCH_CHECK(ctx, name)                          \
  do {                                                      \
    (ctx)->launches++;                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, name); \
  } while (0)

namespace dl {

// RAII stage bracket: records begin/end events on the context stream when profiling is on.
struct StageScope {
  dl_context* ctx;
  int mark = -1;
  StageScope(dl_context* c, const char* name) : ctx(c) {
    if (!c->profiling) return;
    dl_context::Mark m{c->stage_id(name), c->take_event(), c->take_event()};
    cudaEventRecord(m.begin, c->stream);
    mark = (int)c->marks.size();
    c->marks.push_back(m);
  }
  ~StageScope() {
    if (mark >= 0) cudaEventRecord(ctx->marks[mark].end, ctx->stream);
  }
};

struct Arena {  // bump allocator over the context's device scratch
  char* base;
  size_t off = 0;
  explicit Arena(void* p) : base((char*)p) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = (T*)(base + off);
    off += count * sizeof(T);
    return p;
  }
};
inline size_t arena_bytes(std::initializer_list<size_t> sizes) {
  size_t t = 0;
  for (size_t s : sizes) t = ((t + 255) & ~size_t(255)) + s;
  return t + 256;
}

// ---- kernel launchers (each defined in its own .cu), all asynchronous on `stream`, device pointers only.

// Voxel filter over `batch` clouds. Cloud b = rows [b * cap, b * cap + counts[b]) of `points` (stride floats per row).
// keep[b * cap ...] receives the surviving row indices (relative to the cloud) in input order, keep_counts[b] their number.
// table: batch * table_cap uint32 (table_cap a power of two >= 2 * max count); slot: batch * cap uint32.
int launch_voxel_filter(dl_context* ctx, const float* points, int stride, int64_t cap, const int32_t* counts, int batch,
                        float resolution, uint32_t* table, int64_t table_cap, uint32_t* slot, int32_t* keep,
                        int32_t* keep_counts, int32_t* block_counts);
int launch_voxel_indices(dl_context* ctx, const float* points, int stride, int64_t n, float resolution, int32_t* out);

struct AdaptiveParams {
  float max_length, min_num_points, max_range;
};
// One CTA per (cloud, filter) pair: adaptive bisection entirely on the device.
// filters: num_filters parameter blocks; outputs indexed [(b * num_filters + f) * cap ...].
int launch_adaptive_voxel_filter(dl_context* ctx, const float* points, int stride, int64_t cap, const int32_t* counts,
                                 int batch, const AdaptiveParams* filters_dev, int num_filters, uint32_t* table,
                                 int64_t table_cap, uint32_t* scratch /* pairs * 2 * cap */, int32_t* keep, int32_t* keep_counts,
                                 float* passes /* (batch*num_filters) * 32 */, int32_t* num_passes,
                                 int32_t* cropped_counts /* optional, batch*num_filters */,
                                 void* first_pass_scratch /* optional, adaptive_first_pass_bytes(batch * num_filters, cap) */);
size_t adaptive_first_pass_bytes(int pairs, int64_t cap);

struct RtcsmScan {  // one scan of a batched correlative search (dl_rtcsm.cu), device pointers
  const float* points;  // n x 3
  int32_t n;
  const Quatf* cand_q;  // R rotations (composed with the initial pose, normalised)
  const Vec3f* cand_t;  // L translations (composed with the initial pose)
  const double* pen_r;  // R: angle * rotation_delta_cost_weight
  const double* pen_t;  // L: |t| * translation_delta_cost_weight
  int32_t R, L;
  float* scores;                     // optional, R * L
  unsigned long long* best;          // (score bits << 32) | ~index, zeroed before the launch
};
int rtcsm_ctas_for(int64_t R, int64_t L);
int launch_rtcsm_batch(dl_context* ctx, const GridView& grid, const RtcsmScan* scans_dev, const int32_t* cta_prefix_dev, int num_scans,
                       int total_ctas);
int launch_rtcsm_pick(dl_context* ctx, const RtcsmScan* scans_dev, int num_scans, double* pose_out, const int32_t* pose_slot,
                      float* score_out, const int32_t* score_slot);
int launch_max_range_batch(dl_context* ctx, const float* points, int64_t stride_floats, const int32_t* counts, int count_stride,
                           int batch, float init, float* out);

// ---- NLS
struct NlsProblem {  // one scan-to-submap registration problem, device pointers
  const float* cloud[DL_MAX_PAIRS];
  int32_t count[DL_MAX_PAIRS];
  const int32_t* count_dev[DL_MAX_PAIRS];  // optional: read the count from device memory instead
  GridView grid[DL_MAX_PAIRS];
  double target_t[3];
  double initial[7];
  const double* initial_dev;  // optional: 7 doubles on the device override `initial` (and target_t = its translation
                              // unless target_dev is set)
  const double* target_dev;
  const int32_t* enabled_dev;  // optional: the problem is skipped (output untouched) when *enabled_dev == 0
};
struct NlsOptions {
  int num_pairs;
  double occ_weight[DL_MAX_PAIRS];
  double trans_weight, rot_weight;
  int only_yaw, nonmono, max_iter;
  int cluster;  // CTAs per problem (thread-block cluster; 0 / 1 = one CTA): > 1 for clouds of tens of thousands of points, see dl_nls.cu
};
struct NlsOutput {
  double pose[7];
  dl_solve_summary summary;
};
int launch_nls(dl_context* ctx, const NlsOptions& opt, const NlsProblem* problems_dev, int count, NlsOutput* out_dev);
struct FusedOutput {
  double state[16];  // p(3) q(4 wxyz) v(3) ba(3) bg(3)
  dl_solve_summary summary;
};
// IMU term of the fused solve in the SUBMAP frame (built by dl_api.cu on the host or by imu_prepare_kernel on the device):
// state i (fixed), the pre-integrated deltas, gravity, and W = weight^2 * Sigma^-1 (row-major 15x15, order p, theta, v, ba, bg).
struct ImuTerm {
  double pi[3], qi[4], vi[3], bai[3], bgi[3];
  double dp[3], dq[4], dv[3];
  double G[3];
  double sum_dt;
  double W[225];
};
constexpr int kImuTermDoubles = 16 + 10 + 3 + 1 + 225;
static_assert(sizeof(ImuTerm) == kImuTermDoubles * sizeof(double), "ImuTerm layout");
int launch_nls_fused(dl_context* ctx, const NlsOptions& opt, const NlsProblem* problems_dev, const ImuTerm* imu_terms_dev,
                     const double* initial16_dev, int count, FusedOutput* out_dev);
int launch_nls_normal_equations(dl_context* ctx, const NlsOptions& opt, const NlsProblem* problems_dev,
                                const double* at_pose_dev, double* out28_dev);
int launch_imu_preintegrate(dl_context* ctx, int count, const int32_t* offsets, const double* dts, const double* accs,
                            const double* gyrs, const double* biases, int bias_stride, const dl_imu_noise& noise,
                            dl_preintegration* out);
struct DecodeArgs {  // one sensor_msgs/PointCloud2 message (dl_decode.cu)
  const uint8_t* data;
  int64_t n;
  int point_step, offset_x, offset_y, offset_z, offset_time, time_type;
  bool xyz_aligned, time_aligned;  // fields readable with aligned loads
  Rigidf sensor_to_tracking;
  int32_t* tile_counts;  // ceil(n / 256)
  float* rows_out;
  int32_t* num_out;
  double* stamp_offset;
};
int launch_decode_point_cloud2(dl_context* ctx, const DecodeArgs& a);
struct FcsmPair {  // one (node, submap) loop-closure search, device pointers
  GridView hi, lo;
  const float* hi_pts;
  const float* lo_pts;
  int* cells;      // 3 * n_hi: full-resolution cell of every high-resolution point at the guess
  float* lo_rot;   // 3 * n_lo: rotated low-resolution points
  int n_hi, n_lo;
  Rigidf pose;     // float cast of the pose guess (cc:171-173)
  int wxy, wz;     // window half-widths in cells
  float min_score;
  double min_low;
  // search index of the high-resolution grid (dense sliding 8^3 maximum of the 8-bit values); null = exhaustive search
  const uint8_t* m8;
  int m8_org[3], m8_dim[3];
};
struct FcsmPick {
  int found;
  float score, low_resolution_score;
  int offset[3];
  long long num_candidates;
  double pose[7];  // coarse pose (the guess itself when nothing was found)
};
constexpr int kFcsmRun = 8;  // x offsets per search thread (dl_fcsm.cu)
int launch_fcsm_index(dl_context* ctx, const GridView& g, int ox, int oy, int oz, int nx, int ny, int nz, uint8_t* tmp, uint8_t* out);
int launch_fcsm_pruned(dl_context* ctx, const FcsmPair* pairs_dev, int count, int max_points, int max_blocks, int* bounds_dev,
                       int* max_bound_dev, unsigned long long* best_dev, FcsmPick* picks_dev);
int launch_fcsm(dl_context* ctx, const FcsmPair* pairs_dev, int count, int max_points, long long max_threads,
                unsigned long long* best_dev, FcsmPick* picks_dev, float* all_scores_dev);
void compute_odds_table(float probability, uint16_t* table32768);
int grid_insert_device(dl_context* ctx, dl_grid* g, const Vec3f& origin, const float* d_returns, int n, int num_free,
                       const uint16_t* d_hit_table, const uint16_t* d_miss_table, int32_t* d_bbox, uint32_t* d_update_list);
int launch_transform_filter(dl_context* ctx, const float* in, int n, const Rigidf& to_submap, const Vec3f& origin_submap,
                            float max_range, float* all, float* near, int32_t* near_count, int32_t* tile_counts);
// dl_window.cu
int launch_window_optimize(dl_context* ctx, int count, const dl_nav_state* states_i, const double* prior_information,
                           const dl_preintegration* preint, const double* matched_pose, const dl_nav_state* initial_j,
                           const dl_window_options& opt, dl_nav_state* states_i_out, dl_nav_state* states_j_out,
                           double* information_out, dl_solve_summary* summaries);
// dl_histogram.cu
size_t rotational_histogram_scratch_bytes(int64_t n);
int launch_rotational_histogram(dl_context* ctx, Arena& a, const float* d_points, int64_t n, int size, float* d_histogram,
                                int32_t** d_error_out);
// dl_comm.cu: staging buffers of the constraint exchange and the timed all-gather
int comm_reserve(dl_comm* c, size_t bytes_per_rank);
void* comm_send_buffer(dl_comm* c);
void* comm_recv_buffer(dl_comm* c);
int comm_all_gather(dl_comm* c, const void* send_dev, void* recv_dev, size_t bytes, float* ms);
// dl_fcsm.cu: one dl_constraint_row per searched pair from the coarse picks and the refinement's output
int launch_pack_constraint_rows(dl_context* ctx, int n, const FcsmPick* picks, const NlsOutput* refined, const int32_t* submap_ids,
                                const int32_t* node_ids, double translation_weight, double rotation_weight, int rank,
                                dl_constraint_row* rows);
int launch_interpolate(dl_context* ctx, const GridView& grid, int64_t n, const double* xyz, double* out);
int launch_grid_lookup(dl_context* ctx, const GridView& grid, int64_t n, const int32_t* xyz, uint16_t* out);

}  // namespace dl
