// Voxel hashing kernels: first-point-per-voxel filter and its adaptive (bisection) variant.
//
// Replaces sensor::VoxelFilter::Filter / AdaptiveVoxelFilter::Filter
// (C/sensor/internal/voxel_filter.cc:28-131,147-150). The reference inserts 96-bit keys into a
// std::unordered_set in input order and keeps a point iff its insertion succeeded. Here every point
// atomically proposes its input index to an open-addressing table slot keyed by the voxel; the slot keeps
// the MINIMUM index (atomicMin), so the survivor of each voxel is the first point in input order regardless
// of thread scheduling, and an order-preserving compaction returns exactly the reference's output.
// The table stores only point indices (4 B/slot): a slot's key is the voxel of whichever point currently owns
// it, which is invariant under atomicMin among points of the same voxel.
//
// HBM traffic per pass (algorithmic): read stride*4 B per point, write 4 B per survivor. The table (8 B per
// point) lives in L2. Compile with -fmad=false: index = lroundf(x / resolution) must match the CPU bit for bit.
#include "dl_internal.cuh"

namespace dl {
namespace {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kBlock = 256;

__device__ __forceinline__ uint32_t hash_cell(const Int3& c) {
  uint32_t h = (uint32_t)c.x * 73856093u ^ (uint32_t)c.y * 19349663u ^ (uint32_t)c.z * 83492791u;
  h ^= h >> 15;
  h *= 0x2c1b3c6du;
  h ^= h >> 12;
  return h;
}
__device__ __forceinline__ Int3 row_cell(const float* __restrict__ pts, int stride, uint32_t row, float res) {
  const float* p = pts + (size_t)row * stride;
  return cell_index(Vec3f{p[0], p[1], p[2]}, res);
}

// Proposes `id` (position in the filter's input order) for the voxel `c`. `row_of(id)` maps ids to rows.
template <typename RowOf>
__device__ __forceinline__ uint32_t table_insert(uint32_t* table, uint32_t mask, const float* pts, int stride,
                                                 float res, const Int3& c, uint32_t id, RowOf row_of) {
  uint32_t h = hash_cell(c) & mask;
  for (;;) {
    const uint32_t prev = atomicCAS(table + h, kEmpty, id);
    if (prev == kEmpty) return h;
    const Int3 o = row_cell(pts, stride, row_of(prev), res);
    if (o.x == c.x && o.y == c.y && o.z == c.z) {
      atomicMin(table + h, id);
      return h;
    }
    h = (h + 1) & mask;
  }
}

// ------------------------------------------------------------------------------------------- plain filter
__global__ void __launch_bounds__(kBlock) voxel_insert_kernel(const float* __restrict__ points, int stride,
                                                              int64_t cap, const int32_t* __restrict__ counts,
                                                              float res, uint32_t* table, int64_t table_cap,
                                                              uint32_t* slot) {
  const int b = blockIdx.y;
  const int n = counts[b];
  const float* pts = points + (size_t)b * cap * stride;
  uint32_t* tab = table + (size_t)b * table_cap;
  const uint32_t mask = (uint32_t)table_cap - 1;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const Int3 c = row_cell(pts, stride, i, res);
    slot[(size_t)b * cap + i] = table_insert(tab, mask, pts, stride, res, c, (uint32_t)i, [](uint32_t id) { return id; });
  }
}

__device__ __forceinline__ int block_exclusive_scan(int value, int* total) {
  __shared__ int warp_sums[kBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = value;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();  // protects warp_sums across successive calls
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

// Per 256-row tile: number of survivors (a row survives iff it owns its slot).
__global__ void __launch_bounds__(kBlock) voxel_count_kernel(const int32_t* __restrict__ counts, int64_t cap,
                                                             const uint32_t* __restrict__ table, int64_t table_cap,
                                                             const uint32_t* __restrict__ slot, int32_t* block_counts,
                                                             int tiles) {
  const int b = blockIdx.y;
  const int n = counts[b];
  const int i = blockIdx.x * kBlock + threadIdx.x;
  int flag = 0;
  if (i < n) flag = __ldcg(table + (size_t)b * table_cap + slot[(size_t)b * cap + i]) == (uint32_t)i;
  int total;
  block_exclusive_scan(flag, &total);
  if (threadIdx.x == 0) block_counts[(size_t)b * tiles + blockIdx.x] = total;
}

__global__ void __launch_bounds__(kBlock) voxel_scatter_kernel(const int32_t* __restrict__ counts, int64_t cap,
                                                               const uint32_t* __restrict__ table, int64_t table_cap,
                                                               const uint32_t* __restrict__ slot,
                                                               const int32_t* __restrict__ block_counts, int tiles,
                                                               int32_t* keep, int32_t* keep_counts) {
  const int b = blockIdx.y;
  const int n = counts[b];
  __shared__ int tile_base;
  // prefix over the preceding tiles of this cloud (<= 1024 tiles for 262 144 rows)
  int partial = 0;
  for (int t = threadIdx.x; t < (int)blockIdx.x; t += kBlock) partial += block_counts[(size_t)b * tiles + t];
  int total;
  block_exclusive_scan(partial, &total);
  if (threadIdx.x == 0) tile_base = total;
  __syncthreads();
  const int i = blockIdx.x * kBlock + threadIdx.x;
  int flag = 0;
  if (i < n) flag = __ldcg(table + (size_t)b * table_cap + slot[(size_t)b * cap + i]) == (uint32_t)i;
  const int off = block_exclusive_scan(flag, &total);
  if (flag) keep[(size_t)b * cap + tile_base + off] = i;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) keep_counts[b] = tile_base + total;
}

__global__ void voxel_indices_kernel(const float* __restrict__ points, int stride, int64_t n, float res,
                                     int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Int3 c = row_cell(points, stride, (uint32_t)i, res);
  out[3 * i] = c.x;
  out[3 * i + 1] = c.y;
  out[3 * i + 2] = c.z;
}

// ------------------------------------------------------------------------------------------- adaptive filter
// One CTA runs the whole data-dependent pass sequence of AdaptivelyVoxelFiltered for one (cloud, filter) pair,
// so the bisection needs no host round trip: every pass clears the table, re-inserts the range-cropped cloud and
// block-reduces the survivor count; the control flow below is the reference's, statement for statement.
constexpr int kAdaptiveBlock = 1024;

__device__ __forceinline__ int block_sum_1024(int v) {
  __shared__ int ws[32];
  __shared__ int result;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if (lane == 0) ws[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int s = ws[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) result = s;
  }
  __syncthreads();
  return result;
}

// Order-preserving compaction of ids [0, n) with predicate flags computed by `pred`; returns the count.
template <typename Pred, typename Emit>
__device__ __forceinline__ int block_compact_1024(int n, Pred pred, Emit emit) {
  __shared__ int ws[32];
  __shared__ int running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < n; base += kAdaptiveBlock) {
    const int i = base + threadIdx.x;
    const int flag = (i < n) ? (pred(i) ? 1 : 0) : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    const int in_warp = __popc(ballot & ((1u << lane) - 1));
    if (lane == 0) ws[warp] = __popc(ballot);
    __syncthreads();
    int warp_base = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
      const int s = ws[w];
      if (w < warp) warp_base += s;
      tile_total += s;
    }
    const int start = running;
    if (flag) emit(start + warp_base + in_warp, i);
    __syncthreads();
    if (threadIdx.x == 0) running = start + tile_total;
    __syncthreads();
  }
  return running;
}

__global__ void __launch_bounds__(kAdaptiveBlock) adaptive_voxel_kernel(
    const float* __restrict__ points, int stride, int64_t cap, const int32_t* __restrict__ counts,
    const AdaptiveParams* __restrict__ filters, int num_filters, uint32_t* table, int64_t table_cap,
    uint32_t* scratch /* per pair: cap cropped rows + cap slots */, int32_t* keep, int32_t* keep_counts,
    float* passes, int32_t* num_passes, int32_t* cropped_counts) {
  const int pair = blockIdx.x;
  const int b = pair / num_filters;
  const AdaptiveParams opt = filters[pair % num_filters];
  const int n = counts[b];
  const float* pts = points + (size_t)b * cap * stride;
  uint32_t* tab = table + (size_t)pair * table_cap;
  uint32_t* rows = scratch + (size_t)pair * 2 * cap;  // cropped cloud: id -> row
  uint32_t* slot = rows + cap;
  int32_t* out = keep + (size_t)pair * cap;
  float* pass_log = passes + (size_t)pair * 32;
  int npass = 0;

  // FilterByMaxRange (voxel_filter.cc:28-38): norm = sqrt(x^2 + (y^2 + z^2)) <= max_range
  const int m = block_compact_1024(
      n,
      [&](int i) {
        const float* p = pts + (size_t)i * stride;
        return norm3(Vec3f{p[0], p[1], p[2]}) <= opt.max_range;
      },
      [&](int pos, int i) { rows[pos] = (uint32_t)i; });

  if (threadIdx.x == 0 && cropped_counts) cropped_counts[pair] = m;
  auto finish_all = [&]() {  // 'point_cloud' is already sparse enough
    for (int j = threadIdx.x; j < m; j += kAdaptiveBlock) out[j] = (int32_t)rows[j];
    if (threadIdx.x == 0) {
      keep_counts[pair] = m;
      num_passes[pair] = npass;
    }
  };
  if ((float)m <= opt.min_num_points) {
    finish_all();
    return;
  }

  // table sized to the cropped cloud
  uint32_t eff_cap = 64;
  while (eff_cap < 2u * (uint32_t)m) eff_cap <<= 1;
  if (eff_cap > (uint32_t)table_cap) eff_cap = (uint32_t)table_cap;
  const uint32_t mask = eff_cap - 1;

  float last_edge = -1.f;
  auto run_pass = [&](float edge, bool log) -> int {
    if (log) {
      if (threadIdx.x == 0 && npass < 32) pass_log[npass] = edge;
      ++npass;
    }
    for (uint32_t i = threadIdx.x; i < eff_cap; i += kAdaptiveBlock) tab[i] = kEmpty;
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += kAdaptiveBlock) {
      const Int3 c = row_cell(pts, stride, rows[j], edge);
      slot[j] = table_insert(tab, mask, pts, stride, edge, c, (uint32_t)j, [&](uint32_t id) { return rows[id]; });
    }
    __syncthreads();
    int local = 0;
    for (int j = threadIdx.x; j < m; j += kAdaptiveBlock) local += __ldcg(tab + slot[j]) == (uint32_t)j;  // L2 read: atomics bypass L1
    last_edge = edge;
    return block_sum_1024(local);
  };

  // AdaptivelyVoxelFiltered (voxel_filter.cc:40-77)
  float result_edge = opt.max_length;
  int result_count = run_pass(opt.max_length, true);
  bool done = (float)result_count >= opt.min_num_points;
  if (!done) {
    for (float high_length = opt.max_length; high_length > 1e-2f * opt.max_length; high_length /= 2.f) {
      float low_length = high_length / 2.f;
      result_count = run_pass(low_length, true);
      result_edge = low_length;
      if ((float)result_count >= opt.min_num_points) {
        while ((high_length - low_length) / low_length > 1e-1f) {
          const float mid_length = (low_length + high_length) / 2.f;
          const int candidate = run_pass(mid_length, true);
          if ((float)candidate >= opt.min_num_points) {
            low_length = mid_length;
            result_edge = mid_length;
            result_count = candidate;
          } else {
            high_length = mid_length;
          }
        }
        break;
      }
    }
  }
  // materialise `result`: the table must hold the pass that produced it
  if (last_edge != result_edge) run_pass(result_edge, false);
  const int kept = block_compact_1024(
      m, [&](int j) { return __ldcg(tab + slot[j]) == (uint32_t)j; }, [&](int pos, int j) { out[pos] = (int32_t)rows[j]; });
  if (threadIdx.x == 0) {
    keep_counts[pair] = kept;
    num_passes[pair] = npass;
  }
}

}  // namespace

int launch_voxel_filter(dl_context* ctx, const float* points, int stride, int64_t cap, const int32_t* counts, int batch,
                        float resolution, uint32_t* table, int64_t table_cap, uint32_t* slot, int32_t* keep,
                        int32_t* keep_counts, int32_t* block_counts) {
  if (batch <= 0 || cap <= 0) return DL_OK;
  DL_CUDA(ctx, cudaMemsetAsync(table, 0xFF, (size_t)batch * table_cap * sizeof(uint32_t), ctx->stream));
  const int tiles = (int)((cap + kBlock - 1) / kBlock);
  const dim3 grid(tiles, batch);
  voxel_insert_kernel<<<grid, kBlock, 0, ctx->stream>>>(points, stride, cap, counts, resolution, table, table_cap, slot);
  DL_LAUNCH_CHECK(ctx, "voxel_insert_kernel");
  voxel_count_kernel<<<grid, kBlock, 0, ctx->stream>>>(counts, cap, table, table_cap, slot, block_counts, tiles);
  DL_LAUNCH_CHECK(ctx, "voxel_count_kernel");
  voxel_scatter_kernel<<<grid, kBlock, 0, ctx->stream>>>(counts, cap, table, table_cap, slot, block_counts, tiles, keep,
                                                         keep_counts);
  DL_LAUNCH_CHECK(ctx, "voxel_scatter_kernel");
  return DL_OK;
}

int launch_voxel_indices(dl_context* ctx, const float* points, int stride, int64_t n, float resolution, int32_t* out) {
  if (n <= 0) return DL_OK;
  voxel_indices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(points, stride, n, resolution, out);
  DL_LAUNCH_CHECK(ctx, "voxel_indices_kernel");
  return DL_OK;
}

int launch_adaptive_voxel_filter(dl_context* ctx, const float* points, int stride, int64_t cap, const int32_t* counts,
                                 int batch, const AdaptiveParams* filters_dev, int num_filters, uint32_t* table,
                                 int64_t table_cap, uint32_t* scratch, int32_t* keep, int32_t* keep_counts,
                                 float* passes, int32_t* num_passes, int32_t* cropped_counts) {
  if (batch <= 0 || num_filters <= 0) return DL_OK;
  adaptive_voxel_kernel<<<batch * num_filters, kAdaptiveBlock, 0, ctx->stream>>>(
      points, stride, cap, counts, filters_dev, num_filters, table, table_cap, scratch, keep, keep_counts, passes,
      num_passes, cropped_counts);
  DL_LAUNCH_CHECK(ctx, "adaptive_voxel_kernel");
  return DL_OK;
}

}  // namespace dl
