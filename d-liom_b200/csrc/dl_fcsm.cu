// Loop-closure coarse matcher: FastCorrelativeScanMatcher3D::MatchWith3DofInitial
// (SM/fast_correlative_scan_matcher_3d.cc:165-196) without the branch and bound.
//
// The reference prunes the (x, y, z) translation window with a stack of max-pooled 8-bit "precomputation" grids
// (PrecomputationGridStack3D, :57-77; PrecomputeGrid, precomputation_grid_3d.cc:62-81) because a CPU cannot afford the
// ~10^5 leaves x 10^3 points of the full window. The bounds are exact, so branch and bound returns the best-scoring
// leaf that passes the low-resolution gate — which is also what scoring EVERY leaf returns. On a B200 the brute-force
// cube is a few hundred microseconds of L1/L2-resident integer gathers, needs no precomputation stack at all (the
// 8-bit value is derived from the uint16 cell on the fly with the reference's float expression,
// precomputation_grid_3d.cc:50-53) and has no data-dependent control flow:
//   cells      c_i = GetCellIndex(pose * p_i)                                 (float, exact)
//   leaf score s(o) = ToProbability( sum_i V8[c_i + o] / float(N) )           (integer sum: order-independent, exact)
//   result     argmax_o s(o) subject to s(o) > min_score and the low-resolution gate (low_resolution_matcher.cc:24-36),
//              lowest linear index (z, y, x order) among equal scores; the reference's own tie order is that of an
//              unstable std::sort.
#include "dl_internal.cuh"

namespace dl {
namespace {

constexpr int kBlock = 128;
constexpr int kTile = 1024;

__global__ void fcsm_cells_kernel(const float* __restrict__ points, int n, Rigidf pose, float resolution, int* __restrict__ cells) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Int3 c = cell_index(apply(pose, Vec3f{points[3 * i], points[3 * i + 1], points[3 * i + 2]}), resolution);
  cells[3 * i] = c.x; cells[3 * i + 1] = c.y; cells[3 * i + 2] = c.z;
}

// ConvertToPrecomputationGrid's per-cell expression
__device__ __forceinline__ int precomputation_value(uint16_t v) {
  const float kMin = 0.1f, kMax = 1.f - 0.1f;
  return round_to_int((value_to_probability(v) - kMin) * (255.f / (kMax - kMin)));
}

__global__ void __launch_bounds__(kBlock) fcsm_score_kernel(GridView grid, const int* __restrict__ cells, int n, int wxy,
                                                            int wz, float* __restrict__ scores) {
  __shared__ int tile[kTile * 3];
  const int side = 2 * wxy + 1;
  const long long K = (long long)side * side * (2 * wz + 1);
  const long long idx = (long long)blockIdx.x * kBlock + threadIdx.x;
  const bool active = idx < K;
  const int ox = active ? (int)(idx % side) - wxy : 0;
  const int oy = active ? (int)((idx / side) % side) - wxy : 0;
  const int oz = active ? (int)(idx / ((long long)side * side)) - wz : 0;
  int sum = 0;
  for (int base = 0; base < n; base += kTile) {
    const int count = min(kTile, n - base);
    __syncthreads();
    for (int j = threadIdx.x; j < count * 3; j += kBlock) tile[j] = cells[base * 3 + j];
    __syncthreads();
    if (active)
      for (int j = 0; j < count; ++j)
        sum += precomputation_value(grid_value(grid, tile[3 * j] + ox, tile[3 * j + 1] + oy, tile[3 * j + 2] + oz));
  }
  if (active) {
    const float kMin = 0.1f, kMax = 1.f - 0.1f;
    scores[idx] = kMin + ((float)sum / (float)n) * ((kMax - kMin) / 255.f);  // PrecomputationGrid3D::ToProbability(sum / float(n))
  }
}

__global__ void fcsm_argmax_kernel(const float* __restrict__ scores, long long K, float min_score, unsigned long long* best) {
  unsigned long long packed = 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < K; i += (long long)gridDim.x * blockDim.x) {
    const float s = scores[i];
    if (s > min_score && s > 0.f) {
      const unsigned long long p = ((unsigned long long)__float_as_uint(s) << 32) | (0xFFFFFFFFull - (unsigned long long)i);
      packed = p > packed ? p : packed;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, packed, d);
    packed = o > packed ? o : packed;
  }
  if ((threadIdx.x & 31) == 0 && packed) atomicMax(best, packed);
}

// low_resolution_matcher: mean nearest-voxel probability, float sum in point order (one thread: the order matters)
__global__ void fcsm_gate_kernel(GridView lo, const float* __restrict__ points, int n, Rigidf pose, float* out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  float score = 0.f;
  for (int i = 0; i < n; ++i) {
    const Int3 c = cell_index(apply(pose, Vec3f{points[3 * i], points[3 * i + 1], points[3 * i + 2]}), lo.resolution);
    score += value_to_probability(grid_value(lo, c.x, c.y, c.z));
  }
  *out = score / (float)n;
}

__global__ void fcsm_reject_kernel(float* scores, long long idx) { scores[idx] = -1.f; }

}  // namespace

int launch_fcsm_cells(dl_context* ctx, const float* points, int n, const Rigidf& pose, float resolution, int* cells) {
  fcsm_cells_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(points, n, pose, resolution, cells);
  DL_LAUNCH_CHECK(ctx, "fcsm_cells_kernel");
  return DL_OK;
}
int launch_fcsm_scores(dl_context* ctx, const GridView& grid, const int* cells, int n, int wxy, int wz, float* scores) {
  const long long side = 2 * wxy + 1, K = side * side * (2 * wz + 1);
  fcsm_score_kernel<<<(unsigned)((K + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(grid, cells, n, wxy, wz, scores);
  DL_LAUNCH_CHECK(ctx, "fcsm_score_kernel");
  return DL_OK;
}
int launch_fcsm_argmax(dl_context* ctx, const float* scores, long long K, float min_score, unsigned long long* best) {
  DL_CUDA(ctx, cudaMemsetAsync(best, 0, sizeof(unsigned long long), ctx->stream));
  fcsm_argmax_kernel<<<(unsigned)std::min<long long>(kNumSMs * 4, (K + 255) / 256), 256, 0, ctx->stream>>>(scores, K, min_score, best);
  DL_LAUNCH_CHECK(ctx, "fcsm_argmax_kernel");
  return DL_OK;
}
int launch_fcsm_gate(dl_context* ctx, const GridView& lo, const float* points, int n, const Rigidf& pose, float* out) {
  fcsm_gate_kernel<<<1, 32, 0, ctx->stream>>>(lo, points, n, pose, out);
  DL_LAUNCH_CHECK(ctx, "fcsm_gate_kernel");
  return DL_OK;
}
int launch_fcsm_reject(dl_context* ctx, float* scores, long long idx) {
  fcsm_reject_kernel<<<1, 1, 0, ctx->stream>>>(scores, idx);
  DL_LAUNCH_CHECK(ctx, "fcsm_reject_kernel");
  return DL_OK;
}

}  // namespace dl
