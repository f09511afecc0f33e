// Exhaustive correlative scan matcher: the correlation-score cube.
//
// Replaces RealTimeCorrelativeScanMatcher3D::Match / ScoreCandidate
// (SM/real_time_correlative_scan_matcher_3d.cc:34-53, :97-113). The reference transforms the whole cloud once
// per candidate and sums nearest-voxel probabilities in float, sequentially in point order. To reproduce every
// score BIT FOR BIT, each candidate here is owned by one thread that walks the cloud in the same order with the
// same float operations (compiled -fmad=false); the cloud is streamed through shared memory in 512-point tiles
// (one coalesced read per block per tile, then warp-broadcast reads), and the candidate-dependent part of the
// pose (rotation per `r`, translation per `l`) is precomputed on the host with the reference's float ops.
// Candidate index = l * R + r = the reference's emplace order (loops z,y,x,rz,ry,rx; :74-92).
//
// Argmax keeps the reference's strict '>' semantics: (score bits << 32 | ~index) is maximised with a 64-bit
// atomicMax, so among equal scores the lowest index wins — scores are positive floats, whose bit patterns order
// like the values.
//
// Algorithmic traffic: per rotation the cloud is read once and every (point, translation) reads one
// 2-byte voxel: R * (12 N + 2 N L) bytes. The voxel reads are L1/L2 hits (a translation window touches <= 8 bricks).
#include "dl_internal.cuh"

namespace dl {
namespace {

constexpr int kBlock = 128;
constexpr int kTile = 512;

__global__ void __launch_bounds__(kBlock) rtcsm_score_kernel(GridView grid, RtcsmLaunch p) {
  __shared__ float tile[kTile * 3];
  const int64_t K = p.R * p.L;
  const int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const bool active = idx < K;
  const int64_t l = active ? idx / p.R : 0;
  const int64_t r = active ? idx - l * p.R : 0;
  const Rigidf cand{p.cand_t[l], p.cand_q[r]};
  const float res = grid.resolution;

  float score = 0.f;
  for (int64_t base = 0; base < p.n; base += kTile) {
    const int count = (int)min((int64_t)kTile, p.n - base);
    __syncthreads();
    for (int j = threadIdx.x; j < count * 3; j += kBlock) tile[j] = p.points[base * 3 + j];
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int j = 0; j < count; ++j) {
        const Vec3f w = apply(cand, Vec3f{tile[3 * j], tile[3 * j + 1], tile[3 * j + 2]});
        const Int3 c = cell_index(w, res);
        score += value_to_probability(grid_value(grid, c.x, c.y, c.z));
      }
    }
  }
  unsigned long long packed = 0ull;
  if (active) {
    score /= (float)p.n;
    // float * double -> double; exp in double; narrowed on assignment (cc:103-110)
    const double a = p.pen_t[l] + p.pen_r[r];
    score = (float)((double)score * exp(-(a * a)));
    if (p.scores) p.scores[idx] = score;
    if (score > 0.f) packed = ((unsigned long long)__float_as_uint(score) << 32) | (0xFFFFFFFFull - (unsigned long long)idx);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, packed, d);
    packed = o > packed ? o : packed;
  }
  if ((threadIdx.x & 31) == 0 && packed) atomicMax(p.best_packed, packed);
}

__global__ void max_range_kernel(const float* __restrict__ points, int64_t n, float* out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    m = fmaxf(m, norm3(Vec3f{points[3 * i], points[3 * i + 1], points[3 * i + 2]}));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  // non-negative floats order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned*)out, __float_as_uint(m));
}

}  // namespace

int launch_rtcsm(dl_context* ctx, const GridView& grid, const RtcsmLaunch& p) {
  const int64_t K = p.R * p.L;
  if (K <= 0 || p.n <= 0) return DL_OK;
  const int64_t blocks = (K + kBlock - 1) / kBlock;
  rtcsm_score_kernel<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(grid, p);
  DL_LAUNCH_CHECK(ctx, "rtcsm_score_kernel");
  return DL_OK;
}

// out must be pre-set to the float `init` (3 * resolution in the reference, cc:63-66).
int launch_max_range(dl_context* ctx, const float* points, int64_t n, float init, float* out) {
  DL_CUDA(ctx, cudaMemcpyAsync(out, &init, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  if (n <= 0) return DL_OK;
  const int blocks = (int)min((int64_t)kNumSMs * 4, (n + 255) / 256);
  max_range_kernel<<<blocks, 256, 0, ctx->stream>>>(points, n, out);
  DL_LAUNCH_CHECK(ctx, "max_range_kernel");
  return DL_OK;
}

}  // namespace dl
