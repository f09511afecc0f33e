// Fused scan front half for the batched front end: 7 launches per sub-batch instead of the 11 of the stage-wise path
// (dl_voxel.cu + dl_ingest.cu, which stay as the standalone filter API and as a cross-check in the tests).
//
//   A  fe_first_filter_insert   first voxel filter (LTB:393-395): every point proposes its index for its voxel
//                               (atomicCAS claim + atomicMin), one 128-bit load per point.
//   B  fe_ingest_second_insert  for each first-filter survivor: deskew + transform + range gate (LTB:426-472) and
//                               immediately the SECOND voxel filter's insert (LTB:479-484) keyed on the local-frame
//                               voxel — no compaction in between: ids stay the original input indices, which
//                               preserves "first point in input order wins" for free.
//   C1 fe_count_tiles / C2 fe_scatter_tracking   ordered compaction of the second filter's survivors, fused with
//                               the frame change back to tracking (TransformRangeData with current_pose^-1, LTB:485-487).
//
// The second filter's table stores packed 63-bit voxel keys (3 x 21 bits) + a min-index array, so a collision
// never has to read another thread's freshly written point (no fence, no race). Keys outside +-2^20 voxels (> 75 km
// at 0.075 m) set an error flag and the call fails with DL_ERR_ARG; the generic dl_voxel_filter has no such limit.
// Returns and misses share the table (bit 63 distinguishes them).
//
// Algorithmic traffic per raw point: A reads 16 B; B reads 4 B slot + 4 B owner (+16 B row, writes 13 B for
// survivors); C reads 1 B class (+ survivors' 12 B, writes 12 B per output point).
#include <cstdlib>

#include "dl_internal.cuh"
#include "dl_pipeline.cuh"

namespace dl {
namespace {

constexpr uint32_t kEmpty32 = 0xFFFFFFFFu;
constexpr unsigned long long kEmpty64 = 0xFFFFFFFFFFFFFFFFull;
constexpr int kBlock = 256;

__device__ __forceinline__ uint32_t hash_cell(const Int3& c) {
  uint32_t h = (uint32_t)c.x * 73856093u ^ (uint32_t)c.y * 19349663u ^ (uint32_t)c.z * 83492791u;
  h ^= h >> 15;
  h *= 0x2c1b3c6du;
  h ^= h >> 12;
  return h;
}

__device__ __forceinline__ Vec3f load_xyz(const float* __restrict__ rows, int row_floats, uint32_t i) {
  if (row_floats == 3) {
    const float* p = rows + (size_t)i * 3;
    return {p[0], p[1], p[2]};
  }
  const float4 v = __ldg((const float4*)(rows + (size_t)i * row_floats));  // rows are 16- or 32-byte records
  return {v.x, v.y, v.z};
}

// Time of row i of scan b: the row's fourth float, or (12-byte rows) the value of the run that contains the row.
__device__ __forceinline__ float point_time(const FrontendArgs& a, int b, const float* __restrict__ rows, int rf, int i) {
  if (rf != 3) return rows[(size_t)i * rf + 3];
  if (a.times) return a.times[(size_t)b * a.in_cap + i];
  int lo = a.run_offsets[b], hi = a.run_offsets[b + 1] - 1;  // last run whose first row is <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (a.run_first_row[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return a.run_value[lo];
}

// 12-byte rows: one float per row from the runs, written once per batch (one warp per run, contiguous stores), so that the
// latency-bound ingest kernel pays one 4-byte load per survivor instead of an 11-step search of the run table.
__global__ void __launch_bounds__(kBlock) fe_expand_times(FrontendArgs a, float* __restrict__ times) {
  const int b = blockIdx.y;
  const int r0 = a.run_offsets[b], r1 = a.run_offsets[b + 1];
  const int r = r0 + blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  if (r >= r1) return;
  const int n = a.counts[b];
  const int begin = max(0, a.run_first_row[r]), end = min(n, r + 1 < r1 ? a.run_first_row[r + 1] : n);  // clamped to the scan
  const float t = a.run_value[r];
  float* out = times + (size_t)b * a.in_cap;
  for (int i = begin + (threadIdx.x & 31); i < end; i += 32) out[i] = t;
}

// ---------------------------------------------------------------------------------------------------- A
__global__ void __launch_bounds__(kBlock) fe_first_filter_insert(FrontendArgs a) {
  const int b = a.first_scan + blockIdx.y;
  const int n = a.counts[b];
  const float* rows = a.ranges + (size_t)b * a.in_cap * a.row_floats;
  uint32_t* tab = a.table1 + (size_t)b * a.tcap1;
  const uint32_t mask = (uint32_t)a.tcap1 - 1;
  const CellDivider res = make_divider(a.first_resolution);
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const Int3 c = cell_index(load_xyz(rows, a.row_floats, i), res);
    uint32_t h = hash_cell(c) & mask;
    for (;;) {
      const uint32_t prev = atomicCAS(tab + h, kEmpty32, (uint32_t)i);
      if (prev == kEmpty32) break;
      const Int3 o = cell_index(load_xyz(rows, a.row_floats, prev), res);
      if (o.x == c.x && o.y == c.y && o.z == c.z) {
        if ((uint32_t)i < prev) atomicMin(tab + h, (uint32_t)i);  // the owner only ever decreases
        break;
      }
      h = (h + 1) & mask;
    }
  }
}

// ---------------------------------------------------------------------------------------------------- B
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

// pose applied to the point with per-point time t (LTB:430-445)
__device__ __forceinline__ Rigidf point_pose(const ScanConstants& sc, bool no_deskew, double scan_period, float t) {
  if (no_deskew) return to_float(sc.cur);
  const double s = (scan_period + (double)t) / scan_period;
  return to_float(compose(sc.prev, interpolate_pose(s, sc)));
}

__device__ __forceinline__ bool pack_key(const Int3& c, bool miss, unsigned long long* key) {
  const int lim = 1 << 20;
  if (c.x < -lim || c.x >= lim - 1 || c.y < -lim || c.y >= lim - 1 || c.z < -lim || c.z >= lim - 1) return false;
  *key = ((unsigned long long)(c.x + lim) << 42) | ((unsigned long long)(c.y + lim) << 21) | (unsigned long long)(c.z + lim) |
         (miss ? (1ull << 63) : 0ull);
  return true;
}

// Heavy per-survivor work of kernel B (double slerp, pose composition, transform, gate, second-filter insert).
__device__ __forceinline__ int ingest_survivor(const FrontendArgs& a, int b, const float* rows, int rf, const ScanConstants& sc,
                                               bool no_deskew, unsigned long long* keys, uint32_t* mins, uint32_t mask2, int i) {
  float4 h;
  if (rf == 3) {
    const float* p = rows + (size_t)i * 3;
    h = make_float4(p[0], p[1], p[2], point_time(a, b, rows, rf, i));
  } else {
    h = __ldg((const float4*)(rows + (size_t)i * rf));
  }
  const unsigned long long origin_index = rf >= 8 ? *(const unsigned long long*)(rows + (size_t)i * rf + 4) : 0ull;
  const float* o = a.origins + 3 * origin_index;
  const Rigidf pose = point_pose(sc, no_deskew, a.scan_period, h.w);
  const Vec3f hit = apply(pose, Vec3f{h.x, h.y, h.z});
  const Vec3f org = apply(pose, Vec3f{o[0], o[1], o[2]});
  const Vec3f delta = sub(hit, org);
  const float range = norm3(delta);
  Vec3f outp = hit;
  int cls = 0;
  if (range >= a.min_range) {
    if (range <= a.max_range) {
      cls = 1;
    } else {
      cls = 2;
      outp = add(org, mul(a.max_range / range, delta));
    }
  }
  if (cls) {
    // one aligned 16-byte record per survivor: local-frame point + class (1 return, 2 miss) in .w
    ((float4*)a.local)[(size_t)b * a.cap + i] = make_float4(outp.x, outp.y, outp.z, __int_as_float(cls));
    const Int3 c = cell_index(outp, make_divider(a.second_resolution));
    unsigned long long key;
    if (!pack_key(c, cls == 2, &key)) {
      a.error_flag[b] = 1;  // per scan: only this scan's result is invalidated
      cls = 0;
    } else {
      uint32_t hh = (hash_cell(c) ^ (cls == 2 ? 0x9e3779b9u : 0u)) & mask2;
      for (;;) {
        const unsigned long long prev = atomicCAS(keys + hh, kEmpty64, key);
        if (prev == kEmpty64 || prev == key) {
          // (index << 1) | miss: ordered by index (a voxel key holds one class only), and the winner's class reaches
          // the compaction kernels through the byte map without a second look at the survivor record
          atomicMin(mins + hh, ((uint32_t)i << 1) | (cls == 2 ? 1u : 0u));
          break;
        }
        hh = (hh + 1) & mask2;
      }
    }
  }
  return cls;
}

// Only ~30 % of the raw points survive the first filter, so running the heavy path under the survivor predicate
// would leave most lanes idle. Each warp instead appends its survivors to a small shared-memory queue and drains
// it 32 at a time with all lanes busy.
__global__ void __launch_bounds__(kBlock) fe_ingest_second_insert(FrontendArgs a) {
  __shared__ int queue[kBlock / 32][64];
  const int b = a.first_scan + blockIdx.y;
  const int n = a.counts[b];
  const int rf = a.row_floats;
  const float* rows = a.ranges + (size_t)b * a.in_cap * rf;
  const uint32_t* tab = a.table1 + (size_t)b * a.tcap1;
  unsigned long long* keys = a.keys2 + (size_t)b * a.tcap2;
  uint32_t* mins = a.min2 + (size_t)b * a.tcap2;
  const uint32_t mask2 = (uint32_t)a.tcap2 - 1;
  const ScanConstants& sc = a.scans[b];
  const bool no_deskew = n > 0 && (double)fabsf(point_time(a, b, rows, rf, 0)) < 1e-3;  // first survivor is row 0 (LTB:430-433)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* q = queue[warp];
  int queued = 0, survivors = 0, returns = 0, last = -1;
  // The first filter's survivors are exactly the non-empty table slots: stream the table (coalesced) instead of
  // probing it once per raw point. Order does not matter here: the second filter keys on the original index.
  if (n == 0) return;
  for (int h = blockIdx.x * kBlock + threadIdx.x; h < (int)a.tcap1; h += gridDim.x * kBlock) {
    const uint32_t owner = __ldcg(tab + h);
    const bool surv = owner != kEmpty32;
    const int i = (int)owner;
    const unsigned ballot = __ballot_sync(0xffffffffu, surv);
    if (surv) {
      q[queued + __popc(ballot & ((1u << lane) - 1))] = i;
      last = max(last, i);
    }
    queued += __popc(ballot);
    survivors += surv;
    __syncwarp();
    if (queued >= 32) {
      returns += ingest_survivor(a, b, rows, rf, sc, no_deskew, keys, mins, mask2, q[lane]) == 1;
      __syncwarp();
      const int rest = queued - 32;
      const int moved = lane < rest ? q[32 + lane] : 0;
      __syncwarp();
      if (lane < rest) q[lane] = moved;
      queued = rest;
      __syncwarp();
    }
  }
  if (lane < queued) returns += ingest_survivor(a, b, rows, rf, sc, no_deskew, keys, mins, mask2, q[lane]) == 1;
  // bookkeeping: survivor counts and the LAST first-filter survivor (hits_poses.back(), LTB:476)
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    survivors += __shfl_xor_sync(0xffffffffu, survivors, d);
    returns += __shfl_xor_sync(0xffffffffu, returns, d);
    last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
  }
  if (lane == 0) {
    if (survivors) atomicAdd(a.n_first + b, survivors);
    if (returns) atomicAdd(a.n_returns_local + b, returns);
    if (last >= 0) atomicMax(a.last_index + b, last);
  }
}

// ---------------------------------------------------------------------------------------------------- C
__device__ __forceinline__ int block_exclusive_scan(int value, int* total) {
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

// win[i] is set by fe_mark_winners for the points that own their second-filter voxel: 1 return, 2 miss, 0 neither.
__device__ __forceinline__ int second_filter_class(const FrontendArgs& a, int b, int i, int n) {
  if (i >= n) return 0;
  return a.win[(size_t)b * a.cap + i];
}

// The second filter's survivors are the min-index entries of its non-empty slots: stream the table once.
__global__ void __launch_bounds__(kBlock) fe_mark_winners(FrontendArgs a) {
  const int b = a.first_scan + blockIdx.y;
  if (a.counts[b] == 0) return;
  const uint32_t* mins = a.min2 + (size_t)b * a.tcap2;
  uint8_t* win = a.win + (size_t)b * a.cap;
  for (int h = blockIdx.x * kBlock + threadIdx.x; h < (int)a.tcap2; h += gridDim.x * kBlock) {
    const uint32_t m = __ldcg(mins + h);
    if (m != kEmpty32) win[m >> 1] = (uint8_t)(1 + (m & 1u));  // store only: no read-modify-write latency
  }
}

__global__ void __launch_bounds__(kBlock) fe_count_tiles(FrontendArgs a) {
  const int b = a.first_scan + blockIdx.y;
  const int n = a.counts[b];
  if ((int)blockIdx.x * kBlock >= n) {
    if (threadIdx.x == 0) {
      a.tile_counts[((size_t)b * a.tiles + blockIdx.x) * 2] = 0;
      a.tile_counts[((size_t)b * a.tiles + blockIdx.x) * 2 + 1] = 0;
    }
    return;
  }
  const int cls = second_filter_class(a, b, blockIdx.x * kBlock + threadIdx.x, n);
  const int r = __syncthreads_count(cls == 1);
  const int m = __syncthreads_count(cls == 2);
  if (threadIdx.x == 0) {
    a.tile_counts[((size_t)b * a.tiles + blockIdx.x) * 2] = r;
    a.tile_counts[((size_t)b * a.tiles + blockIdx.x) * 2 + 1] = m;
  }
}

// Exclusive prefix of the tile counts of one scan (one CTA per scan), in place; totals go to n_returns / n_misses.
__global__ void __launch_bounds__(kBlock) fe_tile_prefix(FrontendArgs a) {
  const int b = a.first_scan + blockIdx.x;
  const int n = a.counts[b];
  const int my_tiles = (n + kBlock - 1) / kBlock;
  int32_t* tc = a.tile_counts + (size_t)b * a.tiles * 2;
  __shared__ int carry_r, carry_m;
  if (threadIdx.x == 0) carry_r = carry_m = 0;
  __syncthreads();
  for (int base = 0; base < my_tiles; base += kBlock) {
    const int t = base + threadIdx.x;
    const int r = t < my_tiles ? tc[2 * t] : 0, m = t < my_tiles ? tc[2 * t + 1] : 0;
    int tr, tm;
    const int er = block_exclusive_scan(r, &tr);
    const int em = block_exclusive_scan(m, &tm);
    if (t < my_tiles) {
      tc[2 * t] = carry_r + er;
      tc[2 * t + 1] = carry_m + em;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      carry_r += tr;
      carry_m += tm;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    a.n_returns[b] = carry_r;
    a.n_misses[b] = carry_m;
  }
}

__global__ void __launch_bounds__(kBlock) fe_scatter_tracking(FrontendArgs a) {
  const int b = a.first_scan + blockIdx.y;
  const int n = a.counts[b];
  if ((int)blockIdx.x * kBlock >= n) return;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  const int cls = second_filter_class(a, b, i, n);
  const unsigned lane_mask = (1u << (threadIdx.x & 31)) - 1;
  const unsigned br = __ballot_sync(0xffffffffu, cls == 1), bm = __ballot_sync(0xffffffffu, cls == 2);
  __shared__ int wr[kBlock / 32], wm[kBlock / 32];
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    wr[warp] = __popc(br);
    wm[warp] = __popc(bm);
  }
  __syncthreads();
  if (!cls) return;
  int off = a.tile_counts[((size_t)b * a.tiles + blockIdx.x) * 2 + (cls == 2)];
  for (int w = 0; w < warp; ++w) off += cls == 1 ? wr[w] : wm[w];
  off += __popc((cls == 1 ? br : bm) & lane_mask);
  const float* bp = a.back_pose + 7 * b;
  const Rigidf back{{bp[0], bp[1], bp[2]}, {bp[3], bp[4], bp[5], bp[6]}};
  const float4 l = ((const float4*)a.local)[(size_t)b * a.cap + i];
  const Vec3f q = apply(back, Vec3f{l.x, l.y, l.z});
  float* dst = (cls == 1 ? a.returns_tracking : a.misses_tracking) + ((size_t)b * a.cap + off) * 3;
  dst[0] = q.x; dst[1] = q.y; dst[2] = q.z;
}

// current_pose = pose of the LAST first-filter survivor (hits_poses.back(), LTB:476) and its inverse, once per scan.
__global__ void fe_current_pose(FrontendArgs a, int batch) {
  const int b = a.first_scan + blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.first_scan + batch) return;
  const int rf = a.row_floats;
  const float* rows = a.ranges + (size_t)b * a.in_cap * rf;
  const int last = a.last_index[b];
  Rigidf cur = to_float(a.scans[b].cur);
  if (last >= 0) {
    const bool no_deskew = (double)fabsf(point_time(a, b, rows, rf, 0)) < 1e-3;
    cur = point_pose(a.scans[b], no_deskew, a.scan_period, point_time(a, b, rows, rf, last));
  }
  const Rigidf back = inverse(cur);
  float* cp = a.current_pose + 7 * b;
  cp[0] = cur.t.x; cp[1] = cur.t.y; cp[2] = cur.t.z; cp[3] = cur.q.w; cp[4] = cur.q.x; cp[5] = cur.q.y; cp[6] = cur.q.z;
  float* bp = a.back_pose + 7 * b;
  bp[0] = back.t.x; bp[1] = back.t.y; bp[2] = back.t.z; bp[3] = back.q.w; bp[4] = back.q.x; bp[5] = back.q.y; bp[6] = back.q.z;
}

__global__ void fe_reset_counters(FrontendArgs a, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  a.n_first[b] = 0;
  a.n_returns_local[b] = 0;
  a.last_index[b] = -1;
  if (a.counts[b] == 0) {  // nothing will write these for an empty scan
    a.n_returns[b] = 0;
    a.n_misses[b] = 0;
    const Rigidf cur = to_float(a.scans[b].cur);
    float* cp = a.current_pose + 7 * b;
    cp[0] = cur.t.x; cp[1] = cur.t.y; cp[2] = cur.t.z; cp[3] = cur.q.w; cp[4] = cur.q.x; cp[5] = cur.q.y; cp[6] = cur.q.z;
  }
  a.error_flag[b] = 0;
}

}  // namespace

int launch_fe_prepare(dl_context* ctx, const FrontendArgs& a, int batch) {
  DL_CUDA(ctx, cudaMemsetAsync(a.table1, 0xFF, (size_t)batch * a.tcap1 * sizeof(uint32_t), ctx->stream));
  DL_CUDA(ctx, cudaMemsetAsync(a.keys2, 0xFF, (size_t)batch * a.tcap2 * sizeof(unsigned long long), ctx->stream));
  DL_CUDA(ctx, cudaMemsetAsync(a.min2, 0xFF, (size_t)batch * a.tcap2 * sizeof(uint32_t), ctx->stream));
  DL_CUDA(ctx, cudaMemsetAsync(a.win, 0, (size_t)batch * a.cap, ctx->stream));
  fe_reset_counters<<<(batch + 127) / 128, 128, 0, ctx->stream>>>(a, batch);
  DL_LAUNCH_CHECK(ctx, "fe_reset_counters");
  return DL_OK;
}

int launch_fe_expand_times(dl_context* ctx, const FrontendArgs& a, int batch, int max_runs_per_scan, float* times_out) {
  if (batch <= 0 || max_runs_per_scan <= 0) return DL_OK;
  fe_expand_times<<<dim3((max_runs_per_scan + kBlock / 32 - 1) / (kBlock / 32), batch), kBlock, 0, ctx->stream>>>(a, times_out);
  DL_LAUNCH_CHECK(ctx, "fe_expand_times");
  return DL_OK;
}

// Kernel A for scans [first_scan, first_scan + num_scans): lets the host overlap the upload of later scans.
int launch_fe_first_filter(dl_context* ctx, FrontendArgs a, int first_scan, int num_scans) {
  if (num_scans <= 0) return DL_OK;
  a.first_scan = first_scan;
  const int tiles = (int)std::min<int64_t>((a.cap + kBlock - 1) / kBlock, 128);
  fe_first_filter_insert<<<dim3(tiles, num_scans), kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "fe_first_filter_insert");
  return DL_OK;
}

int launch_fe_rest(dl_context* ctx, FrontendArgs a, int first_scan, int batch) {
  if (batch <= 0) return DL_OK;
  a.first_scan = first_scan;
  const int tiles = (int)std::min<int64_t>((a.cap + kBlock - 1) / kBlock, 128);
  int per_scan = 48;  // measured best of {8, 20, 32, 48, 64}: enough CTAs in flight to hide the random-access latency
  if (const char* env = std::getenv("DLIOM_INGEST_GRID")) per_scan = std::max(1, std::atoi(env));
  fe_ingest_second_insert<<<dim3(per_scan, batch), kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "fe_ingest_second_insert");
  fe_mark_winners<<<dim3(tiles, batch), kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "fe_mark_winners");
  fe_current_pose<<<(batch + 127) / 128, 128, 0, ctx->stream>>>(a, batch);
  DL_LAUNCH_CHECK(ctx, "fe_current_pose");
  fe_count_tiles<<<dim3(a.tiles, batch), kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "fe_count_tiles");
  fe_tile_prefix<<<batch, kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "fe_tile_prefix");
  fe_scatter_tracking<<<dim3(a.tiles, batch), kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "fe_scatter_tracking");
  return DL_OK;
}

}  // namespace dl
