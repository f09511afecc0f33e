// Range-data insertion into the device-resident probability grid (the step right after the hot path).
//
// Replaces RangeDataInserter3D::Insert + InsertMissesIntoGrid (C/mapping/3d/range_data_inserter_3d.cc:27-51, :76-92),
// HybridGrid::ApplyLookupTable / FinishUpdate (C/mapping/3d/hybrid_grid.h:494-520) and the growth path of
// DynamicGrid / NestedGrid::mutable_value (hybrid_grid.h:285-300, :165-176, :389-407) so that the device grid is the
// PRIMARY copy: no per-scan host insert + dirty-brick upload.
//
// Semantics preserved exactly: within one Insert every cell is updated at most once (the update marker, bit 15);
// all hits are applied before any miss, so a hit wins over a miss in the same cell; all hits use one table and all
// misses another, so the order inside each group is irrelevant and the two groups parallelise freely:
//     cell' = hit_table[cell]  if the cell is a hit cell, else miss_table[cell] if it is a miss cell, else cell.
// Concurrent updates of one uint16 cell go through a 32-bit compare-and-swap on the containing word.
// Structure growth is lock-free and spin-free: (1) claim missing top entries, (2) claim missing node entries — one
// kernel each, so nobody ever waits for an allocation made in the same kernel — then (3) hits, (4) misses,
// (5) FinishUpdate over the list of touched cells.
#include "dl_internal.cuh"
#include "dl_pipeline.cuh"

namespace dl {
namespace {

constexpr int kBlock = 256;

struct InsertArgs {
  const float* returns;  // n x 3, already in the grid's frame
  int n;
  Vec3f origin;
  float resolution;
  int bits;
  int num_free;
  int32_t* top;
  int32_t* nodes;
  uint16_t* bricks;
  int32_t* counters;     // [0] nodes in use, [1] bricks in use, [2] update-list length, [3] nodes / [4] bricks this Insert adds
  int32_t* bbox;         // min xyz, max xyz
  uint32_t* update_list;
  const uint16_t* hit_table;
  const uint16_t* miss_table;
};

__device__ __forceinline__ Int3 hit_cell(const InsertArgs& a, int i) {
  return cell_index(Vec3f{a.returns[3 * i], a.returns[3 * i + 1], a.returns[3 * i + 2]}, a.resolution);
}

// Visits the hit cell (phase 0) or the miss cells (phase 1) of ray i, exactly as range_data_inserter_3d.cc:33-50.
template <typename F>
__device__ __forceinline__ void for_each_cell(const InsertArgs& a, int i, int phase, F f) {
  const Int3 h = hit_cell(a, i);
  if (phase == 0) {
    f(h);
    return;
  }
  const Int3 o = cell_index(a.origin, a.resolution);
  const Int3 d{h.x - o.x, h.y - o.y, h.z - o.z};
  const int num_samples = max(abs(d.x), max(abs(d.y), abs(d.z)));
  for (int position = max(0, num_samples - a.num_free); position < num_samples; ++position)
    f(Int3{o.x + d.x * position / num_samples, o.y + d.y * position / num_samples, o.z + d.z * position / num_samples});
}

__global__ void ins_bbox_kernel(InsertArgs a) {
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
  auto add = [&](const Int3& c) {
    lo[0] = min(lo[0], c.x); lo[1] = min(lo[1], c.y); lo[2] = min(lo[2], c.z);
    hi[0] = max(hi[0], c.x); hi[1] = max(hi[1], c.y); hi[2] = max(hi[2], c.z);
  };
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < a.n; i += gridDim.x * kBlock) add(hit_cell(a, i));
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.num_free > 0 && a.n > 0) add(cell_index(a.origin, a.resolution));
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
      hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
    }
  }
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 3; ++k) {
      atomicMin(a.bbox + k, lo[k]);
      atomicMax(a.bbox + 3 + k, hi[k]);
    }
}

// Grow(): every axis doubles, the old content moves to the centre (hybrid_grid.h:389-407).
__global__ void grid_grow_kernel(const int32_t* old_top, int old_bits, int32_t* new_top) {
  const int n = 1 << old_bits;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * n * n) return;
  const int x = e & (n - 1), y = (e >> old_bits) & (n - 1), z = e >> (2 * old_bits);
  const int o = 1 << (old_bits - 1), nb = old_bits + 1;
  new_top[((((z + o) << nb) + (y + o)) << nb) + (x + o)] = old_top[e];
}

__device__ __forceinline__ void shifted(const InsertArgs& a, const Int3& c, unsigned* sx, unsigned* sy, unsigned* sz) {
  const int half = (64 << a.bits) >> 1;
  *sx = (unsigned)(c.x + half); *sy = (unsigned)(c.y + half); *sz = (unsigned)(c.z + half);
}

// level 0: make sure the top entry has a node; level 1: make sure the node entry has a brick. Two passes per level so that the
// pools are reserved for EXACTLY the entries this Insert creates (round 1 reserved the worst case, one node and one brick per
// touched cell: ~400 MB per grid for a 30k-point scan with two free-space voxels, a few hundred bricks actually used):
//   ASSIGN = 0  mark every missing entry this Insert touches (-1 -> -2) and count them in counters[3 + LEVEL];
//   ASSIGN = 1  after the host has grown the pool by that count: give every marked entry its slot (-2 -> index).
template <int LEVEL, int ASSIGN>
__global__ void __launch_bounds__(kBlock) ins_claim_kernel(InsertArgs a, int phase) {
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < a.n; i += gridDim.x * kBlock) {
    for_each_cell(a, i, phase, [&](const Int3& c) {
      unsigned sx, sy, sz;
      shifted(a, c, &sx, &sy, &sz);
      int32_t* entry = a.top + ((((sz >> 6) << a.bits) + (sy >> 6)) << a.bits) + (sx >> 6);
      if (LEVEL == 1) {
        const int node = *(volatile int32_t*)entry;  // assigned by the previous level
        entry = a.nodes + (size_t)node * 512 + ((((sz >> 3) & 7) << 6) | (((sy >> 3) & 7) << 3) | ((sx >> 3) & 7));
      }
      if (ASSIGN == 0) {
        if (*(volatile int32_t*)entry == -1 && atomicCAS(entry, -1, -2) == -1) atomicAdd(a.counters + 3 + LEVEL, 1);
      } else {
        if (*(volatile int32_t*)entry == -2 && atomicCAS(entry, -2, -3) == -2) {
          const int idx = atomicAdd(a.counters + LEVEL, 1);  // pool slots are pre-initialised (-1 nodes / 0 bricks)
          *(volatile int32_t*)entry = idx;
        }
      }
    });
  }
}

__global__ void __launch_bounds__(kBlock) ins_apply_kernel(InsertArgs a, int phase) {
  const uint16_t* table = phase == 0 ? a.hit_table : a.miss_table;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < a.n; i += gridDim.x * kBlock) {
    for_each_cell(a, i, phase, [&](const Int3& c) {
      unsigned sx, sy, sz;
      shifted(a, c, &sx, &sy, &sz);
      const int node = a.top[((((sz >> 6) << a.bits) + (sy >> 6)) << a.bits) + (sx >> 6)];
      const int brick = a.nodes[(size_t)node * 512 + ((((sz >> 3) & 7) << 6) | (((sy >> 3) & 7) << 3) | ((sx >> 3) & 7))];
      const uint32_t cell = (uint32_t)brick * 512u + (((sz & 7) << 6) | ((sy & 7) << 3) | (sx & 7));
      unsigned* word = (unsigned*)a.bricks + (cell >> 1);
      const int shift = (cell & 1) * 16;
      unsigned old = *(volatile unsigned*)word;
      for (;;) {
        const uint16_t v = (uint16_t)(old >> shift);
        if (v >= 32768) return;  // already updated in this Insert (ApplyLookupTable returns false)
        const unsigned desired = (old & ~(0xFFFFu << shift)) | ((unsigned)table[v] << shift);
        const unsigned seen = atomicCAS(word, old, desired);
        if (seen == old) {
          a.update_list[atomicAdd(a.counters + 2, 1)] = cell;
          return;
        }
        old = seen;
      }
    });
  }
}

// FinishUpdate: remove the update marker from every cell touched by this Insert.
__global__ void ins_finish_kernel(InsertArgs a) {
  const int count = a.counters[2];
  for (int k = blockIdx.x * kBlock + threadIdx.x; k < count; k += gridDim.x * kBlock) {
    const uint32_t cell = a.update_list[k];
    atomicSub((unsigned*)a.bricks + (cell >> 1), 32768u << ((cell & 1) * 16));  // the two halves of a word are different cells
  }
}

__global__ void transform_filter_kernel(const float* __restrict__ in, int n, Rigidf to_submap, Vec3f origin_submap,
                                        float max_range, float* __restrict__ all, float* __restrict__ near,
                                        int32_t* near_count, int32_t* tile_counts, int pass) {
  // pass 0: transform + per-tile count of in-range points; pass 1: ordered scatter of the in-range ones
  __shared__ int warp_sums[kBlock / 32];
  const int i = blockIdx.x * kBlock + threadIdx.x;
  Vec3f p{0, 0, 0};
  int flag = 0;
  if (i < n) {
    p = apply(to_submap, Vec3f{in[3 * i], in[3 * i + 1], in[3 * i + 2]});
    flag = norm3(sub(p, origin_submap)) <= max_range;   // FilterRangeDataByMaxRange, submap_3d.cc:42-51
    if (pass == 0) { all[3 * i] = p.x; all[3 * i + 1] = p.y; all[3 * i + 2] = p.z; }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = __popc(ballot);
  __syncthreads();
  int base = 0, total = 0;
  for (int w = 0; w < kBlock / 32; ++w) {
    if (w < warp) base += warp_sums[w];
    total += warp_sums[w];
  }
  if (pass == 0) {
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
    return;
  }
  int tile_base = 0;
  for (int t = 0; t < (int)blockIdx.x; ++t) tile_base += tile_counts[t];
  if (flag) {
    float* o = near + 3 * (size_t)(tile_base + base + __popc(ballot & ((1u << lane) - 1)));
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *near_count = tile_base + total;
}

}  // namespace

// ComputeLookupTableToApplyOdds (probability_values.cc:70-80), host float arithmetic (-ffp-contract=off).
void compute_odds_table(float probability, uint16_t* table) {
  auto clampf = [](float v, float lo, float hi) { return v > hi ? hi : (v < lo ? lo : v); };
  auto to_value = [&](float p) {
    const float lo = 0.1f, hi = 1.f - 0.1f;
    return (uint16_t)((int)lroundf((clampf(p, lo, hi) - lo) * (32766.f / (hi - lo))) + 1);
  };
  const float odds = probability / (1.f - probability);
  auto from_odds = [](float o) { return o / (o + 1.f); };
  table[0] = to_value(from_odds(odds)) + 32768;
  for (int cell = 1; cell != 32768; ++cell) {
    const float p = value_to_probability((uint16_t)cell);
    table[cell] = to_value(from_odds(odds * (p / (1.f - p)))) + 32768;
  }
}

int grid_reserve_pools(dl_grid* g, int pending_counter, int level);
int grid_ensure_device_state(dl_grid* g);

// One RangeDataInserter3D::Insert on the device grid. `d_returns` (n x 3 floats, grid frame) is device memory.
int grid_insert_device(dl_context* ctx, dl_grid* g, const Vec3f& origin, const float* d_returns, int n, int num_free,
                       const uint16_t* d_hit_table, const uint16_t* d_miss_table, int32_t* d_bbox, uint32_t* d_update_list) {
  if (n <= 0) return DL_OK;
  DL_TRY_STATUS(grid_ensure_device_state(g));
  InsertArgs a{};
  a.returns = d_returns; a.n = n; a.origin = origin; a.resolution = g->resolution; a.num_free = num_free;
  a.bbox = d_bbox; a.update_list = d_update_list; a.hit_table = d_hit_table; a.miss_table = d_miss_table;
  const int blocks = std::min(kNumSMs * 8, (n + kBlock - 1) / kBlock);
  // 1. which cells will be touched -> does the top level have to grow (CHECK_LE(new_bits, 8))?
  const int32_t init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, -0x7fffffff, -0x7fffffff, -0x7fffffff};
  DL_CUDA(ctx, cudaMemcpyAsync(d_bbox, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  ins_bbox_kernel<<<blocks, kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "ins_bbox_kernel");
  int32_t bbox[6];
  DL_CUDA(ctx, cudaMemcpyAsync(bbox, d_bbox, sizeof(bbox), cudaMemcpyDeviceToHost, ctx->stream));
  DL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (;;) {
    const int half = (64 << g->bits) >> 1;
    bool fits = true;
    for (int k = 0; k < 3; ++k) fits = fits && bbox[k] >= -half && bbox[3 + k] < half;
    if (fits) break;
    if (g->bits + 1 > 8) return ctx->fail(DL_ERR_GRID_RANGE, "cell index outside +-8192 cells");
    const size_t new_size = (size_t)8 << (3 * g->bits);
    int32_t* grown = nullptr;
    DL_CUDA(ctx, cudaMalloc((void**)&grown, new_size * sizeof(int32_t)));
    DL_CUDA(ctx, cudaMemsetAsync(grown, 0xFF, new_size * sizeof(int32_t), ctx->stream));
    const int old_cells = 1 << (3 * g->bits);
    grid_grow_kernel<<<(old_cells + 255) / 256, 256, 0, ctx->stream>>>(g->d_top, g->bits, grown);
    DL_LAUNCH_CHECK(ctx, "grid_grow_kernel");
    DL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DL_CUDA(ctx, cudaFree(g->d_top));
    g->d_top = grown;
    g->d_top_cap = new_size;
    g->bits += 1;
    // from here on the device copy is ahead of the host mirror, whatever happens next: a later failure (range, out of
    // memory) must not leave g->bits describing a host `top` of the old size
    g->mirror_stale = true;
    g->version++;
  }
  // 2. structure growth with exact reservations: mark + count the missing nodes, grow the node pool by that count, assign;
  //    then the same for the bricks (whose node entries now exist)
  a.bits = g->bits; a.top = g->d_top; a.nodes = g->d_nodes; a.bricks = g->d_bricks; a.counters = g->d_counters;
  DL_CUDA(ctx, cudaMemsetAsync(g->d_counters + 2, 0, 3 * sizeof(int32_t), ctx->stream));
  const int phases = num_free > 0 ? 2 : 1;
  for (int phase = 0; phase < phases; ++phase) {
    ins_claim_kernel<0, 0><<<blocks, kBlock, 0, ctx->stream>>>(a, phase);
    DL_LAUNCH_CHECK(ctx, "ins_claim_kernel<0,0>");
  }
  g->mirror_stale = true;  // top entries are marked: the device copy is the only consistent one until the Insert completes
  g->version++;
  DL_TRY_STATUS(grid_reserve_pools(g, 3, 0));
  a.nodes = g->d_nodes;
  for (int phase = 0; phase < phases; ++phase) {
    ins_claim_kernel<0, 1><<<blocks, kBlock, 0, ctx->stream>>>(a, phase);
    DL_LAUNCH_CHECK(ctx, "ins_claim_kernel<0,1>");
  }
  for (int phase = 0; phase < phases; ++phase) {
    ins_claim_kernel<1, 0><<<blocks, kBlock, 0, ctx->stream>>>(a, phase);
    DL_LAUNCH_CHECK(ctx, "ins_claim_kernel<1,0>");
  }
  DL_TRY_STATUS(grid_reserve_pools(g, 4, 1));
  a.bricks = g->d_bricks;
  for (int phase = 0; phase < phases; ++phase) {
    ins_claim_kernel<1, 1><<<blocks, kBlock, 0, ctx->stream>>>(a, phase);
    DL_LAUNCH_CHECK(ctx, "ins_claim_kernel<1,1>");
  }
  ins_apply_kernel<<<blocks, kBlock, 0, ctx->stream>>>(a, 0);
  DL_LAUNCH_CHECK(ctx, "ins_apply_kernel(hits)");
  if (num_free > 0) {
    ins_apply_kernel<<<blocks, kBlock, 0, ctx->stream>>>(a, 1);
    DL_LAUNCH_CHECK(ctx, "ins_apply_kernel(misses)");
  }
  ins_finish_kernel<<<blocks, kBlock, 0, ctx->stream>>>(a);
  DL_LAUNCH_CHECK(ctx, "ins_finish_kernel");
  g->mirror_stale = true;
  g->version++;
  return DL_OK;
}

int launch_transform_filter(dl_context* ctx, const float* in, int n, const Rigidf& to_submap, const Vec3f& origin_submap,
                            float max_range, float* all, float* near, int32_t* near_count, int32_t* tile_counts) {
  if (n <= 0) return DL_OK;
  const int tiles = (n + kBlock - 1) / kBlock;
  for (int pass = 0; pass < 2; ++pass) {
    transform_filter_kernel<<<tiles, kBlock, 0, ctx->stream>>>(in, n, to_submap, origin_submap, max_range, all, near,
                                                               near_count, tile_counts, pass);
    DL_LAUNCH_CHECK(ctx, "transform_filter_kernel");
  }
  return DL_OK;
}

}  // namespace dl
