// RotationalScanMatcher::ComputeHistogram on the device (SURVEY 8a a13 / VERDICT "missing" 4): the histogram of horizontal
// directions between angular neighbours that LocalTrajectoryBuilder3D::InsertIntoSubmap stores with every inserted node
// (LTB:605-610) and the loop-closure matcher compares (SM/rotational_scan_matcher.cc:31-121, :159-170).
//
// The reference buckets the cloud into 0.2 m slices (std::map -> ascending slice order, points in input order), sorts each
// slice by the angle around its centroid and walks it sequentially, adding into ONE float histogram; every float sum here
// keeps that order:
//   1. key = (slice, input index) -> bitonic sort: slices contiguous, ascending, input order inside        (whole CTA)
//   2. centroid of a slice = sequential float sum over its points, one thread per slice                     (:54-61)
//   3. key = (slice, atan2 of the offset from the centroid, index); points closer than 0.2 m dropped -> sort (:94-121)
//   4. one thread per slice walks its sorted points with the reference's `last`-point logic and emits one
//      (bucket, value) event per point                                                                      (:63-92)
//   5. one thread per bucket adds its events in (slice, point) order = the order of the reference's += chain (:31-52)
// One CTA per cloud; the sorts run in global memory (the arrays are L2-resident: 8 B per point).
// Float parity: the sums are ordered like the reference's, but atan2f / sqrtf are the device's (atan2f <= 2 ulp), so a point
// whose angle sits on a bucket or ordering boundary can land differently than on the CPU: compared with a tolerance in the
// tests, not bit for bit. Compiled -fmad=false like everything that mirrors the reference's float expressions.
#include "dl_internal.cuh"

namespace dl {
namespace {

constexpr int kThreads = 1024;
constexpr unsigned long long kPad = 0xFFFFFFFFFFFFFFFFull;

struct HistogramArgs {
  const float* points;  // n x 3, already rotated into the gravity-aligned frame
  int n, np2;           // np2 = n rounded up to a power of two
  int size;             // histogram buckets
  unsigned long long* keys;   // np2
  int* slice_first;     // n + 1: first sorted position of each distinct slice (compact list), then the end
  float* centroid;      // 2 per distinct slice
  int* ev_bucket;       // n
  float* ev_value;      // n
  int* counters;        // [0] number of distinct slices, [1] error
  float* histogram;     // size
};

__device__ __forceinline__ unsigned order_bits(float f) {  // monotone map float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ void bitonic_sort(unsigned long long* keys, int np2) {
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += kThreads) {
        const int partner = i ^ j;
        if (partner > i) {
          const unsigned long long a = keys[i], b = keys[partner];
          const bool ascending = (i & k) == 0;
          if ((a > b) == ascending) {
            keys[i] = b;
            keys[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// boundaries of the runs of equal `slice` (top 20 bits of the key) in the sorted array -> compact, ordered list of run starts.
// Every thread scans one contiguous chunk; a block-wide exclusive scan of the per-chunk run counts places them in order.
__device__ void find_slices(const HistogramArgs& a, int count) {
  __shared__ int warp_sums[kThreads / 32];
  __shared__ int total_s;
  const int chunk = (count + kThreads - 1) / kThreads;
  const int begin = min(count, (int)threadIdx.x * chunk), end = min(count, begin + chunk);
  auto starts_run = [&](int i) { return i == 0 || (a.keys[i] >> 44) != (a.keys[i - 1] >> 44); };
  int mine = 0;
  for (int i = begin; i < end; ++i) mine += starts_run(i);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int v = warp_sums[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += o;
    }
    warp_sums[lane] = v;  // inclusive over warps
    if (lane == 31) total_s = v;
  }
  __syncthreads();
  int pos = inc - mine + (warp ? warp_sums[warp - 1] : 0);
  for (int i = begin; i < end; ++i)
    if (starts_run(i)) a.slice_first[pos++] = i;
  if (threadIdx.x == 0) {
    a.slice_first[total_s] = count;
    a.counters[0] = total_s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) rotational_histogram_kernel(HistogramArgs a) {
  const float kSliceHeight = 0.2f, kMinDistance = 0.2f, kMaxDistance = 0.9f;
  const float kPi = 3.14159274101257324f;  // (float)M_PI
  // ---- 1. (slice, index) keys
  for (int i = threadIdx.x; i < a.np2; i += kThreads) {
    unsigned long long key = kPad;
    if (i < a.n) {
      const int slice = round_to_int(a.points[3 * i + 2] / kSliceHeight);
      if (slice < -(1 << 19) || slice >= (1 << 19)) a.counters[1] = 1;
      key = ((unsigned long long)(unsigned)(slice + (1 << 19)) << 44) | (unsigned long long)i;
    }
    a.keys[i] = key;
  }
  if (threadIdx.x < a.size) a.histogram[threadIdx.x] = 0.f;
  __syncthreads();
  bitonic_sort(a.keys, a.np2);
  find_slices(a, a.n);
  const int num_slices = a.counters[0];
  // ---- 2. centroids (sequential float sums in input order, per slice)
  for (int s = threadIdx.x; s < num_slices; s += kThreads) {
    float sx = 0.f, sy = 0.f;
    const int b = a.slice_first[s], e = a.slice_first[s + 1];
    for (int k = b; k < e; ++k) {
      const int i = (int)(a.keys[k] & 0xFFFFFFFFFFFull);
      sx += a.points[3 * i];
      sy += a.points[3 * i + 1];
    }
    const float cnt = (float)(e - b);
    a.centroid[2 * s] = sx / cnt;
    a.centroid[2 * s + 1] = sy / cnt;
  }
  __syncthreads();
  // ---- 3. (slice, angle, index) keys of the points far enough from their slice's centroid
  for (int s = 0; s < num_slices; ++s) {
    const int b = a.slice_first[s], e = a.slice_first[s + 1];
    const float cx = a.centroid[2 * s], cy = a.centroid[2 * s + 1];
    for (int k = b + threadIdx.x; k < e; k += kThreads) {
      const unsigned long long key = a.keys[k];
      const int i = (int)(key & 0xFFFFFFFFFFFull);
      const float dx = a.points[3 * i] - cx, dy = a.points[3 * i + 1] - cy;
      unsigned long long out = kPad;
      if (!(sqrtf(dx * dx + dy * dy) < kMinDistance))
        out = ((unsigned long long)s << 44) | ((unsigned long long)(order_bits(atan2f(dy, dx)) >> 8) << 20) | (unsigned long long)(k - b);
      // 20 bits slice ordinal | 24 bits angle | 20 bits position inside the slice (ties and the lost low angle bits resolve
      // towards input order; std::sort leaves that order unspecified)
      a.keys[k] = out;
      a.ev_bucket[k] = i;  // position in the first sort -> point index (the second key carries the position inside the slice)
    }
  }
  __syncthreads();
  // 24 angle bits cannot order nearly equal angles: runs that tie on them are fixed up below with the exact float compare
  bitonic_sort(a.keys, a.np2);
  // ---- 4. per slice: sequential walk in sorted order
  for (int s = threadIdx.x; s < num_slices; s += kThreads) {
    const int b0 = a.slice_first[s];
    // sorted segment of slice s: find it by binary search on the leading 20 bits
    int lo = 0, hi = a.n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((a.keys[mid] >> 44) < (unsigned long long)s) lo = mid + 1; else hi = mid;
    }
    const int begin = lo;
    hi = a.n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((a.keys[mid] >> 44) <= (unsigned long long)s) lo = mid + 1; else hi = mid;
    }
    const int end = lo;
    if (begin >= end) continue;
    const float cx = a.centroid[2 * s], cy = a.centroid[2 * s + 1];
    auto point_of = [&](int k) { return a.ev_bucket[b0 + (int)(a.keys[k] & 0xFFFFFull)]; };
    // insertion-sort fix-up inside runs whose 24-bit angle prefix ties: exact float angle, then input order
    // (runs are short: neighbouring points of one slice rarely share 24 leading angle bits)
    for (int k = begin + 1; k < end; ++k) {
      int m = k;
      while (m > begin && ((a.keys[m] >> 20) == (a.keys[m - 1] >> 20))) {
        const int im = point_of(m), ip = point_of(m - 1);
        const float am = atan2f(a.points[3 * im + 1] - cy, a.points[3 * im] - cx);
        const float ap = atan2f(a.points[3 * ip + 1] - cy, a.points[3 * ip] - cx);
        if (am < ap) {
          const unsigned long long t = a.keys[m];
          a.keys[m] = a.keys[m - 1];
          a.keys[m - 1] = t;
          --m;
        } else {
          break;
        }
      }
    }
    int il = point_of(begin);
    float lx = a.points[3 * il], ly = a.points[3 * il + 1];
    for (int k = begin; k < end; ++k) {
      const int i = point_of(k);
      const float px = a.points[3 * i], py = a.points[3 * i + 1];
      const float dx = px - lx, dy = py - ly;
      const float ccx = px - cx, ccy = py - cy;
      const float distance = sqrtf(dx * dx + dy * dy);
      const float direction_norm = sqrtf(ccx * ccx + ccy * ccy);
      int bucket = -1;
      float value = 0.f;
      if (!(distance < kMinDistance || direction_norm < kMinDistance)) {
        if (distance > kMaxDistance) {
          lx = px; ly = py;
        } else {
          float angle = atan2f(dy, dx);
          const float dot = (dx / distance) * (ccx / direction_norm) + (dy / distance) * (ccy / direction_norm);
          value = fmaxf(0.f, 1.f - fabsf(dot));
          while (angle > kPi) angle -= kPi;
          while (angle < 0.f) angle += kPi;
          const float zero_to_one = angle / kPi;
          bucket = min(max(round_to_int((float)a.size * zero_to_one - 0.5f), 0), a.size - 1);
        }
      }
      // events are stored at the sorted position: (slice, point) order = array order
      a.ev_value[k] = value;
      reinterpret_cast<int*>(a.centroid + 2 * a.n)[k] = bucket;  // second half of the centroid buffer: bucket per sorted position
    }
  }
  __syncthreads();
  // ---- 5. per bucket: sequential float sum in (slice, point) order
  const int* ev_b = reinterpret_cast<const int*>(a.centroid + 2 * a.n);
  // number of kept points = first padded key
  __shared__ int kept;
  if (threadIdx.x == 0) {
    int lo = 0, hi = a.n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (a.keys[mid] != kPad) lo = mid + 1; else hi = mid;
    }
    kept = lo;
  }
  __syncthreads();
  if (threadIdx.x < a.size) {
    float sum = 0.f;
    for (int k = 0; k < kept; ++k)
      if (ev_b[k] == (int)threadIdx.x) sum += a.ev_value[k];
    a.histogram[threadIdx.x] = sum;
  }
}

}  // namespace

size_t rotational_histogram_scratch_bytes(int64_t n) {
  int64_t np2 = 64;
  while (np2 < n) np2 <<= 1;
  return arena_bytes({(size_t)np2 * 8, (size_t)(n + 1) * 4, (size_t)n * 16, (size_t)n * 4, (size_t)n * 4, 64, 1024 * 4});
}

// d_points: n x 3 floats on the device. d_histogram: `size` floats (size <= 1024). Scratch is carved from `a`.
int launch_rotational_histogram(dl_context* ctx, Arena& a, const float* d_points, int64_t n, int size, float* d_histogram,
                                int32_t** d_error_out) {
  if (size < 1 || size > kThreads) return ctx->fail(DL_ERR_ARG, "histogram size must be in [1, 1024]");
  if (n > (1 << 20)) return ctx->fail(DL_ERR_ARG, "more than 2^20 points in a rotational histogram");
  HistogramArgs h{};
  h.points = d_points;
  h.n = (int)n;
  int np2 = 64;
  while (np2 < n) np2 <<= 1;
  h.np2 = np2;
  h.size = size;
  h.keys = a.take<unsigned long long>(np2);
  h.slice_first = a.take<int>(n + 1);
  h.centroid = a.take<float>(4 * (size_t)n + 4);  // 2 n centroid floats (upper bound) + n bucket ints
  h.ev_bucket = a.take<int>(n + 1);
  h.ev_value = a.take<float>(n + 1);
  h.counters = a.take<int>(2);
  h.histogram = d_histogram;
  DL_CUDA(ctx, cudaMemsetAsync(h.counters, 0, 2 * sizeof(int), ctx->stream));
  if (n == 0) {
    DL_CUDA(ctx, cudaMemsetAsync(d_histogram, 0, sizeof(float) * size, ctx->stream));
  } else {
    rotational_histogram_kernel<<<1, kThreads, 0, ctx->stream>>>(h);
    DL_LAUNCH_CHECK(ctx, "rotational_histogram_kernel");
  }
  if (d_error_out) *d_error_out = h.counters + 1;
  return DL_OK;
}

}  // namespace dl
