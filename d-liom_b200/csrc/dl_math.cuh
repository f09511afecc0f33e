// Fixed-order float/double vector, quaternion and rigid-transform arithmetic shared by host and device code.
//
// The reference computes these with Eigen 3 on x86-64 (no FMA: /root/reference/src/cartographer/cmake/functions.cmake:75,92-95).
// To reproduce its voxel indices bit for bit every expression below has ONE evaluation order, compiled with
// -fmad=false (device) / -ffp-contract=off (host), IEEE division and square root. Orders follow Eigen 3.3's
// scalar paths: 3-term reductions are a0 + (a1 + a2), 4-term ones (a0 + a1) + (a2 + a3);
// q * v = (v + w * uv) + q.vec x uv with uv = 2 (q.vec x v).
// Reference interfaces: C/transform/rigid_transform.h:124-219, C/transform/transform.h:33-37,85-99,
// C/common/port.h:41-43 (RoundToInt = lround, ties away from zero).
#pragma once
#include <cstring>
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define DL_HD __host__ __device__ __forceinline__
#else
#define DL_HD inline
#endif

namespace dl {

template <typename T>
struct Vec3 {
  T x, y, z;
};
template <typename T>
struct Quat {  // w, x, y, z
  T w, x, y, z;
};
template <typename T>
struct Rigid {
  Vec3<T> t;
  Quat<T> q;
};
using Vec3f = Vec3<float>;
using Vec3d = Vec3<double>;
using Quatf = Quat<float>;
using Quatd = Quat<double>;
using Rigidf = Rigid<float>;
using Rigidd = Rigid<double>;

template <typename T>
DL_HD Vec3<T> add(const Vec3<T>& a, const Vec3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T>
DL_HD Vec3<T> sub(const Vec3<T>& a, const Vec3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T>
DL_HD Vec3<T> neg(const Vec3<T>& a) { return {-a.x, -a.y, -a.z}; }
template <typename T>
DL_HD Vec3<T> mul(T s, const Vec3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T>
DL_HD T dot3(const Vec3<T>& a, const Vec3<T>& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
template <typename T>
DL_HD Vec3<T> cross3(const Vec3<T>& a, const Vec3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
DL_HD float norm3(const Vec3f& a) { return sqrtf(dot3(a, a)); }
DL_HD double norm3(const Vec3d& a) { return sqrt(dot3(a, a)); }

template <typename T>
DL_HD Vec3<T> rotate(const Quat<T>& q, const Vec3<T>& v) {
  const Vec3<T> qv{q.x, q.y, q.z};
  Vec3<T> uv = cross3(qv, v);
  uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  const Vec3<T> c = cross3(qv, uv);
  return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}
template <typename T>
DL_HD Quat<T> qmul(const Quat<T>& a, const Quat<T>& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
template <typename T>
DL_HD Quat<T> qconj(const Quat<T>& q) { return {q.w, -q.x, -q.y, -q.z}; }
DL_HD Quatf qnormalized(const Quatf& q) {
  const float n = sqrtf((q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w));
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
DL_HD Quatd qnormalized(const Quatd& q) {
  const double n = sqrt((q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w));
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
// Rigid3 * Rigid3 re-normalises the rotation (rigid_transform.h:206-212).
template <typename T>
DL_HD Rigid<T> compose(const Rigid<T>& l, const Rigid<T>& r) {
  return {add(rotate(l.q, r.t), l.t), qnormalized(qmul(l.q, r.q))};
}
template <typename T>
DL_HD Vec3<T> apply(const Rigid<T>& r, const Vec3<T>& p) { return add(rotate(r.q, p), r.t); }
template <typename T>
DL_HD Rigid<T> inverse(const Rigid<T>& r) {
  const Quat<T> qi = qconj(r.q);
  return {neg(rotate(qi, r.t)), qi};
}
DL_HD Rigidf to_float(const Rigidd& r) {
  return {{(float)r.t.x, (float)r.t.y, (float)r.t.z}, {(float)r.q.w, (float)r.q.x, (float)r.q.y, (float)r.q.z}};
}
DL_HD Rigidd to_double(const Rigidf& r) {
  return {{(double)r.t.x, (double)r.t.y, (double)r.t.z}, {(double)r.q.w, (double)r.q.x, (double)r.q.y, (double)r.q.z}};
}
DL_HD Rigidd pose_from7(const double* p) { return {{p[0], p[1], p[2]}, {p[3], p[4], p[5], p[6]}}; }
DL_HD void pose_to7(const Rigidd& r, double* p) {
  p[0] = r.t.x; p[1] = r.t.y; p[2] = r.t.z; p[3] = r.q.w; p[4] = r.q.x; p[5] = r.q.y; p[6] = r.q.z;
}

// lround(x / resolution) per axis: voxel_filter.cc:126-131, hybrid_grid.h:430-435.
DL_HD int round_to_int(float x) { return (int)lroundf(x); }
struct Int3 {
  int x, y, z;
};
DL_HD Int3 cell_index(const Vec3f& p, float resolution) {
  return {round_to_int(p.x / resolution), round_to_int(p.y / resolution), round_to_int(p.z / resolution)};
}

// Same result as cell_index(p, resolution), cheaper: q~ = x * (1/resolution) is within 2^-22 |q| of the IEEE
// quotient fl(x / resolution); unless q~ lies that close to a rounding boundary k + 0.5 (where the two could round
// to different integers) lround(q~) IS lround(fl(x / resolution)). The rare near-boundary case takes the division.
struct CellDivider {
  float resolution, inverse;
};
DL_HD CellDivider make_divider(float resolution) { return {resolution, 1.0f / resolution}; }
// The rounding itself stays on the FMA pipe (no FRND / F2I, which issue at a quarter of the rate and were the hottest lines of
// every voxel kernel in ncu): adding 1.5 * 2^23 to 0 <= aq < 2^22 rounds aq to the nearest integer and leaves that integer in the
// low mantissa bits of the sum; away from a k + 0.5 boundary nearest-even IS lround.
DL_HD int round_div(float x, const CellDivider& d) {
  const float q = x * d.inverse;
  const float aq = fabsf(q);
  if (aq < 4194304.f) {
    const float t = aq + 12582912.f;          // bits: 0x4B400000 + nearest integer
    const float dist = fabsf(aq - (t - 12582912.f));  // exact distance to that integer, <= 0.5
    if (0.5f - dist > aq * 4.8e-7f + 1e-30f) {        // exact; == |frac(aq) - 0.5| of the boundary test
#ifdef __CUDA_ARCH__
      const int k = __float_as_int(t) - 0x4B400000;
#else
      int bits;
      memcpy(&bits, &t, sizeof(bits));
      const int k = bits - 0x4B400000;
#endif
      return q < 0.f ? -k : k;
    }
  }
  return round_to_int(x / d.resolution);
}
DL_HD Int3 cell_index(const Vec3f& p, const CellDivider& d) { return {round_div(p.x, d), round_div(p.y, d), round_div(p.z, d)}; }

// uint16 grid value -> probability: the expression that fills the reference's lookup table
// (probability_values.cc:27-68): value * kScale + (0.1f - kScale), unknown (0) -> 0.1f, marker bit ignored.
// Evaluated inline (two rounded float ops) it is bit-identical to a table read.
DL_HD float value_to_probability(uint16_t v) {
  const float kMin = 0.1f;
  const float kMax = 1.f - kMin;
  const float kScale = (kMax - kMin) / 32766.f;
  const int value = v & 32767;
  return value == 0 ? kMin : (float)value * kScale + (kMin - kScale);
}

}  // namespace dl
