// TEST INFRASTRUCTURE — CPU oracle. Not a product path: only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Small fixed-size vector / quaternion / rigid-transform arithmetic with ONE documented
// evaluation order per operation. The reference gets these from Eigen 3 (un-vendored,
// version unpinned, README "Prerequisites"); the orders below follow Eigen 3.3's
// non-vectorised paths (redux_novec_unroller: a 3-sum is a0 + (a1 + a2), a 4-sum is
// (a0 + a1) + (a2 + a3); QuaternionBase::_transformVector; generic quat_product).
// "Bit-exact" in this repo means: the CUDA path reproduces exactly these orders.
//
// Reference interfaces restated here:
//   transform::Rigid3<T>            C/transform/rigid_transform.h:124-219
//   AngleAxisVectorToRotationQuaternion / GetAngle   C/transform/transform.h:33-37,85-99
//   common::RoundToInt = std::lround                 C/common/port.h:41-43
// (C/ = /root/reference/src/cartographer/cartographer/)
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

template <typename T>
struct V3 {
  T x, y, z;
};
using V3f = V3<float>;
using V3d = V3<double>;

template <typename T>
inline V3<T> operator+(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T>
inline V3<T> operator-(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T>
inline V3<T> operator-(const V3<T>& a) { return {-a.x, -a.y, -a.z}; }
template <typename T, typename S>
inline V3<T> scale(const S& s, const V3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }

// Eigen redux order for 3 terms: a0 + (a1 + a2).
template <typename T>
inline T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
template <typename T>
inline T squared_norm(const V3<T>& a) { return dot(a, a); }
inline float norm(const V3f& a) { return std::sqrt(squared_norm(a)); }
inline double norm(const V3d& a) { return std::sqrt(squared_norm(a)); }

template <typename T>
inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Quaternion stored w, x, y, z (the reference's parameter-block order, ceres_pose.cc:23-28).
template <typename T>
struct Quat {
  T w, x, y, z;
};
using Quatf = Quat<float>;
using Quatd = Quat<double>;

template <typename T>
inline V3<T> vec(const Quat<T>& q) { return {q.x, q.y, q.z}; }

// q * v for a (possibly un-normalised) quaternion, Eigen's unit-quaternion formula:
//   uv = q.vec x v;  uv += uv;  result = (v + w*uv) + q.vec x uv
template <typename T>
inline V3<T> rotate(const Quat<T>& q, const V3<T>& v) {
  V3<T> uv = cross(vec(q), v);
  uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  const V3<T> c = cross(vec(q), uv);
  return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}

// Hamilton product, generic (non-SSE) Eigen order.
template <typename T>
inline Quat<T> qmul(const Quat<T>& a, const Quat<T>& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
          a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

template <typename T>
inline Quat<T> conj(const Quat<T>& q) { return {q.w, -q.x, -q.y, -q.z}; }

// coeffs are stored x,y,z,w in Eigen; 4-term redux = (x^2 + y^2) + (z^2 + w^2).
template <typename T>
inline T qsquared_norm(const Quat<T>& q) { return (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w); }
template <typename T>
inline T qdot(const Quat<T>& a, const Quat<T>& b) {
  return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
}
template <typename T>
inline Quat<T> qnormalized(const Quat<T>& q) {
  const T n = std::sqrt(qsquared_norm(q));
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}

template <typename T>
struct Rigid3 {
  V3<T> t{T(0), T(0), T(0)};
  Quat<T> q{T(1), T(0), T(0), T(0)};
};
using Rigid3f = Rigid3<float>;
using Rigid3d = Rigid3<double>;

// rigid_transform.h:206-212 — the composed rotation is re-normalised.
template <typename T>
inline Rigid3<T> compose(const Rigid3<T>& l, const Rigid3<T>& r) {
  return {rotate(l.q, r.t) + l.t, qnormalized(qmul(l.q, r.q))};
}
// rigid_transform.h:214-219
template <typename T>
inline V3<T> apply(const Rigid3<T>& r, const V3<T>& p) { return rotate(r.q, p) + r.t; }
// rigid_transform.h:155-159
template <typename T>
inline Rigid3<T> inverse(const Rigid3<T>& r) {
  const Quat<T> qi = conj(r.q);
  return {-rotate(qi, r.t), qi};
}
inline Rigid3f cast_f(const Rigid3d& r) {
  return {{(float)r.t.x, (float)r.t.y, (float)r.t.z}, {(float)r.q.w, (float)r.q.x, (float)r.q.y, (float)r.q.z}};
}
inline Rigid3d cast_d(const Rigid3f& r) {
  return {{(double)r.t.x, (double)r.t.y, (double)r.t.z}, {(double)r.q.w, (double)r.q.x, (double)r.q.y, (double)r.q.z}};
}

// transform.h:85-99, float instantiation. kCutoffAngle is a double constant, so the
// comparison promotes the float squared norm to double.
inline Quatf angle_axis_to_quat(const V3f& aa) {
  float s = 0.5f;
  float w = 1.f;
  const double kCutoff = 1e-8;
  if ((double)squared_norm(aa) > kCutoff) {
    const float n = norm(aa);
    // sin(norm / 2.) with float norm and double literal: computed in double, narrowed on assignment.
    s = (float)(std::sin((double)n / 2.) / (double)n);
    w = (float)std::cos((double)n / 2.);
  }
  return {w, s * aa.x, s * aa.y, s * aa.z};
}
inline Quatd angle_axis_to_quat(const V3d& aa) {
  double s = 0.5, w = 1.;
  if (squared_norm(aa) > 1e-8) {
    const double n = norm(aa);
    s = std::sin(n / 2.) / n;
    w = std::cos(n / 2.);
  }
  return {w, s * aa.x, s * aa.y, s * aa.z};
}

// transform.h:33-37
inline float rotation_angle(const Quatf& q) {
  return 2.f * std::atan2(norm(vec(q)), std::fabs(q.w));
}
inline double rotation_angle(const Quatd& q) {
  return 2. * std::atan2(norm(vec(q)), std::fabs(q.w));
}

inline int round_to_int(float x) { return (int)std::lround(x); }
inline int round_to_int(double x) { return (int)std::lround(x); }

struct I3 {
  int x, y, z;
};

// voxel_filter.cc:126-131 and hybrid_grid.h:430-435: float / float then lround, per axis.
inline I3 cell_index(const V3f& p, float resolution) {
  return {round_to_int(p.x / resolution), round_to_int(p.y / resolution), round_to_int(p.z / resolution)};
}

}  // namespace orc
