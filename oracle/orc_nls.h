// TEST INFRASTRUCTURE — CPU oracle (see orc_math.h header).
//
// Point-to-probability-grid nonlinear least squares, restated from
//   SM/interpolated_grid.h:50-146            (tricubic-smoothstep interpolation, float cell selection)
//   SM/occupied_space_cost_function_3d.h:46-80 (r_i = s * (1 - M(T p_i)), q NOT normalised)
//   SM/translation_delta_cost_functor_3d.h:38-44, SM/rotation_delta_cost_functor_3d.h:42-53
//   C/common/math.h:74-81 (QuaternionProduct), C/mapping/internal/3d/rotation_parameterization.h:27-39
//   SM/ceres_scan_matcher_3d.cc:63-123       (problem assembly, DENSE_QR, weights / sqrt(N))
//   (SM/ = C/mapping/internal/3d/scan_matching/)
// and from the THIRD-PARTY solver the reference links but does not vendor: Ceres-Solver 1.13.0
// (pinned by /root/reference/src/cartographer/scripts/install_ceres.sh:20): forward-mode Jet
// autodiff, QuaternionParameterization, TrustRegionMinimizer + LevenbergMarquardtStrategy +
// DenseQRSolver (Eigen householderQr) with Solver::Options defaults. That part is restated from
// the published algorithm (RECALLED — not checkable against /root/reference); the parity anchor is
// the reference's own test ceres_scan_matcher_3d_test.cc:34-116 plus a scipy cross-check of the optimum.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include "orc_grid.h"
#include "orc_math.h"

namespace orc {

// ---------------------------------------------------------------- Jet<double, 7>
constexpr int kJetN = 7;
struct Jet {
  double a;
  double v[kJetN];
  Jet() : a(0) { for (double& d : v) d = 0; }
  Jet(double s) : a(s) { for (double& d : v) d = 0; }  // NOLINT: implicit like ceres::Jet
  Jet(double s, int k) : a(s) { for (double& d : v) d = 0; v[k] = 1; }
};
inline Jet operator+(const Jet& f, const Jet& g) { Jet h; h.a = f.a + g.a; for (int i = 0; i < kJetN; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
inline Jet operator-(const Jet& f, const Jet& g) { Jet h; h.a = f.a - g.a; for (int i = 0; i < kJetN; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
inline Jet operator-(const Jet& f) { Jet h; h.a = -f.a; for (int i = 0; i < kJetN; ++i) h.v[i] = -f.v[i]; return h; }
inline Jet operator+(const Jet& f, double s) { Jet h = f; h.a = f.a + s; return h; }
inline Jet operator+(double s, const Jet& f) { Jet h = f; h.a = s + f.a; return h; }
inline Jet operator-(const Jet& f, double s) { Jet h = f; h.a = f.a - s; return h; }
inline Jet operator-(double s, const Jet& f) { Jet h; h.a = s - f.a; for (int i = 0; i < kJetN; ++i) h.v[i] = -f.v[i]; return h; }
inline Jet operator*(const Jet& f, const Jet& g) { Jet h; h.a = f.a * g.a; for (int i = 0; i < kJetN; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
inline Jet operator*(const Jet& f, double s) { Jet h; h.a = f.a * s; for (int i = 0; i < kJetN; ++i) h.v[i] = f.v[i] * s; return h; }
inline Jet operator*(double s, const Jet& f) { Jet h; h.a = f.a * s; for (int i = 0; i < kJetN; ++i) h.v[i] = f.v[i] * s; return h; }
// Jet / scalar multiplies by the inverse (ceres/jet.h).
inline Jet operator/(const Jet& f, double s) { const double si = 1.0 / s; Jet h; h.a = f.a * si; for (int i = 0; i < kJetN; ++i) h.v[i] = f.v[i] * si; return h; }
inline Jet& operator+=(Jet& f, const Jet& g) { f = f + g; return f; }
inline double scalar_part(double x) { return x; }
inline double scalar_part(const Jet& x) { return x.a; }

// ---------------------------------------------------------------- interpolation
class InterpolatedGrid {
 public:
  explicit InterpolatedGrid(const HybridGrid& g) : g_(g) {}

  template <typename T>
  T GetProbability(const T& x, const T& y, const T& z) const {
    const double sx = scalar_part(x), sy = scalar_part(y), sz = scalar_part(z);
    // CenterOfLowerVoxel (interpolated_grid.h:120-139): containing cell from the float-narrowed
    // point, centre in float, compared against the DOUBLE coordinate.
    V3f c = g_.GetCenterOfCell(g_.GetCellIndex(V3f{(float)sx, (float)sy, (float)sz}));
    if ((double)c.x > sx) c.x -= g_.resolution();
    if ((double)c.y > sy) c.y -= g_.resolution();
    if ((double)c.z > sz) c.z -= g_.resolution();
    const double x1 = c.x, y1 = c.y, z1 = c.z;
    // float + float, widened afterwards (:112-117)
    const double x2 = c.x + g_.resolution(), y2 = c.y + g_.resolution(), z2 = c.z + g_.resolution();

    const I3 i1 = g_.GetCellIndex(V3f{(float)x1, (float)y1, (float)z1});
    auto q = [&](int dx, int dy, int dz) -> double { return g_.GetProbability(I3{i1.x + dx, i1.y + dy, i1.z + dz}); };
    const double q111 = q(0, 0, 0), q112 = q(0, 0, 1), q121 = q(0, 1, 0), q122 = q(0, 1, 1);
    const double q211 = q(1, 0, 0), q212 = q(1, 0, 1), q221 = q(1, 1, 0), q222 = q(1, 1, 1);

    const T nx = (x - x1) / (x2 - x1);
    const T ny = (y - y1) / (y2 - y1);
    const T nz = (z - z1) / (z2 - z1);
    const T nxx = nx * nx, nxxx = nx * nxx;
    const T nyy = ny * ny, nyyy = ny * nyy;
    const T nzz = nz * nz, nzzz = nz * nzz;

    const T q11 = (q111 - q112) * nzzz * 2. + (q112 - q111) * nzz * 3. + q111;
    const T q12 = (q121 - q122) * nzzz * 2. + (q122 - q121) * nzz * 3. + q121;
    const T q21 = (q211 - q212) * nzzz * 2. + (q212 - q211) * nzz * 3. + q211;
    const T q22 = (q221 - q222) * nzzz * 2. + (q222 - q221) * nzz * 3. + q221;
    const T q1 = (q11 - q12) * nyyy * 2. + (q12 - q11) * nyy * 3. + q11;
    const T q2 = (q21 - q22) * nyyy * 2. + (q22 - q21) * nyy * 3. + q21;
    return (q1 - q2) * nxxx * 2. + (q2 - q1) * nxx * 3. + q1;
  }

 private:
  const HybridGrid& g_;
};

// ---------------------------------------------------------------- problem
struct CloudAndGrid {
  const float* pts;  // 3 floats per point
  int64_t n;
  const HybridGrid* grid;
};

struct CeresMatcherOptions {
  std::vector<double> occupied_space_weight;
  double translation_weight = 0;
  double rotation_weight = 0;
  bool only_optimize_yaw = false;
  bool use_nonmonotonic_steps = false;
  int max_num_iterations = 12;
};

enum Termination { kConvergence = 0, kNoConvergence = 1, kFailure = 2 };

struct IterationLog {
  int iteration;
  double cost, cost_change, gradient_max_norm, step_norm, relative_decrease, trust_region_radius;
  bool step_is_valid, step_is_successful;
};

struct SolveSummary {
  double initial_cost = 0, final_cost = 0;
  int num_successful_steps = 0, num_unsuccessful_steps = 0;
  int termination = kNoConvergence;
  std::string message;
  std::vector<IterationLog> iterations;
  int num_residual_evaluations = 0, num_jacobian_evaluations = 0;
};

class ScanMatchProblem {
 public:
  ScanMatchProblem(const CeresMatcherOptions& opt, const V3d& target_translation, const Rigid3d& initial,
                   const std::vector<CloudAndGrid>& pairs)
      : opt_(opt), target_t_(target_translation), pairs_(pairs) {
    for (size_t i = 0; i < pairs.size(); ++i) {
      scaling_.push_back(opt.occupied_space_weight[i] / std::sqrt((double)pairs[i].n));
      num_residuals_ += (int)pairs[i].n;
    }
    if (opt.translation_weight > 0.) num_residuals_ += 3;
    if (opt.rotation_weight > 0.) num_residuals_ += 3;
    target_q_inv_ = {initial.q.w, -initial.q.x, -initial.q.y, -initial.q.z};
  }

  int num_residuals() const { return num_residuals_; }
  int num_local() const { return opt_.only_optimize_yaw ? 4 : 6; }
  int num_ambient() const { return 7; }

  // x = [t(3), q(4) wxyz]. Fills residuals; if J != nullptr also the row-major
  // num_residuals x num_local Jacobian in the local parameterisation.
  void Evaluate(const double* x, double* residuals, double* J) const {
    const int nl = num_local();
    int row = 0;
    // local-parameterisation Jacobian of the quaternion block (4 x local_q)
    double plus_jac[4][3];
    int local_q = 3;
    if (J) {
      if (!opt_.only_optimize_yaw) {
        // ceres::QuaternionParameterization::ComputeJacobian, x = w,x,y,z
        const double q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
        const double pj[4][3] = {{-q1, -q2, -q3}, {q0, q3, -q2}, {-q3, q0, q1}, {q2, -q1, q0}};
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) plus_jac[i][j] = pj[i][j];
      } else {
        // AutoDiffLocalParameterization<YawOnlyQuaternionPlus,4,1>: d/d(delta) at delta = 0 of
        // q_delta (x) x with q_delta = (sqrt(1 - d^2), 0, 0, d): dq_delta/dd = (0,0,0,1).
        local_q = 1;
        const double q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
        // (0,0,0,1) (x) (q0,q1,q2,q3) = (-q3, -q2, q1, q0)
        plus_jac[0][0] = -q3; plus_jac[1][0] = -q2; plus_jac[2][0] = q1; plus_jac[3][0] = q0;
      }
    }
    auto write_row = [&](int r, const double* amb /*7*/) {
      double* out = J + (size_t)r * nl;
      out[0] = amb[0]; out[1] = amb[1]; out[2] = amb[2];
      for (int j = 0; j < local_q; ++j) {
        double s = 0;
        for (int i = 0; i < 4; ++i) s += amb[3 + i] * plus_jac[i][j];
        out[3 + j] = s;
      }
    };

    for (size_t k = 0; k < pairs_.size(); ++k) {
      const InterpolatedGrid ig(*pairs_[k].grid);
      const double s = scaling_[k];
      for (int64_t i = 0; i < pairs_[k].n; ++i, ++row) {
        const float* p = pairs_[k].pts + 3 * i;
        if (!J) {
          const V3<double> t{x[0], x[1], x[2]};
          const Quat<double> q{x[3], x[4], x[5], x[6]};
          const V3<double> w = rotate(q, V3<double>{(double)p[0], (double)p[1], (double)p[2]}) + t;
          const double prob = ig.GetProbability(w.x, w.y, w.z);
          residuals[row] = s * (1. - prob);
        } else {
          const V3<Jet> t{Jet(x[0], 0), Jet(x[1], 1), Jet(x[2], 2)};
          const Quat<Jet> q{Jet(x[3], 3), Jet(x[4], 4), Jet(x[5], 5), Jet(x[6], 6)};
          const V3<Jet> w = rotate(q, V3<Jet>{Jet((double)p[0]), Jet((double)p[1]), Jet((double)p[2])}) + t;
          const Jet prob = ig.GetProbability(w.x, w.y, w.z);
          const Jet r = s * (1. - prob);
          residuals[row] = r.a;
          write_row(row, r.v);
        }
      }
    }
    if (opt_.translation_weight > 0.) {
      const double s = opt_.translation_weight;
      const double tt[3] = {target_t_.x, target_t_.y, target_t_.z};
      for (int a = 0; a < 3; ++a, ++row) {
        residuals[row] = s * (x[a] - tt[a]);
        if (J) {
          double amb[7] = {0, 0, 0, 0, 0, 0, 0};
          amb[a] = s;
          write_row(row, amb);
        }
      }
    }
    if (opt_.rotation_weight > 0.) {
      const double s = opt_.rotation_weight;
      const Quat<double>& z = target_q_inv_;
      const double* w = x + 3;
      // math.h:74-81 with z = target^-1, w = current
      const double d1 = z.w * w[1] + z.x * w[0] + z.y * w[3] - z.z * w[2];
      const double d2 = z.w * w[2] - z.x * w[3] + z.y * w[0] + z.z * w[1];
      const double d3 = z.w * w[3] + z.x * w[2] - z.y * w[1] + z.z * w[0];
      const double d[3] = {d1, d2, d3};
      const double dd[3][4] = {{z.x, z.w, -z.z, z.y}, {z.y, z.z, z.w, -z.x}, {z.z, -z.y, z.x, z.w}};
      for (int a = 0; a < 3; ++a, ++row) {
        residuals[row] = s * d[a];
        if (J) {
          double amb[7] = {0, 0, 0, s * dd[a][0], s * dd[a][1], s * dd[a][2], s * dd[a][3]};
          write_row(row, amb);
        }
      }
    }
  }

  // x (+) delta, ambient 7 <- local.
  void Plus(const double* x, const double* delta, double* out) const {
    out[0] = x[0] + delta[0]; out[1] = x[1] + delta[1]; out[2] = x[2] + delta[2];
    const Quat<double> q{x[3], x[4], x[5], x[6]};
    Quat<double> r = q;
    if (!opt_.only_optimize_yaw) {
      // ceres::QuaternionParameterization::Plus
      const double n = std::sqrt(delta[3] * delta[3] + delta[4] * delta[4] + delta[5] * delta[5]);
      if (n > 0.0) {
        const double sbd = std::sin(n) / n;
        r = qmul(Quat<double>{std::cos(n), sbd * delta[3], sbd * delta[4], sbd * delta[5]}, q);
      }
    } else {
      double d = delta[3];
      if (d > 0.5) d = 0.5;
      if (d < -0.5) d = -0.5;
      r = qmul(Quat<double>{std::sqrt(1. - d * d), 0., 0., d}, q);
    }
    out[3] = r.w; out[4] = r.x; out[5] = r.y; out[6] = r.z;
  }

 private:
  CeresMatcherOptions opt_;
  V3d target_t_;
  Quat<double> target_q_inv_;
  std::vector<CloudAndGrid> pairs_;
  std::vector<double> scaling_;
  int num_residuals_ = 0;
};

// ---------------------------------------------------------------- dense QR least squares
// Solves min || A y - b || for a tall row-major m x n matrix by Householder QR
// (what Eigen's householderQr().solve() computes). A and b are overwritten.
inline void householder_qr_solve(double* A, double* b, int m, int n, double* y) {
  for (int k = 0; k < n; ++k) {
    double tail = 0;
    for (int i = k + 1; i < m; ++i) tail += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    const double c0 = A[(size_t)k * n + k];
    double beta, tau;
    if (tail <= std::numeric_limits<double>::min()) {
      tau = 0;
      beta = c0;
    } else {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < m; ++i) A[(size_t)i * n + k] /= (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[(size_t)k * n + k] = beta;
    if (tau != 0) {
      // apply H = I - tau v v^T (v_k = 1) to the remaining columns and to b
      for (int j = k + 1; j <= n; ++j) {
        auto col = [&](int i) -> double& { return j < n ? A[(size_t)i * n + j] : b[i]; };
        double s = col(k);
        for (int i = k + 1; i < m; ++i) s += A[(size_t)i * n + k] * col(i);
        s *= tau;
        col(k) -= s;
        for (int i = k + 1; i < m; ++i) col(i) -= s * A[(size_t)i * n + k];
      }
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[(size_t)k * n + j] * y[j];
    y[k] = s / A[(size_t)k * n + k];
  }
}

// ---------------------------------------------------------------- trust-region LM (Ceres 1.13 defaults)
struct LmConstants {
  double initial_trust_region_radius = 1e4;
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  int max_num_consecutive_invalid_steps = 5;
  int max_consecutive_nonmonotonic_steps = 5;
};

// Dense Cholesky solve of the symmetric positive definite system A y = b (A row-major n x n, destroyed). False if a pivot is
// not positive. The linear algebra of ceres' *_NORMAL_CHOLESKY solvers, without their sparsity.
inline bool cholesky_solve_dense(double* A, const double* b, int n, double* y) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  std::vector<double> z(n);
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * z[k];
    z[i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * y[k];
    y[i] = s / A[(size_t)i * n + i];
  }
  return true;
}

enum LinearSolver { kDenseQr = 0, kNormalCholesky = 1 };

template <typename Problem>
inline void solve_trust_region(const Problem& problem, bool use_nonmonotonic_steps, int max_num_iterations,
                               double* parameters /*ambient size, in-out*/, SolveSummary* summary,
                               LinearSolver linear_solver = kDenseQr) {
  const LmConstants c;
  const int m = problem.num_residuals();
  const int n = problem.num_local();
  const int na = problem.num_ambient();
  std::vector<double> x(parameters, parameters + na), cand(na);
  std::vector<double> res(m), J((size_t)m * n), Jaug((size_t)(m + n) * n), rhs(m + n);
  std::vector<double> scale(n), diag(n), lmdiag(n), g(n), step(n), delta(n), model(m);
  double x_cost = 0, x_norm = 0, minimum_cost = std::numeric_limits<double>::max();
  double radius = c.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_consecutive_invalid = 0;

  IterationLog it{};
  auto norm7 = [na](const double* a) { double s = 0; for (int i = 0; i < na; ++i) s += a[i] * a[i]; return std::sqrt(s); };

  auto evaluate_gradient_and_jacobian = [&]() {
    problem.Evaluate(x.data(), res.data(), J.data());
    summary->num_residual_evaluations++;
    summary->num_jacobian_evaluations++;
    double cs = 0;
    for (int i = 0; i < m; ++i) cs += res[i] * res[i];
    x_cost = 0.5 * cs;
    for (int j = 0; j < n; ++j) g[j] = 0;
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) g[j] += J[(size_t)i * n + j] * res[i];
    if (it.iteration == 0) {
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += J[(size_t)i * n + j] * J[(size_t)i * n + j];
        scale[j] = 1.0 / (1.0 + std::sqrt(s));
      }
    }
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) J[(size_t)i * n + j] *= scale[j];
    // projected gradient: x - Plus(x, -g), ambient max-norm
    std::vector<double> ng(n), px(na);
    for (int j = 0; j < n; ++j) ng[j] = -g[j];
    problem.Plus(x.data(), ng.data(), px.data());
    double mx = 0;
    for (int i = 0; i < na; ++i) mx = std::max(mx, std::fabs(x[i] - px[i]));
    it.gradient_max_norm = mx;
  };

  // nonmonotonic step evaluator (trust_region_step_evaluator.cc)
  const int max_nonmono = use_nonmonotonic_steps ? c.max_consecutive_nonmonotonic_steps : 0;
  double ev_minimum, ev_current, ev_reference, ev_candidate, ev_acc_ref = 0, ev_acc_cand = 0;
  int ev_num_nonmono = 0;

  // ---- iteration zero
  it = IterationLog{};
  it.iteration = 0;
  x_norm = norm7(x.data());
  evaluate_gradient_and_jacobian();
  summary->initial_cost = x_cost;
  it.cost = x_cost;
  it.step_is_valid = true;
  it.step_is_successful = true;
  ev_minimum = ev_current = ev_reference = ev_candidate = x_cost;

  auto finish = [&](int term, const char* msg) {
    summary->termination = term;
    summary->message = msg;
  };

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it.step_is_successful) {
      summary->num_successful_steps++;
      if (x_cost < minimum_cost) {
        minimum_cost = x_cost;
        for (int i = 0; i < na; ++i) parameters[i] = x[i];
      }
    } else {
      summary->num_unsuccessful_steps++;
    }
    it.trust_region_radius = radius;
    summary->iterations.push_back(it);
    if (it.iteration >= max_num_iterations) { finish(kNoConvergence, "max iterations"); break; }
    if (it.step_is_successful && it.gradient_max_norm <= c.gradient_tolerance) { finish(kConvergence, "gradient tolerance"); break; }
    if (radius <= c.min_trust_region_radius) { finish(kConvergence, "min trust region radius"); break; }

    const int iteration = it.iteration + 1;
    it = IterationLog{};
    it.iteration = iteration;
    it.gradient_max_norm = summary->iterations.back().gradient_max_norm;

    // ---- ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep + DenseQRSolver)
    if (!reuse_diagonal) {
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += J[(size_t)i * n + j] * J[(size_t)i * n + j];
        diag[j] = std::min(std::max(s, c.min_lm_diagonal), c.max_lm_diagonal);
      }
    }
    for (int j = 0; j < n; ++j) lmdiag[j] = std::sqrt(diag[j] / radius);
    bool step_ok = true;
    if (linear_solver == kNormalCholesky) {
      // (J^T J + D^T D) y = J^T r, the system the *_NORMAL_CHOLESKY solvers factor (J is the column-scaled Jacobian)
      std::vector<double> H((size_t)n * n, 0.0), b(n, 0.0);
      for (int i = 0; i < m; ++i) {
        const double* row = J.data() + (size_t)i * n;
        for (int a = 0; a < n; ++a) {
          if (row[a] == 0.0) continue;
          b[a] += row[a] * res[i];
          for (int c2 = 0; c2 <= a; ++c2) H[(size_t)a * n + c2] += row[a] * row[c2];
        }
      }
      for (int a = 0; a < n; ++a) H[(size_t)a * n + a] += lmdiag[a] * lmdiag[a];
      step_ok = cholesky_solve_dense(H.data(), b.data(), n, step.data());
    } else {
      std::copy(J.begin(), J.end(), Jaug.begin());
      for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Jaug[(size_t)(m + i) * n + j] = (i == j) ? lmdiag[j] : 0.0;
      for (int i = 0; i < m; ++i) rhs[i] = res[i];
      for (int i = 0; i < n; ++i) rhs[m + i] = 0;
      householder_qr_solve(Jaug.data(), rhs.data(), m + n, n, step.data());
    }
    for (int j = 0; j < n; ++j) { if (!std::isfinite(step[j])) step_ok = false; step[j] = -step[j]; }
    reuse_diagonal = true;

    double model_cost_change = 0;
    if (step_ok) {
      for (int i = 0; i < m; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += J[(size_t)i * n + j] * step[j];
        model[i] = s;
      }
      double s = 0;
      for (int i = 0; i < m; ++i) s += model[i] * (res[i] + model[i] / 2.0);
      model_cost_change = -s;
    }
    it.step_is_valid = step_ok && model_cost_change > 0.0;
    if (!it.step_is_valid) {
      // HandleInvalidStep
      if (++num_consecutive_invalid >= c.max_num_consecutive_invalid_steps) { finish(kFailure, "too many invalid steps"); break; }
      radius *= 0.5;
      reuse_diagonal = true;
      it.cost = x_cost;
      it.step_is_successful = false;
      continue;
    }
    num_consecutive_invalid = 0;
    for (int j = 0; j < n; ++j) delta[j] = step[j] * scale[j];

    // ---- ComputeCandidatePointAndEvaluateCost
    problem.Plus(x.data(), delta.data(), cand.data());
    std::vector<double> cres(m);
    problem.Evaluate(cand.data(), cres.data(), nullptr);
    summary->num_residual_evaluations++;
    double cs = 0;
    for (int i = 0; i < m; ++i) cs += cres[i] * cres[i];
    double candidate_cost = 0.5 * cs;
    if (!std::isfinite(candidate_cost)) candidate_cost = std::numeric_limits<double>::max();

    // ---- ParameterToleranceReached
    {
      double s = 0;
      for (int i = 0; i < na; ++i) s += (x[i] - cand[i]) * (x[i] - cand[i]);
      it.step_norm = std::sqrt(s);
      if (it.step_norm <= c.parameter_tolerance * (x_norm + c.parameter_tolerance)) {
        finish(kConvergence, "parameter tolerance");  // the unfinished iteration is not recorded
        break;
      }
    }
    // ---- FunctionToleranceReached
    it.cost_change = x_cost - candidate_cost;
    if (std::fabs(it.cost_change) <= c.function_tolerance * x_cost) {
      finish(kConvergence, "function tolerance");
      break;
    }
    // ---- IsStepSuccessful
    {
      const double relative = (ev_current - candidate_cost) / model_cost_change;
      const double historical = (ev_reference - candidate_cost) / (ev_acc_ref + model_cost_change);
      it.relative_decrease = std::max(relative, historical);
    }
    if (it.relative_decrease > c.min_relative_decrease) {
      // HandleSuccessfulStep
      x = cand;
      x_norm = norm7(x.data());
      evaluate_gradient_and_jacobian();
      it.cost = x_cost;
      it.step_is_successful = true;
      // strategy StepAccepted
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(c.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      // evaluator StepAccepted
      ev_current = candidate_cost;
      ev_acc_cand += model_cost_change;
      ev_acc_ref += model_cost_change;
      if (ev_current < ev_minimum) {
        ev_minimum = ev_current;
        ev_num_nonmono = 0;
        ev_candidate = ev_current;
        ev_acc_cand = 0;
      } else {
        ++ev_num_nonmono;
        if (ev_current > ev_candidate) {
          ev_candidate = ev_current;
          ev_acc_cand = 0;
        }
      }
      if (ev_num_nonmono == max_nonmono) {
        ev_reference = ev_candidate;
        ev_acc_ref = ev_acc_cand;
      }
    } else {
      // HandleUnsuccessfulStep
      it.step_is_successful = false;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      it.cost = candidate_cost;
    }
  }
  // SetSummaryFinalCost (solver.cc): min over the cost of every recorded iteration.
  summary->final_cost = summary->initial_cost;
  for (const IterationLog& l : summary->iterations) summary->final_cost = std::min(summary->final_cost, l.cost);
}

// CeresScanMatcher3D::Match (ceres_scan_matcher_3d.cc:71-123)
inline void ceres_scan_match(const CeresMatcherOptions& opt, const V3d& target_translation, const Rigid3d& initial,
                             const std::vector<CloudAndGrid>& pairs, Rigid3d* pose, SolveSummary* summary) {
  ScanMatchProblem problem(opt, target_translation, initial, pairs);
  double x[7] = {initial.t.x, initial.t.y, initial.t.z, initial.q.w, initial.q.x, initial.q.y, initial.q.z};
  solve_trust_region(problem, opt.use_nonmonotonic_steps, opt.max_num_iterations, x, summary);
  pose->t = {x[0], x[1], x[2]};
  pose->q = {x[3], x[4], x[5], x[6]};
}

}  // namespace orc
