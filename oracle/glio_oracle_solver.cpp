// glio_oracle_solver.cpp — CPU restatement of the reference's window solve: a ceres::Problem holding
// LidarPlaneNormFactor residual blocks (+ a few generic host factors standing in for IMU / prior / GNSS) solved
// the way ceres::Solve does it for the options of Estimator.cpp:2424-2433 / 3275-3284.
//
// TEST INFRASTRUCTURE ONLY (see glio_oracle.h).  PARITY UNPINNED.
//
// This follows Ceres 2.0.0 literally, on an explicit (compressed-row) Jacobian:
//   ProgramEvaluator/ResidualBlock::Evaluate      ceres.tgz::internal/ceres/residual_block.cc:70-197
//   TrustRegionMinimizer                          .../trust_region_minimizer.cc:66-800
//   DoglegStrategy (traditional + subspace)       .../dogleg_strategy.cc:79-697
//   SparseNormalCholeskySolver                    .../sparse_normal_cholesky_solver.cc:59-113  (dense Cholesky here)
//   TrustRegionStepEvaluator                      .../trust_region_step_evaluator.cc:38-110
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "glio_oracle.h"

namespace {

template <int N> struct J_ { double a; double v[N]; };
// small dynamic-width dual number for the host factors (width up to 32)
struct Dual {
  double a; double v[32];
  Dual() : a(0) { std::memset(v, 0, sizeof(v)); }
  Dual(double s) : a(s) { std::memset(v, 0, sizeof(v)); }  // NOLINT
  Dual(double s, int k) : a(s) { std::memset(v, 0, sizeof(v)); v[k] = 1.0; }
};
inline Dual operator+(const Dual& f, const Dual& g) { Dual h; h.a = f.a + g.a; for (int i = 0; i < 32; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
inline Dual operator-(const Dual& f, const Dual& g) { Dual h; h.a = f.a - g.a; for (int i = 0; i < 32; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
inline Dual operator-(const Dual& f) { Dual h; h.a = -f.a; for (int i = 0; i < 32; ++i) h.v[i] = -f.v[i]; return h; }
inline Dual operator*(const Dual& f, const Dual& g) { Dual h; h.a = f.a * g.a; for (int i = 0; i < 32; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
inline Dual operator/(const Dual& f, const Dual& g) { Dual h; const double gi = 1.0 / g.a, fg = f.a * gi; h.a = fg; for (int i = 0; i < 32; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi; return h; }
inline Dual dsqrt(const Dual& f) { Dual h; const double t = std::sqrt(f.a); h.a = t; const double s = 1.0 / (2.0 * t); for (int i = 0; i < 32; ++i) h.v[i] = f.v[i] * s; return h; }

template <class T> void cross3(const T a[3], const T b[3], T o[3]) { T o0 = a[1] * b[2] - a[2] * b[1], o1 = a[2] * b[0] - a[0] * b[2], o2 = a[0] * b[1] - a[1] * b[0]; o[0] = o0; o[1] = o1; o[2] = o2; }
template <class T> void qrot(const T q[4], const T v[3], T o[3]) {
  const T u[3] = {q[1], q[2], q[3]}; T uv[3]; cross3(u, v, uv); uv[0] = uv[0] + uv[0]; uv[1] = uv[1] + uv[1]; uv[2] = uv[2] + uv[2];
  T c[3]; cross3(u, uv, c); T r0 = v[0] + q[0] * uv[0] + c[0], r1 = v[1] + q[0] * uv[1] + c[1], r2 = v[2] + q[0] * uv[2] + c[2]; o[0] = r0; o[1] = r1; o[2] = r2; }
template <class T> void qmul(const T a[4], const T b[4], T o[4]) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]; o[0] = w; o[1] = x; o[2] = y; o[3] = z; }
template <class T> void qconj(const T q[4], T o[4]) { o[0] = q[0]; o[1] = -q[1]; o[2] = -q[2]; o[3] = -q[3]; }

struct Prior { int kf; double t0[3], q0[4], sb0[9], sw[15]; };
struct Between { int i, j; double dp[3], dq[4], dv[3], dt, sw[15]; };
struct Range { int kf; double lever[3], sat[3], rho, w; };
// MarginalizationFactor (GLIO/src/MarginalizationFactor.cpp:222-330) of the previous window, numbered for THIS window:
// kept blocks = KF0: t, q, speed-bias; KF k = 1..W-2: t, q   (n = 6W + 3 residuals); see glio_b200/csrc/marg.h for the order
struct MargPrior { int W = 0, n = 0; std::vector<double> LJ, lr, x0_pose; double x0_sb[9]; };

struct Opt {
  int32_t max_num_iterations, dogleg_type, use_nonmonotonic_steps, max_consecutive_nonmonotonic_steps;
  double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  int32_t max_num_consecutive_invalid_steps, jacobi_scaling;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  int32_t fuse_candidate_jacobian, reserved;   /* reserved = trust region strategy: 0 DOGLEG, 1 LEVENBERG_MARQUARDT over normal-equation Cholesky,
                                                  2 LEVENBERG_MARQUARDT over DENSE_QR (the front end's options, LidarOdometry.cpp:521-530) */
};
struct Iter { int32_t iteration, step_is_valid, step_is_successful, reserved; double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease, trust_region_radius, mu; };
struct Summary { int32_t termination, num_iterations, num_successful_steps, num_unsuccessful_steps, num_evaluations, num_jacobian_evaluations, num_linear_solves, num_valid_steps; double initial_cost, final_cost; char message[128]; };

struct Problem {
  int W; bool has_sb; int nt, na, n;
  std::vector<double> x;   // ambient [t3 q4 (sb9)] per KF
  double q_lb[4], t_lb[3], huber, lidar_unused;
  // unary lidar
  std::vector<int32_t> kf; std::vector<float> cp, nsd; std::vector<double> score;
  // binary lidar (batch)
  std::vector<int32_t> bkc, bko; std::vector<float> bcp; std::vector<double> bnc, bscore;
  std::vector<Prior> priors; std::vector<Between> betweens; std::vector<Range> ranges;
  MargPrior marg;
  // CRS jacobian
  std::vector<double> r; std::vector<int64_t> rowptr; std::vector<int32_t> col; std::vector<double> val;
  int64_t nrows = 0;
  int mode = 0, nthreads = 1;
};

void quat_plus_jac(const double* x, double* P) { go_quat_plus_jacobian(x, P); }

// number of rows / nnz layout is fixed for a problem: unary rows (6 nnz), prior rows 15 x nt, between 15 x 2nt, range 1 x 6
void layout(Problem& P) {
  const int64_t N = (int64_t)P.kf.size(), NB = (int64_t)P.bkc.size();
  P.nrows = N + NB + 15 * (int64_t)P.priors.size() + 15 * (int64_t)P.betweens.size() + (int64_t)P.ranges.size() + (int64_t)P.marg.n;
  P.rowptr.assign(P.nrows + 1, 0);
  int64_t row = 0, nnz = 0;
  for (int64_t i = 0; i < N; ++i) { P.rowptr[row++] = nnz; nnz += 6; }
  for (int64_t i = 0; i < NB; ++i) { P.rowptr[row++] = nnz; nnz += 12; }
  for (size_t i = 0; i < P.priors.size(); ++i) for (int k = 0; k < 15; ++k) { P.rowptr[row++] = nnz; nnz += P.nt; }
  for (size_t i = 0; i < P.betweens.size(); ++i) for (int k = 0; k < 15; ++k) { P.rowptr[row++] = nnz; nnz += 2 * P.nt; }
  for (size_t i = 0; i < P.ranges.size(); ++i) { P.rowptr[row++] = nnz; nnz += 6; }
  for (int i = 0; i < P.marg.n; ++i) { P.rowptr[row++] = nnz; nnz += P.nt + 6 * (P.marg.W - 2); }
  P.rowptr[row] = nnz;
  P.col.assign(nnz, 0); P.val.assign(nnz, 0.0); P.r.assign(P.nrows, 0.0);
}

// evaluate a host factor with Duals over the ambient blocks of up to two keyframes; fill tangent rows.
// amb index map: kf A -> dual slots [0,16), kf B -> [16,32)
void finish_rows(Problem& P, const double* x, int nres, const Dual* res, int kfA, int kfB, int64_t row0, bool want_jac, double* cost) {
  // Ceres: no loss on these; cost = 0.5 |r|^2
  double s = 0;
  for (int k = 0; k < nres; ++k) { P.r[row0 + k] = res[k].a; s += res[k].a * res[k].a; }
  *cost += 0.5 * s;
  if (!want_jac) return;
  const int kfs[2] = {kfA, kfB};
  for (int k = 0; k < nres; ++k) {
    int64_t p = P.rowptr[row0 + k];
    for (int b = 0; b < 2; ++b) {
      if (kfs[b] < 0) continue;
      const int kf = kfs[b]; const double* xa = x + (size_t)P.na * kf; const double* dv = res[k].v + 16 * b;
      double Pm[12]; quat_plus_jac(xa + 3, Pm);
      for (int c = 0; c < 3; ++c) { P.col[p] = P.nt * kf + c; P.val[p++] = dv[c]; }
      for (int c = 0; c < 3; ++c) { double a = 0; for (int m = 0; m < 4; ++m) a += dv[3 + m] * Pm[3 * m + c]; P.col[p] = P.nt * kf + 3 + c; P.val[p++] = a; }
      if (P.has_sb) for (int c = 0; c < 9; ++c) { P.col[p] = P.nt * kf + 6 + c; P.val[p++] = dv[7 + c]; }
    }
  }
}

// MarginalizationFactor::Evaluate (MarginalizationFactor.cpp:232-330), literal: dx per kept block (quaternions: 2 vec of the
// NORMALISED q0^-1 q, sign flipped when w < 0), residual = linearized_residuals + linearized_jacobians dx, and the analytic
// ambient Jacobians the factor returns (quaternion block: +-2 J_cols Qleft(q0^-1).bottomRightCorner<3,4>(), which ignores the
// normalisation).  Jamb[i] is n x 16 per keyframe slot: columns t(3) q(4: w,x,y,z) sb(9).
struct MargEval { std::vector<double> res; std::vector<std::vector<double>> Jamb; };
void marg_prior_eval(const Problem& P, const double* x, MargEval& E) {
  const MargPrior& M = P.marg; const int n = M.n, W = M.W;
  std::vector<double> dx(n, 0.0);
  E.Jamb.assign(W - 1, std::vector<double>((size_t)n * 16, 0.0));
  for (int k = 0; k <= W - 2; ++k) {
    const double* xa = x + (size_t)P.na * k; const double* x0 = &M.x0_pose[(size_t)7 * k];
    const int it = k == 0 ? 0 : 15 + 6 * (k - 1), iq = it + 3;
    for (int c = 0; c < 3; ++c) dx[it + c] = xa[c] - x0[c];
    // Quaterniond(x0).inverse() * Quaterniond(x)
    const double n2 = x0[3] * x0[3] + x0[4] * x0[4] + x0[5] * x0[5] + x0[6] * x0[6];
    const double qi[4] = {x0[3] / n2, -x0[4] / n2, -x0[5] / n2, -x0[6] / n2};
    const double* q = xa + 3;
    const double e[4] = {qi[0] * q[0] - qi[1] * q[1] - qi[2] * q[2] - qi[3] * q[3], qi[0] * q[1] + qi[1] * q[0] + qi[2] * q[3] - qi[3] * q[2],
                         qi[0] * q[2] + qi[2] * q[0] + qi[3] * q[1] - qi[1] * q[3], qi[0] * q[3] + qi[3] * q[0] + qi[1] * q[2] - qi[2] * q[1]};
    const double nrm = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3]);
    const double sgn = e[0] < 0 ? -1.0 : 1.0;
    for (int c = 0; c < 3; ++c) dx[iq + c] = sgn * 2.0 * e[1 + c] / nrm;
    if (k == 0) for (int c = 0; c < 9; ++c) dx[6 + c] = (P.has_sb ? xa[7 + c] : M.x0_sb[c]) - M.x0_sb[c];
    // Qleft(q0^-1).bottomRightCorner<3,4>() = [vec | w I + skew(vec)]
    const double QL[3][4] = {{qi[1], qi[0], -qi[3], qi[2]}, {qi[2], qi[3], qi[0], -qi[1]}, {qi[3], -qi[2], qi[1], qi[0]}};
    std::vector<double>& J = E.Jamb[k];
    for (int r = 0; r < n; ++r) {
      const double* lj = &M.LJ[(size_t)r * n];
      for (int c = 0; c < 3; ++c) J[(size_t)r * 16 + c] = lj[it + c];
      for (int c = 0; c < 4; ++c) { double a = 0; for (int m = 0; m < 3; ++m) a += lj[iq + m] * QL[m][c]; J[(size_t)r * 16 + 3 + c] = sgn * 2.0 * a; }
      if (k == 0) for (int c = 0; c < 9; ++c) J[(size_t)r * 16 + 7 + c] = lj[6 + c];
    }
  }
  E.res.assign(n, 0.0);
  for (int r = 0; r < n; ++r) { double a = M.lr[r]; const double* lj = &M.LJ[(size_t)r * n]; for (int c = 0; c < n; ++c) a += lj[c] * dx[c]; E.res[r] = a; }
}

bool evaluate(Problem& P, const double* x, bool want_jac, double* cost_out) {
  const int64_t N = (int64_t)P.kf.size();
  double cost = 0;
  // ---- LiDAR rows (through the per-residual oracle of glio_oracle.cpp) ----
  {
    std::vector<double> poses((size_t)P.W * 7);
    for (int k = 0; k < P.W; ++k) for (int i = 0; i < 7; ++i) poses[7 * k + i] = x[(size_t)P.na * k + i];
    const int T = std::max(1, P.nthreads);
    std::vector<double> ct(T, 0.0);
#pragma omp parallel for num_threads(T) schedule(static)
    for (int t = 0; t < T; ++t) {
      const int64_t lo = N * t / T, hi = N * (t + 1) / T;
      if (hi > lo) {
        std::vector<double> Jt(want_jac ? (size_t)(hi - lo) * 6 : 0);
        go_eval_unary(P.mode, 0, P.W, poses.data(), P.q_lb, P.t_lb, P.huber, hi - lo, P.kf.data() + lo, P.cp.data() + 3 * lo,
                      P.nsd.data() + 4 * lo, P.score.data() + lo, P.r.data() + lo, want_jac ? Jt.data() : nullptr, nullptr, nullptr, nullptr, &ct[t]);
        if (want_jac) for (int64_t i = lo; i < hi; ++i) {
          int64_t p = P.rowptr[i];
          for (int c = 0; c < 6; ++c) { P.col[p + c] = P.nt * P.kf[i] + c; P.val[p + c] = Jt[(size_t)(i - lo) * 6 + c]; }
        }
      }
    }
    for (int t = 0; t < T; ++t) cost += ct[t];
  }
  int64_t row = N;
  // ---- binary LiDAR rows (BinaryLidarPlaneNormFactor, no loss: Estimator.cpp:2768) ----
  {
    const int64_t NB = (int64_t)P.bkc.size();
    if (NB > 0) {
      std::vector<double> poses((size_t)P.W * 7);
      for (int k = 0; k < P.W; ++k) for (int i = 0; i < 7; ++i) poses[7 * k + i] = x[(size_t)P.na * k + i];
      std::vector<double> Jt(want_jac ? (size_t)NB * 12 : 0);
      double ct = 0;
      go_eval_binary(P.mode, P.W, poses.data(), 0.0, NB, P.bkc.data(), P.bko.data(), P.bcp.data(), P.bnc.data(), P.bscore.data(),
                     P.r.data() + row, want_jac ? Jt.data() : nullptr, nullptr, nullptr, nullptr, &ct);
      cost += ct;
      if (want_jac) for (int64_t i = 0; i < NB; ++i) {
        int64_t p = P.rowptr[row + i];
        for (int c = 0; c < 6; ++c) { P.col[p + c] = P.nt * P.bkc[i] + c; P.val[p + c] = Jt[(size_t)i * 12 + c]; }
        for (int c = 0; c < 6; ++c) { P.col[p + 6 + c] = P.nt * P.bko[i] + c; P.val[p + 6 + c] = Jt[(size_t)i * 12 + 6 + c]; }
      }
      row += NB;
    }
  }
  // ---- prior rows ----
  for (const Prior& f : P.priors) {
    const double* xa = x + (size_t)P.na * f.kf;
    Dual t[3] = {Dual(xa[0], 0), Dual(xa[1], 1), Dual(xa[2], 2)};
    Dual q[4] = {Dual(xa[3], 3), Dual(xa[4], 4), Dual(xa[5], 5), Dual(xa[6], 6)};
    Dual res[15];
    for (int c = 0; c < 3; ++c) res[c] = Dual(f.sw[c]) * (t[c] - Dual(f.t0[c]));
    Dual q0[4] = {Dual(f.q0[0]), Dual(f.q0[1]), Dual(f.q0[2]), Dual(f.q0[3])}, q0c[4], dq[4];
    qconj(q0, q0c); qmul(q0c, q, dq);
    for (int c = 0; c < 3; ++c) res[3 + c] = Dual(f.sw[3 + c]) * (Dual(2.0) * dq[1 + c]);
    for (int c = 0; c < 9; ++c) res[6 + c] = P.has_sb ? Dual(f.sw[6 + c]) * (Dual(xa[7 + c], 7 + c) - Dual(f.sb0[c])) : Dual(0.0);
    finish_rows(P, x, 15, res, f.kf, -1, row, want_jac, &cost);
    row += 15;
  }
  // ---- between rows (IMU-like chain) ----
  for (const Between& f : P.betweens) {
    const double* xi = x + (size_t)P.na * f.i; const double* xj = x + (size_t)P.na * f.j;
    Dual ti[3] = {Dual(xi[0], 0), Dual(xi[1], 1), Dual(xi[2], 2)}, qi[4] = {Dual(xi[3], 3), Dual(xi[4], 4), Dual(xi[5], 5), Dual(xi[6], 6)};
    Dual tj[3] = {Dual(xj[0], 16), Dual(xj[1], 17), Dual(xj[2], 18)}, qj[4] = {Dual(xj[3], 19), Dual(xj[4], 20), Dual(xj[5], 21), Dual(xj[6], 22)};
    Dual vi[3], vj[3], bi[6], bj[6];
    for (int c = 0; c < 3; ++c) { vi[c] = P.has_sb ? Dual(xi[7 + c], 7 + c) : Dual(0.0); vj[c] = P.has_sb ? Dual(xj[7 + c], 23 + c) : Dual(0.0); }
    for (int c = 0; c < 6; ++c) { bi[c] = P.has_sb ? Dual(xi[10 + c], 10 + c) : Dual(0.0); bj[c] = P.has_sb ? Dual(xj[10 + c], 26 + c) : Dual(0.0); }
    Dual qic[4]; qconj(qi, qic);
    Dual d[3] = {tj[0] - ti[0] - vi[0] * Dual(f.dt), tj[1] - ti[1] - vi[1] * Dual(f.dt), tj[2] - ti[2] - vi[2] * Dual(f.dt)};
    Dual rp[3]; qrot(qic, d, rp);
    Dual res[15];
    for (int c = 0; c < 3; ++c) res[c] = Dual(f.sw[c]) * (rp[c] - Dual(f.dp[c]));
    Dual dqc[4] = {Dual(f.dq[0]), Dual(-f.dq[1]), Dual(-f.dq[2]), Dual(-f.dq[3])}, qij[4], e[4];
    qmul(qic, qj, qij); qmul(dqc, qij, e);
    for (int c = 0; c < 3; ++c) res[3 + c] = Dual(f.sw[3 + c]) * (Dual(2.0) * e[1 + c]);
    Dual dvv[3] = {vj[0] - vi[0], vj[1] - vi[1], vj[2] - vi[2]}, rv[3]; qrot(qic, dvv, rv);
    for (int c = 0; c < 3; ++c) res[6 + c] = Dual(f.sw[6 + c]) * (rv[c] - Dual(f.dv[c]));
    for (int c = 0; c < 6; ++c) res[9 + c] = Dual(f.sw[9 + c]) * (bj[c] - bi[c]);
    finish_rows(P, x, 15, res, f.i, f.j, row, want_jac, &cost);
    row += 15;
  }
  // ---- range rows (pseudorange-like) ----
  for (const Range& f : P.ranges) {
    const double* xa = x + (size_t)P.na * f.kf;
    Dual t[3] = {Dual(xa[0], 0), Dual(xa[1], 1), Dual(xa[2], 2)}, q[4] = {Dual(xa[3], 3), Dual(xa[4], 4), Dual(xa[5], 5), Dual(xa[6], 6)};
    Dual lv[3] = {Dual(f.lever[0]), Dual(f.lever[1]), Dual(f.lever[2])}, pw[3]; qrot(q, lv, pw);
    Dual d[3] = {pw[0] + t[0] - Dual(f.sat[0]), pw[1] + t[1] - Dual(f.sat[1]), pw[2] + t[2] - Dual(f.sat[2])};
    Dual res[1]; res[0] = Dual(f.w) * (dsqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) - Dual(f.rho));
    // range rows only carry the pose part (6 nnz)
    P.r[row] = res[0].a; cost += 0.5 * res[0].a * res[0].a;
    if (want_jac) {
      int64_t p = P.rowptr[row]; double Pm[12]; quat_plus_jac(xa + 3, Pm);
      for (int c = 0; c < 3; ++c) { P.col[p] = P.nt * f.kf + c; P.val[p++] = res[0].v[c]; }
      for (int c = 0; c < 3; ++c) { double a = 0; for (int m = 0; m < 4; ++m) a += res[0].v[3 + m] * Pm[3 * m + c]; P.col[p] = P.nt * f.kf + 3 + c; P.val[p++] = a; }
    }
    row += 1;
  }
  // ---- marginalisation prior rows (no loss function: Estimator.cpp:2153-2158) ----
  if (P.marg.n > 0) {
    MargEval E; marg_prior_eval(P, x, E);
    const int n = P.marg.n, W = P.marg.W;
    double s2 = 0;
    for (int r = 0; r < n; ++r) { P.r[row + r] = E.res[r]; s2 += E.res[r] * E.res[r]; }
    cost += 0.5 * s2;
    if (want_jac) for (int r = 0; r < n; ++r) {
      int64_t p = P.rowptr[row + r];
      for (int k = 0; k <= W - 2; ++k) {
        const double* xa = x + (size_t)P.na * k; const double* dv = &E.Jamb[k][(size_t)r * 16];
        double Pm[12]; quat_plus_jac(xa + 3, Pm);
        for (int c = 0; c < 3; ++c) { P.col[p] = P.nt * k + c; P.val[p++] = dv[c]; }
        for (int c = 0; c < 3; ++c) { double a = 0; for (int m = 0; m < 4; ++m) a += dv[3 + m] * Pm[3 * m + c]; P.col[p] = P.nt * k + 3 + c; P.val[p++] = a; }
        if (k == 0 && P.has_sb) for (int c = 0; c < 9; ++c) { P.col[p] = P.nt * k + 6 + c; P.val[p++] = dv[7 + c]; }
      }
    }
    row += n;
  }
  *cost_out = cost;
  return std::isfinite(cost);
}

void plus(const Problem& P, const double* x, const double* d, double* o) {
  for (int k = 0; k < P.W; ++k) {
    const double* xa = x + (size_t)P.na * k; double* oa = o + (size_t)P.na * k; const double* dk = d + (size_t)P.nt * k;
    for (int c = 0; c < 3; ++c) oa[c] = xa[c] + dk[c];
    go_quat_plus(xa + 3, dk + 3, oa + 3);
    if (P.has_sb) for (int c = 0; c < 9; ++c) oa[7 + c] = xa[7 + c] + dk[6 + c];
  }
}

// CRS helpers (SparseMatrix::LeftMultiply / RightMultiply / SquaredColumnNorm / ScaleColumns)
void left_multiply(const Problem& P, const double* rvec, double* y) { for (int64_t i = 0; i < P.nrows; ++i) for (int64_t p = P.rowptr[i]; p < P.rowptr[i + 1]; ++p) y[P.col[p]] += P.val[p] * rvec[i]; }
void right_multiply(const Problem& P, const double* xv, double* y) { for (int64_t i = 0; i < P.nrows; ++i) { double s = 0; for (int64_t p = P.rowptr[i]; p < P.rowptr[i + 1]; ++p) s += P.val[p] * xv[P.col[p]]; y[i] += s; } }
void sq_col_norm(const Problem& P, double* out) { for (int i = 0; i < P.n; ++i) out[i] = 0; for (size_t p = 0; p < P.val.size(); ++p) out[P.col[p]] += P.val[p] * P.val[p]; }
void scale_columns(Problem& P, const double* s) { for (size_t p = 0; p < P.val.size(); ++p) P.val[p] *= s[P.col[p]]; }

bool dense_cholesky_solve(std::vector<double>& A, int n, const double* b, double* x) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    const double l = std::sqrt(d); A[(size_t)j * n + j] = l;
    for (int i = j + 1; i < n; ++i) { double s = A[(size_t)i * n + j]; for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k]; A[(size_t)i * n + j] = s / l; }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * x[k]; x[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * x[k]; x[i] = s / A[(size_t)i * n + i]; }
  for (int i = 0; i < n; ++i) if (!std::isfinite(x[i])) return false;
  return true;
}

// real parts of all roots of a quartic (highest power first) — Aberth/Durand-Kerner; what
// FindPolynomialRoots(poly, &real, NULL) returns in ceres.tgz::internal/ceres/polynomial.cc
// min |A x - b| for a dense row-major m x n matrix (m >= n) by unpivoted Householder QR: what Ceres' DENSE_QR does with the
// LM-augmented Jacobian [J; D] and [r; 0] (ceres.tgz::internal/ceres/dense_qr_solver.cc:120-153, Eigen householderQr().solve()).
bool dense_qr_solve(std::vector<double>& A, int64_t m, int n, std::vector<double>& b, double* x) {
  for (int k = 0; k < n; ++k) {
    double tail = 0; for (int64_t i = k + 1; i < m; ++i) tail += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    const double c0 = A[(size_t)k * n + k];
    double beta, tau;
    if (tail <= std::numeric_limits<double>::min()) { tau = 0; beta = c0; }
    else {
      beta = std::sqrt(c0 * c0 + tail); if (c0 >= 0) beta = -beta;
      const double den = c0 - beta;
      for (int64_t i = k + 1; i < m; ++i) A[(size_t)i * n + k] /= den;
      tau = (beta - c0) / beta;
    }
    A[(size_t)k * n + k] = beta;
    if (tau != 0) {
      for (int j = k + 1; j < n; ++j) {
        double t = A[(size_t)k * n + j]; for (int64_t i = k + 1; i < m; ++i) t += A[(size_t)i * n + k] * A[(size_t)i * n + j];
        A[(size_t)k * n + j] -= tau * t; for (int64_t i = k + 1; i < m; ++i) A[(size_t)i * n + j] -= tau * A[(size_t)i * n + k] * t;
      }
      double t = b[k]; for (int64_t i = k + 1; i < m; ++i) t += A[(size_t)i * n + k] * b[i];
      b[k] -= tau * t; for (int64_t i = k + 1; i < m; ++i) b[i] -= tau * A[(size_t)i * n + k] * t;
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double sacc = b[i]; for (int j = i + 1; j < n; ++j) sacc -= A[(size_t)i * n + j] * x[j];
    x[i] = sacc / A[(size_t)i * n + i];
  }
  for (int i = 0; i < n; ++i) if (!std::isfinite(x[i])) return false;
  return true;
}

bool poly_real_parts(const double* p5, std::vector<double>* out) {
  out->clear(); int lead = 0; while (lead < 5 && p5[lead] == 0.0) ++lead; int deg = 4 - lead; if (deg < 1) return false;
  typedef std::complex<double> cd; std::vector<cd> c(deg + 1); for (int i = 0; i <= deg; ++i) c[i] = p5[lead + i] / p5[lead];
  double bound = 0; for (int i = 1; i <= deg; ++i) bound = std::max(bound, std::abs(c[i])); bound += 1.0;
  std::vector<cd> z(deg); for (int i = 0; i < deg; ++i) z[i] = std::polar(bound * 0.6, 0.7 + 2.0 * M_PI * i / deg);
  auto ev = [&](cd x) { cd v = c[0]; for (int i = 1; i <= deg; ++i) v = v * x + c[i]; return v; };
  for (int it = 0; it < 1000; ++it) { double ch = 0; for (int i = 0; i < deg; ++i) { cd den = 1.0; for (int j = 0; j < deg; ++j) if (j != i) den *= (z[i] - z[j]); if (std::abs(den) == 0) den = 1e-300; cd dz = ev(z[i]) / den; z[i] -= dz; ch = std::max(ch, std::abs(dz) / std::max(1.0, std::abs(z[i]))); } if (ch < 1e-16) break; }
  for (int i = 0; i < deg; ++i) { if (!std::isfinite(z[i].real())) return false; out->push_back(z[i].real()); }
  return true;
}

}  // namespace

extern "C" {

void* go_problem_create(int W, const double* poses, const double* speed_bias, const double q_lb[4], const double t_lb[3], double huber_delta) {
  Problem* P = new Problem();
  P->W = W; P->has_sb = speed_bias != nullptr; P->nt = P->has_sb ? 15 : 6; P->na = P->has_sb ? 16 : 7; P->n = W * P->nt;
  P->x.assign((size_t)W * P->na, 0.0);
  for (int k = 0; k < W; ++k) { for (int i = 0; i < 7; ++i) P->x[(size_t)P->na * k + i] = poses[7 * k + i]; if (P->has_sb) for (int i = 0; i < 9; ++i) P->x[(size_t)P->na * k + 7 + i] = speed_bias[9 * k + i]; }
  for (int i = 0; i < 4; ++i) P->q_lb[i] = q_lb[i];
  for (int i = 0; i < 3; ++i) P->t_lb[i] = t_lb[i];
  P->huber = huber_delta;
  return P;
}
void go_problem_free(void* h) { delete (Problem*)h; }
void go_problem_add_unary(void* h, int64_t N, const int32_t* kf, const float* cp, const float* nsd, const double* score) {
  Problem* P = (Problem*)h;
  P->kf.insert(P->kf.end(), kf, kf + N); P->cp.insert(P->cp.end(), cp, cp + 3 * N); P->nsd.insert(P->nsd.end(), nsd, nsd + 4 * N); P->score.insert(P->score.end(), score, score + N);
}
void go_problem_add_binary(void* h, int64_t N, const int32_t* kf_c, const int32_t* kf_o, const float* cp, const double* nc, const double* score) {
  Problem* P = (Problem*)h;
  P->bkc.insert(P->bkc.end(), kf_c, kf_c + N); P->bko.insert(P->bko.end(), kf_o, kf_o + N); P->bcp.insert(P->bcp.end(), cp, cp + 3 * N);
  P->bnc.insert(P->bnc.end(), nc, nc + 6 * N); P->bscore.insert(P->bscore.end(), score, score + N);
}
void go_problem_add_prior(void* h, int kf, const double t0[3], const double q0[4], const double* sb0, const double sw[15]) {
  Prior f; f.kf = kf;
  for (int i = 0; i < 3; ++i) f.t0[i] = t0[i];
  for (int i = 0; i < 4; ++i) f.q0[i] = q0[i];
  for (int i = 0; i < 9; ++i) f.sb0[i] = sb0 ? sb0[i] : 0.0;
  for (int i = 0; i < 15; ++i) f.sw[i] = sw[i];
  ((Problem*)h)->priors.push_back(f);
}
void go_problem_add_between(void* h, int i, int j, const double dp[3], const double dq[4], const double dv[3], double dt, const double sw[15]) {
  Between f; f.i = i; f.j = j; for (int k = 0; k < 3; ++k) { f.dp[k] = dp[k]; f.dv[k] = dv[k]; } for (int k = 0; k < 4; ++k) f.dq[k] = dq[k]; f.dt = dt; for (int k = 0; k < 15; ++k) f.sw[k] = sw[k];
  ((Problem*)h)->betweens.push_back(f);
}
void go_problem_add_range(void* h, int kf, const double lever[3], const double sat[3], double rho, double w) {
  Range f; f.kf = kf; for (int k = 0; k < 3; ++k) { f.lever[k] = lever[k]; f.sat[k] = sat[k]; } f.rho = rho; f.w = w;
  ((Problem*)h)->ranges.push_back(f);
}
void go_problem_set_state(void* h, const double* poses, const double* speed_bias) {
  Problem* P = (Problem*)h;
  for (int k = 0; k < P->W; ++k) { for (int i = 0; i < 7; ++i) P->x[(size_t)P->na * k + i] = poses[7 * k + i]; if (P->has_sb && speed_bias) for (int i = 0; i < 9; ++i) P->x[(size_t)P->na * k + 7 + i] = speed_bias[9 * k + i]; }
}
void go_problem_get_state(void* h, double* poses, double* speed_bias) {
  Problem* P = (Problem*)h;
  for (int k = 0; k < P->W; ++k) { for (int i = 0; i < 7; ++i) poses[7 * k + i] = P->x[(size_t)P->na * k + i]; if (P->has_sb && speed_bias) for (int i = 0; i < 9; ++i) speed_bias[9 * k + i] = P->x[(size_t)P->na * k + 7 + i]; }
}
// dense tangent-space normal equations of the host factors alone at the current state (for cross-checking the
// product's analytic host factors): H[n*n], g[n], cost
void go_problem_host_normal_eq(void* h, double* H, double* g, double* cost) {
  Problem* P = (Problem*)h;
  Problem Q = *P; Q.kf.clear(); Q.cp.clear(); Q.nsd.clear(); Q.score.clear(); Q.bkc.clear(); Q.bko.clear(); Q.bcp.clear(); Q.bnc.clear(); Q.bscore.clear();
  layout(Q); double c = 0; evaluate(Q, Q.x.data(), true, &c);
  const int n = Q.n; std::fill(H, H + (size_t)n * n, 0.0); std::fill(g, g + n, 0.0);
  for (int64_t i = 0; i < Q.nrows; ++i) for (int64_t p = Q.rowptr[i]; p < Q.rowptr[i + 1]; ++p) {
    g[Q.col[p]] += Q.val[p] * Q.r[i];
    for (int64_t q = Q.rowptr[i]; q < Q.rowptr[i + 1]; ++q) H[(size_t)Q.col[p] * n + Q.col[q]] += Q.val[p] * Q.val[q];
  }
  *cost = c;
}

void go_problem_set_marg_prior(void* h, int W, const double* lin_jac, const double* lin_res, const double* x0_pose, const double* x0_sb) {
  Problem& P = *(Problem*)h; MargPrior& M = P.marg;
  if (W <= 0 || !lin_jac) { M = MargPrior(); return; }
  M.W = W; M.n = 6 * W + 3;
  M.LJ.assign(lin_jac, lin_jac + (size_t)M.n * M.n); M.lr.assign(lin_res, lin_res + M.n);
  M.x0_pose.assign(x0_pose, x0_pose + (size_t)(W - 1) * 7);
  for (int c = 0; c < 9; ++c) M.x0_sb[c] = x0_sb ? x0_sb[c] : 0.0;
}

namespace {
// cyclic Jacobi (the oracle's own eigen-solver: deliberately not the product's tridiagonal QL)
void jacobi_eigh(std::vector<double> a, int n, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0); for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 80; ++sweep) {
    double off = 0, dg = 0;
    for (int i = 0; i < n; ++i) { dg += a[(size_t)i * n + i] * a[(size_t)i * n + i]; for (int j = i + 1; j < n; ++j) off += a[(size_t)i * n + j] * a[(size_t)i * n + j]; }
    if (off <= 1e-32 * (dg + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
      const double apq = a[(size_t)p * n + q]; if (apq == 0.0) continue;
      const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0)), c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
      for (int k = 0; k < n; ++k) { const double x1 = a[(size_t)k * n + p], x2 = a[(size_t)k * n + q]; a[(size_t)k * n + p] = c * x1 - sn * x2; a[(size_t)k * n + q] = sn * x1 + c * x2; }
      for (int k = 0; k < n; ++k) { const double x1 = a[(size_t)p * n + k], x2 = a[(size_t)q * n + k]; a[(size_t)p * n + k] = c * x1 - sn * x2; a[(size_t)q * n + k] = sn * x1 + c * x2; }
      for (int k = 0; k < n; ++k) { const double x1 = V[(size_t)k * n + p], x2 = V[(size_t)k * n + q]; V[(size_t)k * n + p] = c * x1 - sn * x2; V[(size_t)k * n + q] = sn * x1 + c * x2; }
    }
  }
  w.resize(n); for (int i = 0; i < n; ++i) w[i] = a[(size_t)i * n + i];
}
inline int marg_off(int kf) { return kf == 0 ? 0 : (kf == 1 ? 15 : 30 + 6 * (kf - 2)); }
// ThreadsConstructA (MarginalizationFactor.cpp:3-29) for one residual block given ambient Jacobians in Dual slots of two keyframes
void construct_A(int N, std::vector<double>& A, std::vector<double>& b, int nres, const Dual* res, int kfA, int kfB) {
  const int kfs[2] = {kfA, kfB};
  for (int r = 0; r < nres; ++r) {
    int col[40]; double val[40]; int nz = 0;
    for (int s = 0; s < 2; ++s) {
      if (kfs[s] < 0) continue;
      const double* dv = res[r].v + 16 * s; const int base = marg_off(kfs[s]);
      for (int c = 0; c < 3; ++c) { col[nz] = base + c; val[nz++] = dv[c]; }
      for (int c = 0; c < 3; ++c) { col[nz] = base + 3 + c; val[nz++] = dv[4 + c]; }      // rightCols(3) of the 4-wide quaternion Jacobian
      if (kfs[s] <= 1) for (int c = 0; c < 9; ++c) { col[nz] = base + 6 + c; val[nz++] = dv[7 + c]; }
    }
    for (int i = 0; i < nz; ++i) { b[col[i]] += val[i] * res[r].a; for (int j = 0; j < nz; ++j) A[(size_t)col[i] * N + col[j]] += val[i] * val[j]; }
  }
}
}  // namespace

// MarginalizationInfo::PreMarginalize + Marginalize (MarginalizationFactor.cpp:107-202) over the factors the Estimator hands it
// (Estimator.cpp:2464-2576): the previous prior, [stand-in priors on KF0], the IMU-like factor KF0 -> KF1, every LiDAR factor of
// the window.  Requires speed/bias states.  Outputs in the prior ordering (next window's numbering).
int go_problem_marginalize(void* h, double eps, int mode, double* lin_jac, double* lin_res, double* x0_pose, double* x0_sb, double* A_out, double* b_out) {
  Problem& P = *(Problem*)h;
  if (!P.has_sb || P.W < 2) return -1;
  const int W = P.W, N = 6 * W + 18, m = 15, n = N - m;
  const double* x = P.x.data();
  std::vector<double> A((size_t)N * N, 0.0), b(N, 0.0);
  // LiDAR factors: ambient x,y,z columns, Huber-corrected per ResidualBlockInfo::Evaluate (same corrector as Ceres')
  {
    std::vector<double> poses((size_t)W * 7), Hd((size_t)36 * W * W, 0.0), gd((size_t)6 * W, 0.0); double ct = 0;
    for (int k = 0; k < W; ++k) for (int i = 0; i < 7; ++i) poses[7 * k + i] = x[(size_t)P.na * k + i];
    const int64_t NL = (int64_t)P.kf.size();
    if (NL > 0) go_eval_unary(mode, 1, W, poses.data(), P.q_lb, P.t_lb, P.huber, NL, P.kf.data(), P.cp.data(), P.nsd.data(), P.score.data(), nullptr, nullptr, nullptr, Hd.data(), gd.data(), &ct);
    for (int k = 0; k < W; ++k) for (int p = 0; p < 6; ++p) { b[marg_off(k) + p] += gd[6 * k + p]; for (int q = 0; q < 6; ++q) A[(size_t)(marg_off(k) + p) * N + marg_off(k) + q] += Hd[(size_t)(6 * k + p) * 6 * W + 6 * k + q]; }
  }
  for (const Prior& f : P.priors) {
    if (f.kf != 0) continue;
    const double* xa = x;
    Dual t[3] = {Dual(xa[0], 0), Dual(xa[1], 1), Dual(xa[2], 2)}, q[4] = {Dual(xa[3], 3), Dual(xa[4], 4), Dual(xa[5], 5), Dual(xa[6], 6)}, res[15];
    for (int c = 0; c < 3; ++c) res[c] = Dual(f.sw[c]) * (t[c] - Dual(f.t0[c]));
    Dual q0[4] = {Dual(f.q0[0]), Dual(f.q0[1]), Dual(f.q0[2]), Dual(f.q0[3])}, q0c[4], dq[4]; qconj(q0, q0c); qmul(q0c, q, dq);
    for (int c = 0; c < 3; ++c) res[3 + c] = Dual(f.sw[3 + c]) * (Dual(2.0) * dq[1 + c]);
    for (int c = 0; c < 9; ++c) res[6 + c] = Dual(f.sw[6 + c]) * (Dual(xa[7 + c], 7 + c) - Dual(f.sb0[c]));
    construct_A(N, A, b, 15, res, 0, -1);
  }
  for (const Between& f : P.betweens) {
    if (!(f.i == 0 && f.j == 1)) continue;
    const double* xi = x; const double* xj = x + P.na;
    Dual ti[3] = {Dual(xi[0], 0), Dual(xi[1], 1), Dual(xi[2], 2)}, qi[4] = {Dual(xi[3], 3), Dual(xi[4], 4), Dual(xi[5], 5), Dual(xi[6], 6)};
    Dual tj[3] = {Dual(xj[0], 16), Dual(xj[1], 17), Dual(xj[2], 18)}, qj[4] = {Dual(xj[3], 19), Dual(xj[4], 20), Dual(xj[5], 21), Dual(xj[6], 22)};
    Dual vi[3], vj[3], bi[6], bj[6];
    for (int c = 0; c < 3; ++c) { vi[c] = Dual(xi[7 + c], 7 + c); vj[c] = Dual(xj[7 + c], 23 + c); }
    for (int c = 0; c < 6; ++c) { bi[c] = Dual(xi[10 + c], 10 + c); bj[c] = Dual(xj[10 + c], 26 + c); }
    Dual qic[4]; qconj(qi, qic);
    Dual d[3] = {tj[0] - ti[0] - vi[0] * Dual(f.dt), tj[1] - ti[1] - vi[1] * Dual(f.dt), tj[2] - ti[2] - vi[2] * Dual(f.dt)}, rp[3]; qrot(qic, d, rp);
    Dual res[15];
    for (int c = 0; c < 3; ++c) res[c] = Dual(f.sw[c]) * (rp[c] - Dual(f.dp[c]));
    Dual dqc[4] = {Dual(f.dq[0]), Dual(-f.dq[1]), Dual(-f.dq[2]), Dual(-f.dq[3])}, qij[4], e[4]; qmul(qic, qj, qij); qmul(dqc, qij, e);
    for (int c = 0; c < 3; ++c) res[3 + c] = Dual(f.sw[3 + c]) * (Dual(2.0) * e[1 + c]);
    Dual dvv[3] = {vj[0] - vi[0], vj[1] - vi[1], vj[2] - vi[2]}, rv[3]; qrot(qic, dvv, rv);
    for (int c = 0; c < 3; ++c) res[6 + c] = Dual(f.sw[6 + c]) * (rv[c] - Dual(f.dv[c]));
    for (int c = 0; c < 6; ++c) res[9 + c] = Dual(f.sw[9 + c]) * (bj[c] - bi[c]);
    construct_A(N, A, b, 15, res, 0, 1);
  }
  if (P.marg.n > 0) {
    if (P.marg.W != W) return -2;
    MargEval E; marg_prior_eval(P, x, E);
    const int np = P.marg.n;
    for (int r = 0; r < np; ++r) {
      int col[160]; double val[160]; int nz = 0;
      for (int k = 0; k <= W - 2; ++k) {
        const double* dv = &E.Jamb[k][(size_t)r * 16]; const int base = marg_off(k);
        for (int c = 0; c < 3; ++c) { col[nz] = base + c; val[nz++] = dv[c]; }
        for (int c = 0; c < 3; ++c) { col[nz] = base + 3 + c; val[nz++] = dv[4 + c]; }
        if (k == 0) for (int c = 0; c < 9; ++c) { col[nz] = base + 6 + c; val[nz++] = dv[7 + c]; }
      }
      for (int i = 0; i < nz; ++i) { if (val[i] == 0.0) continue; b[col[i]] += val[i] * E.res[r]; for (int j = 0; j < nz; ++j) A[(size_t)col[i] * N + col[j]] += val[i] * val[j]; }
    }
  }
  if (A_out) std::memcpy(A_out, A.data(), sizeof(double) * A.size());
  if (b_out) std::memcpy(b_out, b.data(), sizeof(double) * b.size());
  // Marginalize (:176-201)
  std::vector<double> Amm((size_t)m * m), w, V;
  for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * N + j] + A[(size_t)j * N + i]);
  jacobi_eigh(Amm, m, w, V);
  std::vector<double> Ainv((size_t)m * m, 0.0);
  for (int k = 0; k < m; ++k) { if (!(w[k] > eps)) continue; for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) Ainv[(size_t)i * m + j] += V[(size_t)i * m + k] * V[(size_t)j * m + k] / w[k]; }
  std::vector<double> T((size_t)n * m, 0.0), Ar((size_t)n * n), br(n);
  for (int i = 0; i < n; ++i) for (int k = 0; k < m; ++k) for (int j = 0; j < m; ++j) T[(size_t)i * m + j] += A[(size_t)(m + i) * N + k] * Ainv[(size_t)k * m + j];
  for (int i = 0; i < n; ++i) {
    double sacc = b[m + i]; for (int k = 0; k < m; ++k) sacc -= T[(size_t)i * m + k] * b[k]; br[i] = sacc;
    for (int j = 0; j < n; ++j) { double v = A[(size_t)(m + i) * N + m + j]; for (int k = 0; k < m; ++k) v -= T[(size_t)i * m + k] * A[(size_t)k * N + m + j]; Ar[(size_t)i * n + j] = v; }
  }
  for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) Ar[(size_t)j * n + i] = Ar[(size_t)i * n + j];     // SelfAdjointEigenSolver reads the lower triangle
  std::vector<double> w2, V2; jacobi_eigh(Ar, n, w2, V2);
  for (int k = 0; k < n; ++k) {
    const double S = w2[k] > eps ? w2[k] : 0.0, Si = w2[k] > eps ? 1.0 / w2[k] : 0.0; double vb = 0;
    for (int j = 0; j < n; ++j) { lin_jac[(size_t)k * n + j] = std::sqrt(S) * V2[(size_t)j * n + k]; vb += V2[(size_t)j * n + k] * br[j]; }
    lin_res[k] = std::sqrt(Si) * vb;
  }
  // keep_block_data in the next window's numbering (addr_shift, Estimator.cpp:2583-2597)
  for (int k = 1; k < W; ++k) for (int i = 0; i < 7; ++i) x0_pose[(size_t)7 * (k - 1) + i] = x[(size_t)P.na * k + i];
  for (int c = 0; c < 9; ++c) x0_sb[c] = x[(size_t)P.na + 7 + c];
  return 0;
}

int go_problem_solve(void* h, const void* options, int mode, int nthreads, void* summary_out, void* iter_log, int iter_cap,
                     double* step_log, int64_t step_cap) {
  Problem& P = *(Problem*)h;
  const Opt& o = *(const Opt*)options;
  P.mode = mode; P.nthreads = nthreads > 0 ? nthreads : 1;
  layout(P);
  const int n = P.n, na = P.W * P.na;
  Summary S; std::memset(&S, 0, sizeof(S)); S.termination = 1;
  std::vector<Iter> iters; std::vector<double> steps;
  std::vector<double> x = P.x, cand(na), proj(na), best = P.x;
  std::vector<double> gradient(n), scale(n, 1.0), diag(n), grad_d(n), gn(n), step(n), delta(n), tmp(n), y(n), neg(n), rhs(n);
  std::vector<double> A((size_t)n * n), model_res;
  double x_cost = 0, cand_cost = 0, minimum_cost = std::numeric_limits<double>::max(), model_cost_change = 0, x_norm = -1;
  double radius = o.initial_trust_region_radius, mu = 1e-8, alpha = 0, dogleg_step_norm = 0; bool reuse = false; int ninvalid = 0;
  std::vector<double> b0(n), b1(n); double sB[4] = {0, 0, 0, 0}, sg[2] = {0, 0}; bool sub1d = false;
  auto dotv = [&](const std::vector<double>& a, const std::vector<double>& b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; };
  auto nrm = [&](const std::vector<double>& a) { return std::sqrt(dotv(a, a)); };
  // step evaluator
  double se_min, se_cur, se_ref, se_cand, se_acc_ref = 0, se_acc_cand = 0; int se_n = 0; const int se_max = o.use_nonmonotonic_steps ? o.max_consecutive_nonmonotonic_steps : 0;

  auto eval_grad_jac = [&](Iter& it, bool first) -> bool {
    S.num_evaluations++; S.num_jacobian_evaluations++;
    if (!evaluate(P, x.data(), true, &x_cost)) return false;
    std::fill(gradient.begin(), gradient.end(), 0.0); left_multiply(P, P.r.data(), gradient.data());
    it.cost = x_cost;
    if (o.jacobi_scaling) {
      if (first) { sq_col_norm(P, scale.data()); for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(scale[i])); }
      scale_columns(P, scale.data());
    }
    for (int i = 0; i < n; ++i) neg[i] = -gradient[i];
    plus(P, x.data(), neg.data(), proj.data());
    double mx = 0, s2 = 0; for (int i = 0; i < na; ++i) { double d = x[i] - proj[i]; mx = std::max(mx, std::fabs(d)); s2 += d * d; }
    it.gradient_max_norm = mx; it.gradient_norm = std::sqrt(s2);
    return true;
  };
  auto traditional = [&]() {
    const double gnorm = nrm(grad_d), gnn = nrm(gn);
    if (gnn <= radius) { step = gn; dogleg_step_norm = gnn; for (int i = 0; i < n; ++i) step[i] /= diag[i]; return; }
    if (gnorm * alpha >= radius) { for (int i = 0; i < n; ++i) step[i] = -(radius / gnorm) * grad_d[i]; dogleg_step_norm = radius; for (int i = 0; i < n; ++i) step[i] /= diag[i]; return; }
    const double b_dot_a = -alpha * dotv(grad_d, gn), a_sq = std::pow(alpha * gnorm, 2.0), bma = a_sq - 2 * b_dot_a + std::pow(gnn, 2);
    const double c = b_dot_a - a_sq, d = std::sqrt(c * c + bma * (std::pow(radius, 2.0) - a_sq));
    const double beta = (c <= 0) ? (d - c) / bma : (radius * radius - a_sq) / (d + c);
    for (int i = 0; i < n; ++i) step[i] = (-alpha * (1.0 - beta)) * grad_d[i] + beta * gn[i];
    dogleg_step_norm = nrm(step); for (int i = 0; i < n; ++i) step[i] /= diag[i];
  };
  auto subspace = [&]() {
    const double gnn = nrm(gn);
    if (gnn <= radius) { step = gn; dogleg_step_norm = gnn; for (int i = 0; i < n; ++i) step[i] /= diag[i]; return; }
    if (sub1d) { const double gnorm = nrm(grad_d); for (int i = 0; i < n; ++i) step[i] = -(radius / gnorm) * grad_d[i]; dogleg_step_norm = radius; for (int i = 0; i < n; ++i) step[i] /= diag[i]; return; }
    const double detB = sB[0] * sB[3] - sB[1] * sB[2], trB = sB[0] + sB[3], r2 = radius * radius;
    const double Ba[4] = {sB[3], -sB[1], -sB[2], sB[0]}; const double bg[2] = {Ba[0] * sg[0] + Ba[1] * sg[1], Ba[2] * sg[0] + Ba[3] * sg[1]};
    double poly[5] = {r2, 2.0 * r2 * trB, r2 * (trB * trB + 2.0 * detB) - (sg[0] * sg[0] + sg[1] * sg[1]), -2.0 * ((sg[0] * bg[0] + sg[1] * bg[1]) - r2 * detB * trB), r2 * detB * detB - (bg[0] * bg[0] + bg[1] * bg[1])};
    std::vector<double> roots; double mn[2] = {0, 0}; bool found = false; double bestf = std::numeric_limits<double>::max();
    if (poly_real_parts(poly, &roots)) for (double yr : roots) {
      const double a = sB[0] + yr, b = sB[1], c = sB[2], d = sB[3] + yr, det = a * d - b * c; if (det == 0) continue;
      const double xi[2] = {-(d * sg[0] - b * sg[1]) / det, -(-c * sg[0] + a * sg[1]) / det};
      const double xn = std::sqrt(xi[0] * xi[0] + xi[1] * xi[1]);
      if (xn > 0 && std::isfinite(xn)) { const double v[2] = {radius / xn * xi[0], radius / xn * xi[1]};
        const double f = 0.5 * (v[0] * (sB[0] * v[0] + sB[1] * v[1]) + v[1] * (sB[2] * v[0] + sB[3] * v[1])) + sg[0] * v[0] + sg[1] * v[1];
        found = true; if (f < bestf) { bestf = f; mn[0] = xi[0]; mn[1] = xi[1]; } }
    }
    if (!found) { traditional(); return; }
    const double gm[2] = {sB[0] * mn[0] + sB[1] * mn[1] + sg[0], sB[2] * mn[0] + sB[3] * mn[1] + sg[1]};
    const double cosang = -(mn[0] * gm[0] + mn[1] * gm[1]) / (std::sqrt(mn[0] * mn[0] + mn[1] * mn[1]) * std::sqrt(gm[0] * gm[0] + gm[1] * gm[1]));
    if (cosang < 0.99) { traditional(); return; }
    for (int i = 0; i < n; ++i) step[i] = b0[i] * mn[0] + b1[i] * mn[1];
    dogleg_step_norm = radius; for (int i = 0; i < n; ++i) step[i] /= diag[i];
  };
  auto subspace_model = [&]() -> bool {
    const double n0 = nrm(grad_d), n1 = nrm(gn); const std::vector<double>& f = n0 >= n1 ? grad_d : gn; const std::vector<double>& s = n0 >= n1 ? gn : grad_d; const double nf = std::max(n0, n1);
    if (!(nf > 0)) return false;
    for (int i = 0; i < n; ++i) b0[i] = f[i] / nf;
    double pr = dotv(b0, s); for (int i = 0; i < n; ++i) b1[i] = s[i] - pr * b0[i];
    pr = dotv(b0, b1); for (int i = 0; i < n; ++i) b1[i] -= pr * b0[i];
    const double ns = nrm(b1);
    if (ns <= nf * 4.0 * std::numeric_limits<double>::epsilon()) { sub1d = true; return true; }
    sub1d = false; for (int i = 0; i < n; ++i) b1[i] /= ns;
    sg[0] = dotv(b0, grad_d); sg[1] = dotv(b1, grad_d);
    std::vector<double> t0(n), t1(n), j0(P.nrows, 0.0), j1(P.nrows, 0.0);
    for (int i = 0; i < n; ++i) { t0[i] = b0[i] / diag[i]; t1[i] = b1[i] / diag[i]; }
    right_multiply(P, t0.data(), j0.data()); right_multiply(P, t1.data(), j1.data());
    sB[0] = dotv(j0, j0); sB[1] = dotv(j0, j1); sB[2] = sB[1]; sB[3] = dotv(j1, j1);
    return true;
  };
  // LevenbergMarquardtStrategy::ComputeStep (ceres.tgz::internal/ceres/levenberg_marquardt_strategy.cc:69-141)
  const int strategy = o.reserved; double lm_decrease = 2.0;
  auto lm_step = [&]() -> int {
    if (!reuse) { sq_col_norm(P, diag.data()); for (int i = 0; i < n; ++i) diag[i] = std::min(std::max(diag[i], o.min_lm_diagonal), o.max_lm_diagonal); }
    reuse = true;
    std::vector<double> lmd(n); for (int i = 0; i < n; ++i) lmd[i] = std::sqrt(diag[i] / radius);
    S.num_linear_solves++;
    bool ok;
    if (strategy == 2) {
      const int64_t m = P.nrows + n;
      std::vector<double> Aug((size_t)m * n, 0.0), baug((size_t)m, 0.0);
      for (int64_t i = 0; i < P.nrows; ++i) { for (int64_t p = P.rowptr[i]; p < P.rowptr[i + 1]; ++p) Aug[(size_t)i * n + P.col[p]] += P.val[p]; baug[i] = P.r[i]; }
      for (int i = 0; i < n; ++i) Aug[(size_t)(P.nrows + i) * n + i] = lmd[i];
      ok = dense_qr_solve(Aug, m, n, baug, y.data());
    } else {
      std::fill(rhs.begin(), rhs.end(), 0.0); left_multiply(P, P.r.data(), rhs.data());
      std::fill(A.begin(), A.end(), 0.0);
      for (int64_t i = 0; i < P.nrows; ++i) for (int64_t p = P.rowptr[i]; p < P.rowptr[i + 1]; ++p) for (int64_t q = P.rowptr[i]; q < P.rowptr[i + 1]; ++q) A[(size_t)P.col[p] * n + P.col[q]] += P.val[p] * P.val[q];
      for (int i = 0; i < n; ++i) A[(size_t)i * n + i] += lmd[i] * lmd[i];
      ok = dense_cholesky_solve(A, n, rhs.data(), y.data());
    }
    if (!ok) return 1;
    for (int i = 0; i < n; ++i) step[i] = -y[i];
    return 0;
  };
  auto compute_step = [&]() -> int {
    if (strategy != 0) return lm_step();
    if (reuse) { if (o.dogleg_type == 0) traditional(); else subspace(); return 0; }
    reuse = true;
    sq_col_norm(P, diag.data());
    for (int i = 0; i < n; ++i) diag[i] = std::sqrt(std::min(std::max(diag[i], o.min_lm_diagonal), o.max_lm_diagonal));
    std::fill(grad_d.begin(), grad_d.end(), 0.0); left_multiply(P, P.r.data(), grad_d.data());
    rhs = grad_d;                                   // J^T r (scaled J)
    for (int i = 0; i < n; ++i) grad_d[i] /= diag[i];
    { std::vector<double> Jg(P.nrows, 0.0); for (int i = 0; i < n; ++i) tmp[i] = grad_d[i] / diag[i]; right_multiply(P, tmp.data(), Jg.data()); alpha = dotv(grad_d, grad_d) / dotv(Jg, Jg); }
    // J^T J (dense) once; regulariser added per attempt
    std::vector<double> JtJ((size_t)n * n, 0.0);
    for (int64_t i = 0; i < P.nrows; ++i) for (int64_t p = P.rowptr[i]; p < P.rowptr[i + 1]; ++p) for (int64_t q = P.rowptr[i]; q < P.rowptr[i + 1]; ++q) JtJ[(size_t)P.col[p] * n + P.col[q]] += P.val[p] * P.val[q];
    bool ok = false;
    while (mu < 1.0) {
      A = JtJ; const double sm = std::sqrt(mu);
      for (int i = 0; i < n; ++i) { const double lm = diag[i] * sm; A[(size_t)i * n + i] += lm * lm; }
      S.num_linear_solves++;
      if (dense_cholesky_solve(A, n, rhs.data(), y.data())) { ok = true; break; }
      mu *= 10.0;
    }
    if (!ok) return 1;
    for (int i = 0; i < n; ++i) gn[i] = -diag[i] * y[i];
    if (o.dogleg_type == 0) traditional(); else { if (!subspace_model()) return 1; subspace(); }
    return 0;
  };

  Iter it; std::memset(&it, 0, sizeof(it)); it.iteration = 0;
  if (!eval_grad_jac(it, true)) { S.termination = 2; snprintf(S.message, sizeof(S.message), "Residual and Jacobian evaluation failed."); goto done; }
  S.initial_cost = x_cost; S.final_cost = x_cost; it.step_is_valid = 1; it.step_is_successful = 1;
  se_min = se_cur = se_ref = se_cand = x_cost;
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it.step_is_successful) { ++S.num_successful_steps; if (x_cost < minimum_cost) { minimum_cost = x_cost; best = x; S.final_cost = x_cost; } } else ++S.num_unsuccessful_steps;
    it.trust_region_radius = radius; it.mu = mu; iters.push_back(it);
    if (it.iteration >= o.max_num_iterations) { S.termination = 1; snprintf(S.message, sizeof(S.message), "Maximum number of iterations reached."); break; }
    if (it.step_is_successful && it.gradient_max_norm <= o.gradient_tolerance) { S.termination = 0; snprintf(S.message, sizeof(S.message), "Gradient tolerance reached."); break; }
    if (radius <= o.min_trust_region_radius) { S.termination = 0; snprintf(S.message, sizeof(S.message), "Minimum trust region radius reached."); break; }
    const Iter prev = it; std::memset(&it, 0, sizeof(it)); it.iteration = prev.iteration + 1;
    const int rc = compute_step(); bool valid = false;
    if (rc == 0) {
      model_res.assign(P.nrows, 0.0); right_multiply(P, step.data(), model_res.data());
      double s = 0; for (int64_t i = 0; i < P.nrows; ++i) s += model_res[i] * (P.r[i] + model_res[i] / 2.0);
      model_cost_change = -s; valid = model_cost_change > 0.0;
      if (valid) { for (int i = 0; i < n; ++i) delta[i] = step[i] * scale[i]; ninvalid = 0; }
    }
    it.step_is_valid = valid;
    if (!valid) {
      if (++ninvalid >= o.max_num_consecutive_invalid_steps) { S.termination = 2; snprintf(S.message, sizeof(S.message), "Number of consecutive invalid steps more than max"); break; }
      if (strategy != 0) { radius /= lm_decrease; lm_decrease *= 2.0; reuse = true; } else { mu *= 10.0; reuse = false; }
      it.cost = x_cost; it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm; continue;
    }
    steps.insert(steps.end(), delta.begin(), delta.end());
    plus(P, x.data(), delta.data(), cand.data());
    {
      // cost-only evaluation at the candidate must not clobber the cached jacobian / residuals of x
      std::vector<double> r_keep = P.r;
      S.num_evaluations++;
      if (!evaluate(P, cand.data(), false, &cand_cost)) cand_cost = std::numeric_limits<double>::max();
      P.r.swap(r_keep);
    }
    { double s2 = 0; for (int i = 0; i < na; ++i) { double d = x[i] - cand[i]; s2 += d * d; } it.step_norm = std::sqrt(s2); }
    if (it.step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { S.termination = 0; snprintf(S.message, sizeof(S.message), "Parameter tolerance reached."); break; }
    it.cost_change = x_cost - cand_cost;
    if (std::fabs(it.cost_change) <= o.function_tolerance * x_cost) { S.termination = 0; snprintf(S.message, sizeof(S.message), "Function tolerance reached."); break; }
    {
      double q;
      if (cand_cost >= std::numeric_limits<double>::max()) q = std::numeric_limits<double>::lowest();
      else q = std::max((se_cur - cand_cost) / model_cost_change, (se_ref - cand_cost) / (se_acc_ref + model_cost_change));
      it.relative_decrease = q;
    }
    if (it.relative_decrease > o.min_relative_decrease) {
      x = cand; { double s2 = 0; for (int i = 0; i < na; ++i) s2 += x[i] * x[i]; x_norm = std::sqrt(s2); }
      if (!eval_grad_jac(it, false)) { S.termination = 2; snprintf(S.message, sizeof(S.message), "Residual and Jacobian evaluation failed."); break; }
      it.step_is_successful = 1;
      if (strategy != 0) {   // LevenbergMarquardtStrategy::StepAccepted (:143-150)
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
        radius = std::min(o.max_trust_region_radius, radius); lm_decrease = 2.0;
      } else {
        if (it.relative_decrease < 0.25) radius *= 0.5;
        if (it.relative_decrease > 0.75) radius = std::max(radius, 3.0 * dogleg_step_norm);
        mu = std::max(1e-8, 2.0 * mu / 10.0);
      }
      reuse = false;
      se_cur = cand_cost; se_acc_cand += model_cost_change; se_acc_ref += model_cost_change;
      if (se_cur < se_min) { se_min = se_cur; se_n = 0; se_cand = se_cur; se_acc_cand = 0.0; }
      else { ++se_n; if (se_cur > se_cand) { se_cand = se_cur; se_acc_cand = 0.0; } }
      if (se_n == se_max) { se_ref = se_cand; se_acc_ref = se_acc_cand; }
    } else {
      it.step_is_successful = 0; if (strategy != 0) { radius /= lm_decrease; lm_decrease *= 2.0; } else radius *= 0.5; reuse = true; it.cost = cand_cost; it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;
    }
  }
done:
  P.x = best;
  S.num_iterations = (int)iters.size(); S.num_valid_steps = (int)(steps.size() / (size_t)n);
  if (summary_out) std::memcpy(summary_out, &S, sizeof(S));
  if (iter_log) std::memcpy(iter_log, iters.data(), sizeof(Iter) * std::min<size_t>(iters.size(), (size_t)iter_cap));
  if (step_log) std::memcpy(step_log, steps.data(), sizeof(double) * std::min<size_t>(steps.size(), (size_t)step_cap));
  return S.termination;
}

}  // extern "C"
