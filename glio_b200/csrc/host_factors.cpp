// host_factors.cpp — analytic host-side (CPU, C++) factors that share the window problem with the LiDAR blocks.
//
// In the reference the non-LiDAR factors (ImuFactor, MarginalizationFactor, dd_psr_factor — SURVEY §2 rows 8-10)
// are host C++ ceres::CostFunctions and stay so by north_star.  Their data (pre-integration, GNSS epochs) are not
// reproducible here, so the tests / benchmark use three generic stand-ins with the same block structure:
//   prior   15 x [t,q,sb]         (marginalisation-prior-like, Estimator.cpp:2153-2158)
//   between 15 x [t,q,sb | t,q,sb] (IMU-like chain between consecutive keyframes, Estimator.cpp:2182-2192)
//   range    1 x [t,q]            (pseudorange-like, Estimator.cpp:1891-1900)
// They are evaluated in the tangent space of Ceres' QuaternionParameterization (q <- [cos|d|, sin|d|/|d| d] (x) q)
// and accumulated into the dense normal equations through the glio_host_factors_fn callback signature.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/glio_b200.h"
#include "hostmath.h"

namespace {

using glio::h_qmul; using glio::h_qrot;

struct Prior { int kf; double t0[3], q0[4], sb0[9], sw[15]; };
struct Between { int i, j; double dp[3], dq[4], dv[3], dt, sw[15]; };
struct Range { int kf; double lever[3], sat[3], rho, w; };

inline void qconj(const double q[4], double o[4]) { o[0] = q[0]; o[1] = -q[1]; o[2] = -q[2]; o[3] = -q[3]; }
inline void rot_mat(const double q[4], double R[9]) {
  double e[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int c = 0; c < 3; ++c) { double o[3]; h_qrot(q, e[c], o); R[0 + c] = o[0]; R[3 + c] = o[1]; R[6 + c] = o[2]; }
}
// M = A^T * skew(v) * s
inline void At_skew(const double A[9], const double v[3], double s, double M[9]) {
  const double S[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double a = 0; for (int k = 0; k < 3; ++k) a += A[3 * k + i] * S[3 * k + j]; M[3 * i + j] = s * a; }
}

struct DenseAdder { double* H; int n; inline void operator()(int i, int j, double v) const { H[(size_t)i * n + j] += v; } };
// lower-band storage (solver.h BandMat): only i >= j is stored
struct BandAdder { double* a; int hb; inline void operator()(int i, int j, double v) const { if (i >= j && i - j <= hb) a[(size_t)i * (hb + 1) + (j - i + hb)] += v; } };

// accumulate an m-row factor with dense tangent Jacobian blocks into H, g
template <class Adder>
inline void accumulate(int m, const double* r, int nb, const int* off, const int* width, const double* const* J /*m x width[b]*/,
                       const Adder& add, double* g) {
  // row by row over the non-zeros only (the factor Jacobians are mostly 3x3 blocks and diagonals)
  int col[64]; double val[64];
  for (int k = 0; k < m; ++k) {
    int nz = 0;
    for (int b = 0; b < nb; ++b)
      for (int p = 0; p < width[b]; ++p) {
        const double v = J[b][k * width[b] + p];
        if (v != 0.0 && nz < 64) { col[nz] = off[b] + p; val[nz] = v; ++nz; }
      }
    for (int i = 0; i < nz; ++i) {
      g[col[i]] += val[i] * r[k];
      for (int j = 0; j < nz; ++j) add(col[i], col[j], val[i] * val[j]);
    }
  }
}

}  // namespace

struct glio_host_factor_set {
  std::vector<Prior> priors;
  std::vector<Between> betweens;
  std::vector<Range> ranges;
};

extern "C" {

glio_host_factor_set* glio_hf_create(void) { return new glio_host_factor_set(); }
void glio_hf_destroy(glio_host_factor_set* s) { delete s; }

void glio_hf_add_prior(glio_host_factor_set* s, int kf, const double t0[3], const double q0[4], const double* sb0, const double sqrt_w[15]) {
  Prior f; f.kf = kf;
  for (int i = 0; i < 3; ++i) f.t0[i] = t0[i];
  for (int i = 0; i < 4; ++i) f.q0[i] = q0[i];
  for (int i = 0; i < 9; ++i) f.sb0[i] = sb0 ? sb0[i] : 0.0;
  for (int i = 0; i < 15; ++i) f.sw[i] = sqrt_w[i];
  s->priors.push_back(f);
}
void glio_hf_add_between(glio_host_factor_set* s, int i, int j, const double dp[3], const double dq[4], const double dv[3], double dt,
                         const double sqrt_w[15]) {
  Between f; f.i = i; f.j = j; f.dt = dt;
  for (int k = 0; k < 3; ++k) { f.dp[k] = dp[k]; f.dv[k] = dv[k]; }
  for (int k = 0; k < 4; ++k) f.dq[k] = dq[k];
  for (int k = 0; k < 15; ++k) f.sw[k] = sqrt_w[k];
  s->betweens.push_back(f);
}
void glio_hf_add_range(glio_host_factor_set* s, int kf, const double lever[3], const double sat[3], double rho, double w) {
  Range f; f.kf = kf; f.rho = rho; f.w = w;
  for (int k = 0; k < 3; ++k) { f.lever[k] = lever[k]; f.sat[k] = sat[k]; }
  s->ranges.push_back(f);
}

}  // extern "C"

template <class Adder>
static int hf_evaluate_impl(void* user, int W, const double* poses, const double* speed_bias, int want_jac, const Adder& add, double* g, double* cost) {
  const glio_host_factor_set& S = *(const glio_host_factor_set*)user;
  const bool sb = speed_bias != nullptr;
  const int nt = sb ? 15 : 6;
  (void)W;
  double c = 0;
  for (const Prior& f : S.priors) {
    const double* t = poses + 7 * f.kf; const double* q = t + 3;
    double r[15]; std::memset(r, 0, sizeof(r));
    for (int k = 0; k < 3; ++k) r[k] = f.sw[k] * (t[k] - f.t0[k]);
    double q0c[4]; qconj(f.q0, q0c);
    double e[4]; h_qmul(q0c, q, e);
    for (int k = 0; k < 3; ++k) r[3 + k] = f.sw[3 + k] * 2.0 * e[1 + k];
    if (sb) for (int k = 0; k < 9; ++k) r[6 + k] = f.sw[6 + k] * (speed_bias[9 * f.kf + k] - f.sb0[k]);
    for (int k = 0; k < 15; ++k) c += 0.5 * r[k] * r[k];
    if (want_jac) {
      std::vector<double> J((size_t)15 * nt, 0.0);
      for (int k = 0; k < 3; ++k) J[(size_t)k * nt + k] = f.sw[k];
      for (int col = 0; col < 3; ++col) {     // d/d delta: 2 vec(q0c (x) (0,e_col) (x) q)
        double ek[4] = {0, 0, 0, 0}; ek[1 + col] = 1.0;
        double a[4], b[4]; h_qmul(q0c, ek, a); h_qmul(a, q, b);
        for (int k = 0; k < 3; ++k) J[(size_t)(3 + k) * nt + 3 + col] = f.sw[3 + k] * 2.0 * b[1 + k];
      }
      if (sb) for (int k = 0; k < 9; ++k) J[(size_t)(6 + k) * nt + 6 + k] = f.sw[6 + k];
      const int off[1] = {nt * f.kf}, width[1] = {nt}; const double* Jp[1] = {J.data()};
      accumulate(15, r, 1, off, width, Jp, add, g);
    }
  }
  for (const Between& f : S.betweens) {
    const double* ti = poses + 7 * f.i; const double* qi = ti + 3;
    const double* tj = poses + 7 * f.j; const double* qj = tj + 3;
    double zero9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double* si = sb ? speed_bias + 9 * f.i : zero9; const double* sj = sb ? speed_bias + 9 * f.j : zero9;
    double qic[4]; qconj(qi, qic);
    double Ri[9]; rot_mat(qi, Ri);      // R_i ; R_i^T v = rotate by conj
    double d[3] = {tj[0] - ti[0] - si[0] * f.dt, tj[1] - ti[1] - si[1] * f.dt, tj[2] - ti[2] - si[2] * f.dt};
    double rp[3]; h_qrot(qic, d, rp);
    double r[15];
    for (int k = 0; k < 3; ++k) r[k] = f.sw[k] * (rp[k] - f.dp[k]);
    double dqc[4]; qconj(f.dq, dqc);
    double qij[4], e[4]; h_qmul(qic, qj, qij); h_qmul(dqc, qij, e);
    for (int k = 0; k < 3; ++k) r[3 + k] = f.sw[3 + k] * 2.0 * e[1 + k];
    double dvv[3] = {sj[0] - si[0], sj[1] - si[1], sj[2] - si[2]}, rv[3]; h_qrot(qic, dvv, rv);
    for (int k = 0; k < 3; ++k) r[6 + k] = f.sw[6 + k] * (rv[k] - f.dv[k]);
    for (int k = 0; k < 6; ++k) r[9 + k] = f.sw[9 + k] * (sj[3 + k] - si[3 + k]);
    for (int k = 0; k < 15; ++k) c += 0.5 * r[k] * r[k];
    if (want_jac) {
      std::vector<double> Ji((size_t)15 * nt, 0.0), Jj((size_t)15 * nt, 0.0);
      double M[9];
      // position rows
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
        const double RiT = Ri[3 * b + a];                           // (R_i^T)[a][b]
        Ji[(size_t)a * nt + b] = -f.sw[a] * RiT; Jj[(size_t)a * nt + b] = f.sw[a] * RiT;
        if (sb) Ji[(size_t)a * nt + 6 + b] = -f.sw[a] * RiT * f.dt;
      }
      At_skew(Ri, d, 2.0, M);
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Ji[(size_t)a * nt + 3 + b] = f.sw[a] * M[3 * a + b];
      // rotation rows: d/d delta_j col = 2 vec(dqc (x) qic (x) (0,e) (x) qj); d/d delta_i = - that
      for (int col = 0; col < 3; ++col) {
        double ek[4] = {0, 0, 0, 0}; ek[1 + col] = 1.0;
        double a1[4], a2[4], a3[4]; h_qmul(qic, ek, a1); h_qmul(a1, qj, a2); h_qmul(dqc, a2, a3);
        for (int k = 0; k < 3; ++k) { Jj[(size_t)(3 + k) * nt + 3 + col] = f.sw[3 + k] * 2.0 * a3[1 + k]; Ji[(size_t)(3 + k) * nt + 3 + col] = -f.sw[3 + k] * 2.0 * a3[1 + k]; }
      }
      // velocity rows
      At_skew(Ri, dvv, 2.0, M);
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
        const double RiT = Ri[3 * b + a];
        Ji[(size_t)(6 + a) * nt + 3 + b] = f.sw[6 + a] * M[3 * a + b];
        if (sb) { Ji[(size_t)(6 + a) * nt + 6 + b] = -f.sw[6 + a] * RiT; Jj[(size_t)(6 + a) * nt + 6 + b] = f.sw[6 + a] * RiT; }
      }
      if (sb) for (int k = 0; k < 6; ++k) { Ji[(size_t)(9 + k) * nt + 9 + k] = -f.sw[9 + k]; Jj[(size_t)(9 + k) * nt + 9 + k] = f.sw[9 + k]; }
      const int off[2] = {nt * f.i, nt * f.j}, width[2] = {nt, nt}; const double* Jp[2] = {Ji.data(), Jj.data()};
      accumulate(15, r, 2, off, width, Jp, add, g);
    }
  }
  for (const Range& f : S.ranges) {
    const double* t = poses + 7 * f.kf; const double* q = t + 3;
    double a[3]; h_qrot(q, f.lever, a);
    double p[3] = {a[0] + t[0] - f.sat[0], a[1] + t[1] - f.sat[1], a[2] + t[2] - f.sat[2]};
    const double pn = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    const double r[1] = {f.w * (pn - f.rho)};
    c += 0.5 * r[0] * r[0];
    if (want_jac) {
      const double ph[3] = {p[0] / pn, p[1] / pn, p[2] / pn};
      double cr[3]; glio::h_cross3(a, ph, cr);
      double J[6] = {f.w * ph[0], f.w * ph[1], f.w * ph[2], 2.0 * f.w * cr[0], 2.0 * f.w * cr[1], 2.0 * f.w * cr[2]};
      const int off[1] = {nt * f.kf}, width[1] = {6}; const double* Jp[1] = {J};
      accumulate(1, r, 1, off, width, Jp, add, g);
    }
  }
  *cost += c;
  return 0;
}

extern "C" {

// glio_host_factors_fn-compatible evaluation (user = glio_host_factor_set*): dense n x n H
int glio_hf_evaluate(void* user, int W, const double* poses, const double* speed_bias, int want_jac, double* H, double* g, double* cost) {
  const int n = W * (speed_bias ? 15 : 6);
  return hf_evaluate_impl(user, W, poses, speed_bias, want_jac, DenseAdder{H, n}, g, cost);
}
// glio_host_factors_band_fn-compatible evaluation: lower-band storage (batch problems)
int glio_hf_evaluate_band(void* user, int K, const double* poses, const double* speed_bias, int want_jac, double* Hband, int hb, double* g, double* cost) {
  return hf_evaluate_impl(user, K, poses, speed_bias, want_jac, BandAdder{Hband, hb}, g, cost);
}

}  // extern "C"
