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
#include "marg.h"

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

struct DenseAdder { static constexpr bool lower_only = false; double* H; int n; inline void operator()(int i, int j, double v) const { H[(size_t)i * n + j] += v; } };
// lower-band storage (solver.h BandMat): only i >= j is stored
// A non-zero contribution outside the declared half bandwidth is an ERROR (a wrong band would silently give a wrong J^T J):
// it raises *dropped, and the evaluation fails.
struct BandAdder {
  static constexpr bool lower_only = true;            // only i >= j is stored: callers skip the upper triangle
  double* a; int hb; int* dropped;
  inline void operator()(int i, int j, double v) const {
    if (i < j) return;
    if (i - j <= hb) a[(size_t)i * (hb + 1) + (j - i + hb)] += v;
    else if (v != 0.0) *dropped = 1;
  }
};

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
      if (Adder::lower_only) { for (int j = 0; j < nz; ++j) if (col[j] <= col[i]) add(col[i], col[j], val[i] * val[j]); }
      else for (int j = 0; j < nz; ++j) add(col[i], col[j], val[i] * val[j]);
    }
  }
}

}  // namespace

struct glio_host_factor_set {
  std::vector<Prior> priors;
  std::vector<Between> betweens;
  std::vector<Range> ranges;
  // the marginalisation prior of the previous window (MarginalizationFactor, MarginalizationFactor.cpp:232-330), in
  // information form: cost = c0 + b.dx + 0.5 dx.A.dx with dx as MarginalizationFactor::Evaluate builds it
  bool has_marg = false;
  glio_marg_prior marg;
  struct Blk { int pidx, size, kf, part; };          // prior index, width, keyframe, 0 t / 1 q / 2 speed-bias
  std::vector<Blk> marg_blocks;
  std::vector<std::pair<int, int>> marg_pairs;       // block pairs (I >= J) whose A_info block is not identically zero
};

namespace {

// dx of MarginalizationFactor::Evaluate (:239-258) and the per-block Jacobians d dx / d(parameter):
//   kind 0: Ceres tangent (ambient Jacobian of the reference, 2 Qleft(q0^-1) rows x,y,z, times QuaternionParameterization's
//           plus-Jacobian) - the solve path;   kind 1: ambient x,y,z quaternion columns - the marginalisation path (:9-17)
struct MargLin { std::vector<double> dx; std::vector<double> P; /* per block: size x size row-major, concatenated */ std::vector<int> poff; };
inline void marg_linearise(const glio_host_factor_set& S, const double* poses, const double* speed_bias, int kind, MargLin& L) {
  const glio_marg_prior& M = S.marg;
  L.dx.assign(M.n, 0.0); L.P.clear(); L.poff.clear();
  for (const auto& B : S.marg_blocks) {
    L.poff.push_back((int)L.P.size());
    L.P.resize(L.P.size() + (size_t)B.size * B.size, 0.0);
    double* P = &L.P[L.poff.back()];
    const double* x0 = &M.x0_pose[(size_t)7 * B.kf];
    if (B.part == 0) {
      for (int k = 0; k < 3; ++k) { L.dx[B.pidx + k] = poses[7 * B.kf + k] - x0[k]; P[3 * k + k] = 1.0; }
    } else if (B.part == 2) {
      for (int k = 0; k < 9; ++k) { L.dx[B.pidx + k] = (speed_bias ? speed_bias[9 * B.kf + k] : M.x0_sb[k]) - M.x0_sb[k]; P[9 * k + k] = 1.0; }
    } else {
      const double* q = poses + 7 * B.kf + 3;
      double q0i[4]; glio::h_qinv(x0 + 3, q0i);
      double e[4]; h_qmul(q0i, q, e);
      const double nrm = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3]);
      const double sgn = e[0] >= 0 ? 1.0 : -1.0;       // the reference tests w of the un-normalised product; same sign
      for (int k = 0; k < 3; ++k) L.dx[B.pidx + k] = sgn * 2.0 * e[1 + k] / nrm;
      for (int c = 0; c < 3; ++c) {
        double ek[4] = {0, 0, 0, 0}; ek[1 + c] = 1.0;
        double a[4]; h_qmul(q0i, ek, a);
        double col[4];
        if (kind == 0) h_qmul(a, q, col);               // d/d delta of q0^-1 (x) ([.., delta] (x) q)
        else { col[0] = a[0]; col[1] = a[1]; col[2] = a[2]; col[3] = a[3]; }   // d/d q_{x,y,z} of q0^-1 (x) q
        for (int k = 0; k < 3; ++k) P[3 * k + c] = sgn * 2.0 * col[1 + k];
      }
    }
  }
}

}  // namespace

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
      double J[15 * 15]; std::memset(J, 0, sizeof(double) * 15 * nt);
      for (int k = 0; k < 3; ++k) J[(size_t)k * nt + k] = f.sw[k];
      for (int col = 0; col < 3; ++col) {     // d/d delta: 2 vec(q0c (x) (0,e_col) (x) q)
        double ek[4] = {0, 0, 0, 0}; ek[1 + col] = 1.0;
        double a[4], b[4]; h_qmul(q0c, ek, a); h_qmul(a, q, b);
        for (int k = 0; k < 3; ++k) J[(size_t)(3 + k) * nt + 3 + col] = f.sw[3 + k] * 2.0 * b[1 + k];
      }
      if (sb) for (int k = 0; k < 9; ++k) J[(size_t)(6 + k) * nt + 6 + k] = f.sw[6 + k];
      const int off[1] = {nt * f.kf}, width[1] = {nt}; const double* Jp[1] = {J};
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
      double Ji[15 * 15], Jj[15 * 15];                  // nt <= 15: on the stack (this loop runs once per keyframe per evaluation)
      std::memset(Ji, 0, sizeof(double) * 15 * nt); std::memset(Jj, 0, sizeof(double) * 15 * nt);
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
      const int off[2] = {nt * f.i, nt * f.j}, width[2] = {nt, nt}; const double* Jp[2] = {Ji, Jj};
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
  if (S.has_marg) {
    const glio_marg_prior& M = S.marg;
    MargLin L; marg_linearise(S, poses, speed_bias, 0, L);
    const int n = M.n;
    std::vector<double> gdx(M.b_info);
    double quad = 0, lin = 0;
    for (int i = 0; i < n; ++i) {
      const double* row = &M.A_info[(size_t)i * n];
      double sacc = 0;
      for (int j = 0; j < n; ++j) if (row[j] != 0.0) sacc += row[j] * L.dx[j];
      gdx[i] += sacc; quad += L.dx[i] * sacc; lin += M.b_info[i] * L.dx[i];
    }
    c += M.c0 + lin + 0.5 * quad;
    if (want_jac) {
      auto toff = [&](const glio_host_factor_set::Blk& B) { return nt * B.kf + (B.part == 0 ? 0 : (B.part == 1 ? 3 : 6)); };
      for (size_t bi = 0; bi < S.marg_blocks.size(); ++bi) {
        const auto& B = S.marg_blocks[bi];
        if (B.part == 2 && !sb) continue;
        const double* P = &L.P[L.poff[bi]];
        for (int a = 0; a < B.size; ++a) { double v = 0; for (int k = 0; k < B.size; ++k) v += P[B.size * k + a] * gdx[B.pidx + k]; g[toff(B) + a] += v; }
      }
      for (const auto& pr : S.marg_pairs) {
        const auto& BI = S.marg_blocks[pr.first]; const auto& BJ = S.marg_blocks[pr.second];
        if ((BI.part == 2 || BJ.part == 2) && !sb) continue;
        const double* PI = &L.P[L.poff[pr.first]]; const double* PJ = &L.P[L.poff[pr.second]];
        double T[81], Hb[81];
        for (int a = 0; a < BI.size; ++a) for (int bcol = 0; bcol < BJ.size; ++bcol) { double v = 0; for (int k = 0; k < BJ.size; ++k) v += M.A_info[(size_t)(BI.pidx + a) * n + BJ.pidx + k] * PJ[BJ.size * k + bcol]; T[a * BJ.size + bcol] = v; }
        for (int a = 0; a < BI.size; ++a) for (int bcol = 0; bcol < BJ.size; ++bcol) { double v = 0; for (int k = 0; k < BI.size; ++k) v += PI[BI.size * k + a] * T[k * BJ.size + bcol]; Hb[a * BJ.size + bcol] = v; }
        const int oi = toff(BI), oj = toff(BJ);
        for (int a = 0; a < BI.size; ++a) for (int bcol = 0; bcol < BJ.size; ++bcol) {
          add(oi + a, oj + bcol, Hb[a * BJ.size + bcol]);
          if (pr.first != pr.second) add(oj + bcol, oi + a, Hb[a * BJ.size + bcol]);
        }
      }
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
  int dropped = 0;
  const int rc = hf_evaluate_impl(user, K, poses, speed_bias, want_jac, BandAdder{Hband, hb, &dropped}, g, cost);
  return dropped ? GLIO_ERR_ARG : rc;       // half_bandwidth does not cover a coupling these factors create
}

}  // extern "C"


// ------------------------------------------------------------------------------------------------------------------
// marginalisation path of the stand-ins (glio_host_marg_fn): what MarginalizationInfo accumulates for the factors that are
// NOT LiDAR (Estimator.cpp:2464-2537): the previous prior, every prior on KF0, the between (IMU-like) factor KF0 -> KF1.
// Jacobians are the AMBIENT x,y,z quaternion columns (ResidualBlockInfo::Evaluate + ThreadsConstructA's rightCols(3),
// MarginalizationFactor.cpp:9-17), marginalisation ordering of marg.h.  A (N x N, full symmetric) and b are accumulated into.
// ------------------------------------------------------------------------------------------------------------------
namespace {
inline void rot_conj_du(const double q[4], const double v[3], int c, double o[3]) {
  // d/d q_c (c = x,y,z) of h_qrot(conj(q), v) = v - 2w(u x v) + 2 u x (u x v), u = q.xyz
  const double u[3] = {q[1], q[2], q[3]};
  double ec[3] = {0, 0, 0}; ec[c] = 1.0;
  double ecv[3], uv[3], t1[3], t2[3];
  glio::h_cross3(ec, v, ecv); glio::h_cross3(u, v, uv); glio::h_cross3(ec, uv, t1); glio::h_cross3(u, ecv, t2);
  for (int k = 0; k < 3; ++k) o[k] = -2.0 * q[0] * ecv[k] + 2.0 * (t1[k] + t2[k]);
}
inline void marg_accumulate(int m, const double* r, int nb, const int* off, const int* width, const double* const* J, double* A, double* b, int N) {
  for (int k = 0; k < m; ++k) {
    int col[64]; double val[64]; int nz = 0;
    for (int bi = 0; bi < nb; ++bi) for (int p = 0; p < width[bi]; ++p) { const double v = J[bi][k * width[bi] + p]; if (v != 0.0 && nz < 64) { col[nz] = off[bi] + p; val[nz] = v; ++nz; } }
    for (int i = 0; i < nz; ++i) { b[col[i]] += val[i] * r[k]; for (int j = 0; j < nz; ++j) A[(size_t)col[i] * N + col[j]] += val[i] * val[j]; }
  }
}
}  // namespace

extern "C" int glio_hf_marg_evaluate(void* user, int W, const double* poses, const double* speed_bias, double* A, double* b) {
  const glio_host_factor_set& S = *(const glio_host_factor_set*)user;
  if (W < 2 || !poses || !speed_bias || !A || !b) return GLIO_ERR_ARG;
  const int N = 6 * W + 18;
  for (const Prior& f : S.priors) {
    if (f.kf != 0) continue;                                  // only factors that touch the dropped keyframe
    const double* t = poses; const double* q = t + 3;
    double r[15], J[15 * 15]; std::memset(J, 0, sizeof(J));
    for (int k = 0; k < 3; ++k) { r[k] = f.sw[k] * (t[k] - f.t0[k]); J[15 * k + k] = f.sw[k]; }
    double q0c[4]; qconj(f.q0, q0c);
    double e[4]; h_qmul(q0c, q, e);
    for (int k = 0; k < 3; ++k) r[3 + k] = f.sw[3 + k] * 2.0 * e[1 + k];
    for (int c = 0; c < 3; ++c) { double ek[4] = {0, 0, 0, 0}; ek[1 + c] = 1.0; double a[4]; h_qmul(q0c, ek, a); for (int k = 0; k < 3; ++k) J[15 * (3 + k) + 3 + c] = f.sw[3 + k] * 2.0 * a[1 + k]; }
    for (int k = 0; k < 9; ++k) { r[6 + k] = f.sw[6 + k] * (speed_bias[k] - f.sb0[k]); J[15 * (6 + k) + 6 + k] = f.sw[6 + k]; }
    const int off[1] = {0}, width[1] = {15}; const double* Jp[1] = {J};
    marg_accumulate(15, r, 1, off, width, Jp, A, b, N);
  }
  for (const Between& f : S.betweens) {
    if (!(f.i == 0 && f.j == 1)) continue;                    // the IMU factor of the dropped keyframe (Estimator.cpp:2521-2535)
    const double* ti = poses; const double* qi = ti + 3; const double* tj = poses + 7; const double* qj = tj + 3;
    const double* si = speed_bias; const double* sj = speed_bias + 9;
    double qic[4]; qconj(qi, qic);
    double Ri[9]; rot_mat(qi, Ri);
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
    double Ji[15 * 15], Jj[15 * 15]; std::memset(Ji, 0, sizeof(Ji)); std::memset(Jj, 0, sizeof(Jj));
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) {
      const double RiT = Ri[3 * c + a];
      Ji[15 * a + c] = -f.sw[a] * RiT; Jj[15 * a + c] = f.sw[a] * RiT; Ji[15 * a + 6 + c] = -f.sw[a] * RiT * f.dt;
      Ji[15 * (6 + a) + 6 + c] = -f.sw[6 + a] * RiT; Jj[15 * (6 + a) + 6 + c] = f.sw[6 + a] * RiT;
    }
    for (int c = 0; c < 3; ++c) {
      double dp_[3], dv_[3]; rot_conj_du(qi, d, c, dp_); rot_conj_du(qi, dvv, c, dv_);
      for (int a = 0; a < 3; ++a) { Ji[15 * a + 3 + c] = f.sw[a] * dp_[a]; Ji[15 * (6 + a) + 3 + c] = f.sw[6 + a] * dv_[a]; }
      double ek[4] = {0, 0, 0, 0}; ek[1 + c] = 1.0;
      double a1[4], a2[4], a3[4];
      h_qmul(qic, ek, a1); h_qmul(dqc, a1, a2);                           // d/d qj_c: dqc (x) qic (x) e_c
      for (int k = 0; k < 3; ++k) Jj[15 * (3 + k) + 3 + c] = f.sw[3 + k] * 2.0 * a2[1 + k];
      h_qmul(ek, qj, a1); h_qmul(dqc, a1, a3);                            // d/d qi_c: dqc (x) (-e_c) (x) qj
      for (int k = 0; k < 3; ++k) Ji[15 * (3 + k) + 3 + c] = -f.sw[3 + k] * 2.0 * a3[1 + k];
    }
    for (int k = 0; k < 6; ++k) { Ji[15 * (9 + k) + 9 + k] = -f.sw[9 + k]; Jj[15 * (9 + k) + 9 + k] = f.sw[9 + k]; }
    const int off[2] = {0, 15}, width[2] = {15, 15}; const double* Jp[2] = {Ji, Jj};
    marg_accumulate(15, r, 2, off, width, Jp, A, b, N);
  }
  if (S.has_marg) {
    const glio_marg_prior& M = S.marg;
    if (M.W != W) return GLIO_ERR_ARG;
    MargLin L; marg_linearise(S, poses, speed_bias, 1, L);
    const int n = M.n;
    std::vector<double> gdx(M.b_info);
    for (int i = 0; i < n; ++i) { const double* row = &M.A_info[(size_t)i * n]; double sacc = 0; for (int j = 0; j < n; ++j) if (row[j] != 0.0) sacc += row[j] * L.dx[j]; gdx[i] += sacc; }
    auto moff = [&](const glio_host_factor_set::Blk& B) { return glio::marg_index_t(B.kf) + (B.part == 0 ? 0 : (B.part == 1 ? 3 : 6)); };
    for (size_t bi = 0; bi < S.marg_blocks.size(); ++bi) {
      const auto& B = S.marg_blocks[bi]; const double* P = &L.P[L.poff[bi]];
      for (int a = 0; a < B.size; ++a) { double v = 0; for (int k = 0; k < B.size; ++k) v += P[B.size * k + a] * gdx[B.pidx + k]; b[moff(B) + a] += v; }
    }
    for (const auto& pr : S.marg_pairs) {
      const auto& BI = S.marg_blocks[pr.first]; const auto& BJ = S.marg_blocks[pr.second];
      const double* PI = &L.P[L.poff[pr.first]]; const double* PJ = &L.P[L.poff[pr.second]];
      double T[81], Hb[81];
      for (int a = 0; a < BI.size; ++a) for (int c = 0; c < BJ.size; ++c) { double v = 0; for (int k = 0; k < BJ.size; ++k) v += M.A_info[(size_t)(BI.pidx + a) * n + BJ.pidx + k] * PJ[BJ.size * k + c]; T[a * BJ.size + c] = v; }
      for (int a = 0; a < BI.size; ++a) for (int c = 0; c < BJ.size; ++c) { double v = 0; for (int k = 0; k < BI.size; ++k) v += PI[BI.size * k + a] * T[k * BJ.size + c]; Hb[a * BJ.size + c] = v; }
      const int oi = moff(BI), oj = moff(BJ);
      for (int a = 0; a < BI.size; ++a) for (int c = 0; c < BJ.size; ++c) {
        A[(size_t)(oi + a) * N + oj + c] += Hb[a * BJ.size + c];
        if (pr.first != pr.second) A[(size_t)(oj + c) * N + oi + a] += Hb[a * BJ.size + c];
      }
    }
  }
  return GLIO_OK;
}

extern "C" int glio_hf_set_marg_prior(glio_host_factor_set* s, const glio_marg_prior* p) {
  if (!s) return GLIO_ERR_ARG;
  s->has_marg = false; s->marg_blocks.clear(); s->marg_pairs.clear();
  if (!p) return GLIO_OK;
  s->marg = *p; s->has_marg = true;
  const int W = p->W;
  s->marg_blocks.push_back({0, 3, 0, 0}); s->marg_blocks.push_back({3, 3, 0, 1}); s->marg_blocks.push_back({6, 9, 0, 2});
  for (int k = 1; k <= W - 2; ++k) { s->marg_blocks.push_back({glio::prior_index_t(k), 3, k, 0}); s->marg_blocks.push_back({glio::prior_index_t(k) + 3, 3, k, 1}); }
  const int n = p->n;
  for (size_t I = 0; I < s->marg_blocks.size(); ++I) for (size_t J = 0; J <= I; ++J) {
    const auto& BI = s->marg_blocks[I]; const auto& BJ = s->marg_blocks[J];
    bool nz = false;
    for (int a = 0; a < BI.size && !nz; ++a) for (int c = 0; c < BJ.size; ++c) if (p->A_info[(size_t)(BI.pidx + a) * n + BJ.pidx + c] != 0.0) { nz = true; break; }
    if (nz) s->marg_pairs.push_back({(int)I, (int)J});
  }
  return GLIO_OK;
}

extern "C" int glio_hf_marg_half_bandwidth(const glio_host_factor_set* s) { return (s && s->has_marg) ? s->marg.half_bandwidth : 0; }
