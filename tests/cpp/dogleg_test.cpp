// dogleg_test.cpp — Ceres' own known-answer tests of DoglegStrategy (the reference vendors Ceres 2.0.0 as
// support_files/ceres-solver.tar.gz; ceres.tgz::internal/ceres/dogleg_strategy_test.cc:44-300), run against the
// product's host minimizer glio::TrustRegionDogleg (glio_b200/csrc/solver.{h,cpp}).  The fixtures and expectations are
// the reference's; what is checked is the first trust-region step from x = 0 on the linear least-squares problem
// r(x) = J x + r0 with the fixture's options (min_lm_diagonal = max_lm_diagonal = 1: no diagonal scaling; the test
// drives the strategy directly, so the minimizer-level Jacobi scaling is off).
// Prints "name ok|FAIL detail" lines; tests/test_dogleg_known_answers.py asserts on them.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

#include "../../glio_b200/csrc/solver.h"

using glio::BandMat;

static std::vector<double> first_step(const std::vector<double>& J /*6x6 row-major*/, const std::vector<double>& r0, int dogleg_type, double radius) {
  const int n = 6;
  std::vector<glio::ParamBlock> blocks{glio::ParamBlock{0, n, 0, n, false, nullptr}};
  glio::SolverOptions o;
  o.max_num_iterations = 1; o.dogleg_type = dogleg_type;
  o.initial_trust_region_radius = radius; o.max_trust_region_radius = radius;
  o.min_lm_diagonal = 1.0; o.max_lm_diagonal = 1.0; o.jacobi_scaling = false;
  o.function_tolerance = 0; o.gradient_tolerance = 0; o.parameter_tolerance = 0;
  glio::TrustRegionDogleg solver(blocks, o);
  glio::EvalFn eval = [&](const double* x, bool want_jac, double* cost, BandMat* H, double* g) -> bool {
    double r[6]; double c = 0;
    for (int i = 0; i < n; ++i) { r[i] = r0[i]; for (int j = 0; j < n; ++j) r[i] += J[i * n + j] * x[j]; c += 0.5 * r[i] * r[i]; }
    *cost = c;
    if (want_jac) {
      H->reset(n, n - 1);
      for (int a = 0; a < n; ++a) {
        double ga = 0; for (int i = 0; i < n; ++i) ga += J[i * n + a] * r[i];
        g[a] = ga;
        for (int b = 0; b <= a; ++b) { double h = 0; for (int i = 0; i < n; ++i) h += J[i * n + a] * J[i * n + b]; H->at(a, b) = h; }
      }
    }
    return true;
  };
  std::vector<double> x(n, 0.0);
  glio::SolverSummary S;
  solver.solve(x.data(), eval, &S);
  return S.steps.size() >= (size_t)n ? std::vector<double>(S.steps.begin(), S.steps.begin() + n) : std::vector<double>();
}

static double norm(const std::vector<double>& v) { double s = 0; for (double a : v) s += a * a; return std::sqrt(s); }
static void report(const char* name, bool ok, const std::vector<double>& x) {
  printf("%s %s", name, ok ? "ok" : "FAIL"); for (double v : x) printf(" %.17g", v); printf("\n");
}


// ---- ceres.tgz::internal/ceres/trust_region_minimizer_test.cc:60-300: Powell's singular function with a subset of the
// columns held fixed, TEST(TrustRegionMinimizer, PowellsSingularFunctionUsingDogleg).  The evaluator (residuals and the
// Jacobian exactly as the test writes them, including its (1 - x4) / (x1 - 1) factors in the f4 column entries) is
// transcribed literally; g = J^T r as the strategy computes it.  Default Solver::Options: 50 iterations, traditional
// dogleg, Jacobi scaling on; tolerances 1e-26, radius 1e4 / 1e20, lm diagonal 1e-6 / 1e32 (:218-236).
static bool powell_case(bool c1, bool c2, bool c3, bool c4, double out[4], int strategy = 0) {
  const bool col[4] = {c1, c2, c3, c4};
  int map[4], n = 0;
  for (int k = 0; k < 4; ++k) map[k] = col[k] ? n++ : -1;
  double p0[4] = {3, -1, 0, 1.0};
  for (int k = 0; k < 4; ++k) if (!col[k]) p0[k] = 0.0;
  std::vector<glio::ParamBlock> blocks{glio::ParamBlock{0, n, 0, n, false, nullptr}};
  glio::SolverOptions o;
  o.max_num_iterations = 50; o.dogleg_type = 0; o.trust_region_strategy = strategy;
  o.initial_trust_region_radius = 1e4; o.max_trust_region_radius = 1e20; o.min_lm_diagonal = 1e-6; o.max_lm_diagonal = 1e32;
  o.function_tolerance = 1e-26; o.gradient_tolerance = 1e-26; o.parameter_tolerance = 1e-26; o.jacobi_scaling = true;
  glio::TrustRegionDogleg solver(blocks, o);
  glio::EvalFn eval = [&](const double* xa, bool want_jac, double* cost, BandMat* H, double* g) -> bool {
    double x[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) if (col[k]) x[k] = xa[map[k]];
    const double x1 = x[0], x2 = x[1], x3 = x[2], x4 = x[3];
    const double f[4] = {x1 + 10.0 * x2, std::sqrt(5.0) * (x3 - x4), std::pow(x2 - 2.0 * x3, 2.0), std::sqrt(10.0) * std::pow(x1 - x4, 2.0)};
    *cost = (f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3]) / 2.0;
    if (want_jac) {
      const double Jc[4][4] = {   // Jc[column][row]
          {1.0, 0.0, 0.0, std::sqrt(10.0) * 2.0 * (x1 - x4) * (1.0 - x4)},
          {10.0, 0.0, 2.0 * (x2 - 2.0 * x3) * (1.0 - 2.0 * x3), 0.0},
          {0.0, std::sqrt(5.0), 2.0 * (x2 - 2.0 * x3) * (x2 - 2.0), 0.0},
          {0.0, -std::sqrt(5.0), 0.0, std::sqrt(10.0) * 2.0 * (x1 - x4) * (x1 - 1.0)}};
      H->reset(n, n - 1);
      for (int a = 0; a < 4; ++a) {
        if (!col[a]) continue;
        double ga = 0; for (int r = 0; r < 4; ++r) ga += Jc[a][r] * f[r];
        g[map[a]] = ga;
        for (int b = 0; b <= a; ++b) { if (!col[b]) continue; double h = 0; for (int r = 0; r < 4; ++r) h += Jc[a][r] * Jc[b][r]; H->at(map[a], map[b]) = h; }
      }
    }
    return true;
  };
  std::vector<double> xs(n);
  for (int k = 0; k < 4; ++k) if (col[k]) xs[map[k]] = p0[k];
  glio::SolverSummary S;
  solver.solve(xs.data(), eval, &S);
  for (int k = 0; k < 4; ++k) out[k] = col[k] ? xs[map[k]] : p0[k];
  bool ok = true;
  for (int k = 0; k < 4; ++k) ok = ok && std::fabs(out[k]) <= 0.001;
  return ok;
}

// ---- ceres.tgz::internal/ceres/polynomial_test.cc:144-210: FindPolynomialRoots known answers for the degrees the
// subspace dogleg uses (<= 4), against glio::detail::real_roots_deg4 (real parts; complex pairs give their real part twice,
// which is what dogleg_strategy.cc:473-507 consumes).  Tolerances and the relative/absolute rule are Ceres' ExpectClose.
static std::vector<double> add_root(const std::vector<double>& p, double r) {
  std::vector<double> q(p.size() + 1, 0.0);
  for (size_t i = 0; i < p.size(); ++i) { q[i] += p[i]; q[i + 1] -= r * p[i]; }
  return q;
}
static double rel_diff(double x, double y) { const double a = std::fabs(x - y); return (x == 0 || y == 0) ? a : a / std::max(std::fabs(x), std::fabs(y)); }
static void poly_case(const char* name, std::vector<double> roots, double eps) {
  std::vector<double> p{1.23};
  for (double r : roots) p = add_root(p, r);
  double poly5[5] = {0, 0, 0, 0, 0};
  for (size_t i = 0; i < p.size(); ++i) poly5[5 - p.size() + i] = p[i];
  std::vector<double> got;
  bool ok = glio::detail::real_roots_deg4(poly5, &got) && got.size() == roots.size();
  std::sort(got.begin(), got.end()); std::sort(roots.begin(), roots.end());
  for (size_t i = 0; ok && i < roots.size(); ++i) ok = rel_diff(got[i], roots[i]) <= eps;
  report(name, ok, got);
}

int main() {
  const double kEps = std::numeric_limits<double>::epsilon(), kLoose = 1e-5;
  // DoglegStrategyFixtureEllipse (dogleg_strategy_test.cc:60-91): J^T J = Q diag(1,2,4,8,16,32) Q^T, minimum at (1,...,1)
  const double basis[36] = {
      -0.1046920933796121, -0.7449367449921986, -0.4190744502875876, -0.4480450716142566, 0.2375351607929440, -0.0363053418882862,
      0.4064975684355914, 0.2681113508511354, -0.7463625494601520, -0.0803264850508117, -0.4463149623021321, 0.0130224954867195,
      -0.5514387729089798, 0.1026621026168657, -0.5008316122125011, 0.5738122212666414, 0.2974664724007106, 0.1296020877535158,
      0.5037835370947156, 0.2668479925183712, -0.1051754618492798, -0.0272739396578799, 0.7947481647088278, -0.1776623363955670,
      -0.4005458426625444, 0.2939330589634109, -0.0682629380550051, -0.2895448882503687, -0.0457239396341685, -0.8139899477847840,
      -0.3247764582762654, 0.4528151365941945, -0.0276683863102816, -0.6155994592510784, 0.1489240599972848, 0.5362574892189350};
  const double D[6] = {1, 2, 4, 8, 16, 32};
  std::vector<double> Je(36), re(6, 0.0);
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Je[i * 6 + j] = std::sqrt(D[i]) * basis[i * 6 + j];       // sqrtD * basis
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) re[i] -= Je[i * 6 + j] * 1.0;                               // -J * ones
  // DoglegStrategyFixtureValley (:99-119): J = diag(1,2,4,8,16,32), minimum at e_2
  std::vector<double> Jv(36, 0.0), rv(6, 0.0);
  for (int i = 0; i < 6; ++i) Jv[i * 6 + i] = D[i];
  rv[2] = -D[2];

  std::vector<double> x;
  x = first_step(Je, re, 0, 2.0); report("TrustRegionObeyedTraditional", !x.empty() && norm(x) <= 2.0 * (1.0 + 4.0 * kEps), x);      // :128-149
  x = first_step(Je, re, 1, 2.0); report("TrustRegionObeyedSubspace", !x.empty() && norm(x) <= 2.0 * (1.0 + 4.0 * kEps), x);         // :151-169
  x = first_step(Je, re, 1, 10.0);                                                                                                     // :171-194
  { bool ok = !x.empty(); for (int i = 0; ok && i < 6; ++i) ok = std::fabs(x[i] - 1.0) <= kLoose; report("CorrectGaussNewtonStep", ok, x); }
  x = first_step(Jv, rv, 1, 0.25);                                                                                                     // :231-256
  { bool ok = !x.empty(); for (int i = 0; ok && i < 6; ++i) ok = std::fabs(x[i] - (i == 2 ? 0.25 : 0.0)) <= kLoose; report("CorrectStepLocalOptimumAlongGradient", ok, x); }
  x = first_step(Jv, rv, 1, 2.0);                                                                                                      // :261-286
  { bool ok = !x.empty(); for (int i = 0; ok && i < 6; ++i) ok = std::fabs(x[i] - (i == 2 ? 1.0 : 0.0)) <= kLoose; report("CorrectStepGlobalOptimumAlongGradient", ok, x); }
  // same two Valley cases with the traditional dogleg (not in Ceres' file; the geometry gives the same answers)
  x = first_step(Jv, rv, 0, 0.25);
  { bool ok = !x.empty(); for (int i = 0; ok && i < 6; ++i) ok = std::fabs(x[i] - (i == 2 ? 0.25 : 0.0)) <= kLoose; report("ValleyTraditionalActive", ok, x); }
  x = first_step(Jv, rv, 0, 2.0);
  { bool ok = !x.empty(); for (int i = 0; ok && i < 6; ++i) ok = std::fabs(x[i] - (i == 2 ? 1.0 : 0.0)) <= kLoose; report("ValleyTraditionalInactive", ok, x); }
  // the 13 column activations of PowellsSingularFunctionUsingDogleg (:284-302); the two excluded there are excluded here
  const bool cases[13][4] = {{1, 1, 1, 0}, {1, 0, 1, 1}, {0, 1, 1, 1}, {1, 1, 0, 0}, {1, 0, 1, 0}, {0, 1, 1, 0}, {1, 0, 0, 1}, {0, 1, 0, 1}, {0, 0, 1, 1}, {1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int c = 0; c < 13; ++c) {
    double out[4];
    const bool ok = powell_case(cases[c][0], cases[c][1], cases[c][2], cases[c][3], out);
    char name[64]; snprintf(name, sizeof(name), "Powell_%d%d%d%d", (int)cases[c][0], (int)cases[c][1], (int)cases[c][2], (int)cases[c][3]);
    report(name, ok, std::vector<double>(out, out + 4));
  }
  // PowellsSingularFunctionUsingLevenbergMarquardt (trust_region_minimizer_test.cc:257-280): the 14 column activations Ceres runs
  const bool lm_cases[14][4] = {{1, 1, 1, 1}, {1, 1, 1, 0}, {1, 0, 1, 1}, {0, 1, 1, 1}, {1, 1, 0, 0}, {1, 0, 1, 0}, {0, 1, 1, 0}, {1, 0, 0, 1}, {0, 1, 0, 1}, {0, 0, 1, 1}, {1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int c = 0; c < 14; ++c) {
    double out[4];
    const bool ok = powell_case(lm_cases[c][0], lm_cases[c][1], lm_cases[c][2], lm_cases[c][3], out, 1);
    char name[64]; snprintf(name, sizeof(name), "PowellLM_%d%d%d%d", (int)lm_cases[c][0], (int)lm_cases[c][1], (int)lm_cases[c][2], (int)lm_cases[c][3]);
    report(name, ok, std::vector<double>(out, out + 4));
  }
  {  // LevenbergMarquardtStrategy radius rule (levenberg_marquardt_strategy_test.cc:80-112, AcceptRejectStepRadiusScaling) driven
     // through the minimizer: a 1-D problem whose evaluator returns scripted costs so that the step qualities are exactly
     // rejected, rejected, 1, 1, 0.25, 1, 1, 1 with initial radius 2 and max radius 20 -> radii 1, 0.25, 0.75, 2.25, 2, 6, 18, 20.
     // With H = 1, g = -1 (r = x - 1 at x = 0 ... kept constant), the LM step is s = 1 / (1 + 1/radius) and the model decrease
     // is s - s*s/2; the scripted candidate cost = cost - quality * model decrease.
    std::vector<glio::ParamBlock> blocks{glio::ParamBlock{0, 1, 0, 1, false, nullptr}};
    glio::SolverOptions o; o.trust_region_strategy = 1; o.max_num_iterations = 8; o.initial_trust_region_radius = 2.0; o.max_trust_region_radius = 20.0;
    o.min_lm_diagonal = 1.0; o.max_lm_diagonal = 1.0; o.jacobi_scaling = false; o.function_tolerance = 0; o.gradient_tolerance = 0; o.parameter_tolerance = 0;
    o.fuse_candidate_jacobian = false; o.min_relative_decrease = 1e-3;
    const double quality[8] = {-0.5, -1.0, 1.0, 1.0, 0.25, 1.0, 1.0, 1.0};   // (a quality of exactly 0 would trip the function tolerance test first)
    const double radii[9] = {2.0, 1.0, 0.25, 0.75, 2.25, 2.0, 6.0, 18.0, 20.0};
    int k = 0; double cur_cost = 100.0;
    glio::EvalFn eval = [&](const double* x, bool want_jac, double* cost, BandMat* H, double* g) -> bool {
      if (want_jac) { *cost = cur_cost; H->reset(1, 0); H->at(0, 0) = 1.0; g[0] = -1.0; return true; }
      const double s = 1.0 / (1.0 + 1.0 / radii[k]), dec = s - 0.5 * s * s;
      const double cand = cur_cost - quality[k] * dec;
      if (quality[k] > 1e-3) cur_cost = cand;
      ++k; *cost = cand; return true;
    };
    std::vector<double> x0(1, 0.0); glio::SolverSummary S;
    glio::TrustRegionDogleg solver(blocks, o);
    solver.solve(x0.data(), eval, &S);
    bool ok = S.iterations.size() == 9; std::vector<double> got;
    for (size_t i = 0; i < S.iterations.size(); ++i) { got.push_back(S.iterations[i].trust_region_radius); if (i < 9) ok = ok && std::fabs(S.iterations[i].trust_region_radius - radii[i]) <= 1e-12 * radii[i]; }
    report("LM_AcceptRejectStepRadiusScaling", ok, got);
  }
  {  // CorrectDiagonalToLinearSolver (levenberg_marquardt_strategy_test.cc:114-160): J = [[0 1 100],[0 1 0]], min/max_lm_diagonal 1e-2 / 1e2,
     // radius 2 -> D^2 = {1e-2, 2, 1e2} / 2 on the diagonal of the regularised normal equations; the first LM step must equal
     // -(J^T J + D^2)^-1 J^T r computed independently.
    std::vector<glio::ParamBlock> blocks{glio::ParamBlock{0, 3, 0, 3, false, nullptr}};
    glio::SolverOptions o; o.trust_region_strategy = 1; o.max_num_iterations = 1; o.initial_trust_region_radius = 2.0; o.max_trust_region_radius = 20.0;
    o.min_lm_diagonal = 1e-2; o.max_lm_diagonal = 1e2; o.jacobi_scaling = false; o.function_tolerance = 0; o.gradient_tolerance = 0; o.parameter_tolerance = 0;
    const double Jm[2][3] = {{0.0, 1.0, 100.0}, {0.0, 1.0, 0.0}}, r0[2] = {1.0, 1.0};
    glio::EvalFn eval = [&](const double* x, bool want_jac, double* cost, BandMat* H, double* g) -> bool {
      double r[2]; for (int i = 0; i < 2; ++i) { r[i] = r0[i]; for (int j = 0; j < 3; ++j) r[i] += Jm[i][j] * x[j]; }
      *cost = 0.5 * (r[0] * r[0] + r[1] * r[1]);
      if (want_jac) { H->reset(3, 2); for (int a = 0; a < 3; ++a) { g[a] = Jm[0][a] * r[0] + Jm[1][a] * r[1]; for (int b = 0; b <= a; ++b) H->at(a, b) = Jm[0][a] * Jm[0][b] + Jm[1][a] * Jm[1][b]; } }
      return true;
    };
    std::vector<double> x0(3, 0.0); glio::SolverSummary S;
    glio::TrustRegionDogleg solver(blocks, o);
    solver.solve(x0.data(), eval, &S);
    // independent 3x3 solve: A = J^T J + diag(1e-2, 2, 1e2)/2 ; x0 column decouples (A00 = 0.005, g0 = 0)
    const double A11 = 2.0 + 1.0, A12 = 100.0, A22 = 10000.0 + 50.0, g1 = 2.0, g2 = 100.0;
    const double det = A11 * A22 - A12 * A12; const double e1 = -(A22 * g1 - A12 * g2) / det, e2 = -(-A12 * g1 + A11 * g2) / det;
    bool ok = S.steps.size() >= 3 && std::fabs(S.steps[0]) <= 1e-15 && std::fabs(S.steps[1] - e1) <= 1e-12 * std::fabs(e1) + 1e-15 && std::fabs(S.steps[2] - e2) <= 1e-12 * std::fabs(e2);
    report("LM_CorrectDiagonalToLinearSolver", ok, S.steps);
  }
  {  // Solver::Options::max_solver_time_in_seconds (trust_region_minimizer.cc:327-329, :605-619): a zero budget stops the
     // solve before the first step with NO_CONVERGENCE, as the reference's front end relies on (LidarOdometry.cpp:527)
    std::vector<glio::ParamBlock> blocks{glio::ParamBlock{0, 6, 0, 6, false, nullptr}};
    glio::SolverOptions o; o.max_solver_time_in_seconds = 0.0;
    glio::TrustRegionDogleg solver(blocks, o);
    glio::EvalFn eval = [&](const double* x, bool want_jac, double* cost, BandMat* H, double* g) -> bool {
      double c = 0; for (int i = 0; i < 6; ++i) c += 0.5 * (x[i] - 1.0) * (x[i] - 1.0);
      *cost = c;
      if (want_jac) { H->reset(6, 5); for (int i = 0; i < 6; ++i) { H->at(i, i) = 1.0; g[i] = x[i] - 1.0; } }
      return true;
    };
    std::vector<double> x0(6, 0.0); glio::SolverSummary S;
    solver.solve(x0.data(), eval, &S);
    report("MaxSolverTimeReached", S.termination == glio::TERM_NO_CONVERGENCE && S.message.find("Maximum solver time reached") == 0 && S.steps.empty() && x0[0] == 0.0, x0);
  }
  poly_case("Poly_LinearPositive", {42.42}, 1e-13); poly_case("Poly_LinearNegative", {-42.42}, 1e-13);
  poly_case("Poly_QuadraticPositive", {1.0, 42.42}, 1e-13); poly_case("Poly_QuadraticOneNegative", {-42.42, 1.0}, 1e-13);
  poly_case("Poly_QuadraticTwoNegative", {-42.42, -1.0}, 1e-13); poly_case("Poly_QuadraticClose", {42.42, 42.43}, 1e-9);
  poly_case("Poly_Quartic", {1.23e-4, 1.23e-1, 1.23e+2, 1.23e+5}, 1e-13); poly_case("Poly_QuarticTwoClusters", {1.23e-1, 2.46e-1, 1.23e+5, 2.46e+5}, 1e-9);
  poly_case("Poly_QuarticTwoZeroRoots", {-42.42, 0.0, 0.0, 42.42}, 2e-9); poly_case("Poly_QuarticMonomial", {0.0, 0.0, 0.0, 0.0}, 1e-13);
  {  // QuadraticPolynomialWithComplexRootsWorks (:174-190): 42.42 +- 4.2i -> real parts 42.42, 42.42
    const double poly5[5] = {0, 0, 1.23, -2 * 42.42 * 1.23, (42.42 * 42.42 + 4.2 * 4.2) * 1.23};
    std::vector<double> got; bool ok = glio::detail::real_roots_deg4(poly5, &got) && got.size() == 2;
    for (size_t i = 0; ok && i < 2; ++i) ok = rel_diff(got[i], 42.42) <= 1e-13;
    report("Poly_QuadraticComplexPair", ok, got);
  }
  return 0;
}
