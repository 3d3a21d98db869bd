// shim_known_answers.cpp — Ceres' own unit-test expectations (the reference vendors Ceres 2.0.0 as
// support_files/ceres-solver.tar.gz) run against the Ceres-API shim of the PRODUCT (glio_b200/shim/ceres/):
//   loss_function_test.cc:47-105        TrivialLoss / HuberLoss / CauchyLoss: rho', rho'' vs symmetric differences
//   corrector_test.cc:58-147            Corrector scalar cases (shim_internal::Corr)
//   local_parameterization_test.cc:232-352  QuaternionParameterization Plus / ComputeJacobian: zero, near-zero, away from zero
//   rotation_test.cc (quaternion product / rotate point identities used by the factors)
//   problem_test.cc:1064-1249, :1363-1394  Problem::Evaluate known answers (cost 7607, residuals, gradient; constant block)
//   autodiff_cost_function_test.cc:42-143   AutoDiffCostFunction residuals / Jacobians (bilinear, ten parameter blocks)
//   jet_test.cc style checks: every Jet function the reference's functors use, derivative vs symmetric differences
// Prints "name ok|FAIL value" lines; tests/test_shim_known_answers.py asserts on them.
#include <cmath>
#include <cstdio>
#include <functional>

#include "ceres/ceres.h"
#include "ceres/rotation.h"

static int n_fail = 0;
static void report(const char* name, bool ok, double v = 0) { printf("%s %s %.17g\n", name, ok ? "ok" : "FAIL", v); if (!ok) ++n_fail; }

// loss_function_test.cc:47-70
static bool loss_valid(const ceres::LossFunction& loss, double s) {
  double rho[3], fwd[3], bwd[3];
  const double kH = 1e-4;
  loss.Evaluate(s, rho); loss.Evaluate(s + kH, fwd); loss.Evaluate(s - kH, bwd);
  const double fd1 = (fwd[0] - bwd[0]) / (2 * kH), fd2 = (fwd[0] - 2 * rho[0] + bwd[0]) / (kH * kH);
  return std::fabs(fd1 - rho[1]) <= 1e-6 && std::fabs(fd2 - rho[2]) <= 1e-6;
}

// local_parameterization_test.cc:232-290 (QuaternionParameterizationTestHelper), with the reference Jacobian taken by
// symmetric differences of Plus at delta = 0 instead of Ceres' autodiff (same quantity)
static bool quat_helper(const double* x, const double* delta, const double* x_plus_delta_ref) {
  const double kTol = 1e-14;
  ceres::QuaternionParameterization p;
  double xpd[4] = {0, 0, 0, 0};
  p.Plus(x, delta, xpd);
  bool ok = true;
  double nrm = 0;
  for (int i = 0; i < 4; ++i) { ok = ok && std::fabs(xpd[i] - x_plus_delta_ref[i]) <= kTol; nrm += xpd[i] * xpd[i]; }
  ok = ok && std::fabs(std::sqrt(nrm) - 1.0) <= kTol;
  double J[12]; p.ComputeJacobian(x, J);
  for (int c = 0; c < 3; ++c) {
    const double h = 1e-6;
    double dp[3] = {0, 0, 0}, dm[3] = {0, 0, 0}, a[4], b[4];
    dp[c] = h; dm[c] = -h;
    p.Plus(x, dp, a); p.Plus(x, dm, b);
    for (int r = 0; r < 4; ++r) ok = ok && std::isfinite(J[r * 3 + c]) && std::fabs(J[r * 3 + c] - (a[r] - b[r]) / (2 * h)) <= 1e-9;
  }
  return ok;
}

template <typename F> static bool jet_fn(F f, double x, std::function<double(double)> g) {
  typedef ceres::Jet<double, 1> J;
  J a; a.a = x; a.v[0] = 1.0;
  const J r = f(a);
  const double h = 1e-6 * std::max(1.0, std::fabs(x));
  const double fd = (g(x + h) - g(x - h)) / (2 * h);
  return std::fabs(r.a - g(x)) <= 1e-14 * std::max(1.0, std::fabs(g(x))) && std::fabs(r.v[0] - fd) <= 1e-7 * std::max(1.0, std::fabs(fd));
}

// problem_test.cc:1064-1099: residual_i = i - sum_j (j+1) * p_j[i]^2, jacobian_j = diag(-2 (j+1) p_j)
template <int kNumResiduals, int kNumParameterBlocks>
class QuadraticCostFunction : public ceres::CostFunction {
 public:
  QuadraticCostFunction() {
    set_num_residuals(kNumResiduals);
    for (int i = 0; i < kNumParameterBlocks; ++i) mutable_parameter_block_sizes()->push_back(kNumResiduals);
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const final {
    for (int i = 0; i < kNumResiduals; ++i) {
      residuals[i] = i;
      for (int j = 0; j < kNumParameterBlocks; ++j) residuals[i] -= (j + 1.0) * parameters[j][i] * parameters[j][i];
    }
    if (jacobians == NULL) return true;
    for (int j = 0; j < kNumParameterBlocks; ++j)
      if (jacobians[j] != NULL)
        for (int r = 0; r < kNumResiduals; ++r) for (int c = 0; c < kNumResiduals; ++c) jacobians[j][r * kNumResiduals + c] = r == c ? -2.0 * (j + 1.0) * parameters[j][r] : 0.0;
    return true;
  }
};

// problem_test.cc:1110-1146 fixture + :1221-1249 / :1363-1394 expectations (cost, residuals, gradient)
static void problem_evaluate_tests() {
  for (int constant_y = 0; constant_y < 2; ++constant_y) {
    double parameters[6]; for (int i = 0; i < 6; ++i) parameters[i] = i + 1.0;
    ceres::Problem problem;
    ceres::CostFunction* cost_function = new QuadraticCostFunction<2, 2>;
    problem.AddResidualBlock(cost_function, NULL, parameters, parameters + 2);        // f(x, y)
    problem.AddResidualBlock(cost_function, NULL, parameters + 2, parameters + 4);    // g(y, z)
    problem.AddResidualBlock(cost_function, NULL, parameters + 4, parameters);        // h(z, x)
    if (constant_y) problem.SetParameterBlockConstant(parameters + 2);
    const double exp_res[6] = {-19.0, -35.0, -59.0, -87.0, -27.0, -43.0};
    const double exp_grad[2][6] = {{146.0, 484.0, 582.0, 1256.0, 1450.0, 2604.0}, {146.0, 484.0, 0.0, 0.0, 1450.0, 2604.0}};
    double cost = 0; std::vector<double> residuals, gradient;
    bool ok = problem.Evaluate(ceres::Problem::EvaluateOptions(), &cost, &residuals, &gradient, NULL);
    ok = ok && cost == 7607.0 && residuals.size() == 6 && gradient.size() == 6;
    for (int i = 0; ok && i < 6; ++i) ok = residuals[i] == exp_res[i] && gradient[i] == exp_grad[constant_y][i];
    report(constant_y ? "ProblemEvaluate_ConstantParameterBlock" : "ProblemEvaluate_MultipleParameterAndResidualBlocks", ok, cost);
  }
}

// autodiff_cost_function_test.cc:42-143
struct BinaryScalarCost {
  explicit BinaryScalarCost(double a) : a_(a) {}
  template <typename T> bool operator()(const T* const x, const T* const y, T* cost) const { cost[0] = x[0] * y[0] + x[1] * y[1] - T(a_); return true; }
  double a_;
};
struct TenParameterCost {
  template <typename T> bool operator()(const T* const x0, const T* const x1, const T* const x2, const T* const x3, const T* const x4, const T* const x5,
                                        const T* const x6, const T* const x7, const T* const x8, const T* const x9, T* cost) const {
    cost[0] = *x0 + *x1 + *x2 + *x3 + *x4 + *x5 + *x6 + *x7 + *x8 + *x9; return true;
  }
};
static void autodiff_tests() {
  {
    ceres::AutoDiffCostFunction<BinaryScalarCost, 1, 2, 2> cf(new BinaryScalarCost(1.0));
    double x[2] = {1, 2}, y[2] = {3, 4}; const double* params[2] = {x, y};
    double jx[2], jy[2]; double* jac[2] = {jx, jy}; double r = 0, r2 = 0;
    bool ok = cf.Evaluate(params, &r, nullptr) && r == 10.0 && cf.Evaluate(params, &r2, jac) && r2 == 10.0 && jx[0] == 3 && jx[1] == 4 && jy[0] == 1 && jy[1] == 2;
    report("AutoDiff_BilinearDifferentiationTest", ok, r);
  }
  {
    ceres::AutoDiffCostFunction<TenParameterCost, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1> cf(new TenParameterCost);
    double p[10], j[10]; const double* params[10]; double* jac[10];
    for (int i = 0; i < 10; ++i) { p[i] = i; params[i] = &p[i]; jac[i] = &j[i]; }
    double r = 0; bool ok = cf.Evaluate(params, &r, nullptr) && r == 45.0 && cf.Evaluate(params, &r, jac) && r == 45.0;
    for (int i = 0; ok && i < 10; ++i) ok = j[i] == 1.0;
    report("AutoDiff_ManyParameterAutodiffInstantiates", ok, r);
  }
}

int main() {
  // ---- losses
  for (double s : {0.357, 1.792}) {
    report("TrivialLoss", loss_valid(ceres::TrivialLoss(), s), s);
    for (double a : {0.7, 1.3}) { report("HuberLoss", loss_valid(ceres::HuberLoss(a), s), a); report("CauchyLoss", loss_valid(ceres::CauchyLoss(a), s), a); }
  }
  // ---- corrector (corrector_test.cc:58-147): scalar residual sqrt(3), jacobian 10
  {
    using ceres::shim_internal::Corr;
    struct C { const char* name; double res, rho[3]; } cs[3] = {{"ScalarCorrection", std::sqrt(3.0), {3.0, 0.1, -0.01}}, {"ScalarCorrectionZeroResidual", 0.0, {0.0, 0.1, -0.01}},
                                                                {"ScalarCorrectionAlphaClamped", std::sqrt(3.0), {3.0, 0.1, -0.1}}};
    for (auto& c : cs) {
      double r = c.res, J = 10.0;
      Corr corr(c.res * c.res, c.rho);
      corr.jac(1, 1, &r, &J); r *= corr.residual_scaling;
      const double kAlpha = 0.0;     // rho'' < 0 or zero residual -> alpha = 0
      report(c.name, std::fabs(r - c.res * std::sqrt(c.rho[1]) / (1 - kAlpha)) <= 1e-6 && std::fabs(J - std::sqrt(c.rho[1]) * (1 - kAlpha) * 10.0) <= 1e-6, J);
    }
    // MultidimensionalGaussNewtonApproximation (corrector_test.cc:152-214), one fixed instance: the corrected J^T J and
    // J^T r must equal the robustified Gauss-Newton terms rho' J^T J + 2 rho'' J^T r r^T J and rho' J^T r
    const double r0[3] = {0.3, -1.1, 0.7}, J0[6] = {1.0, 2.0, -0.5, 0.4, 0.25, -3.0};   // 3 x 2 row-major
    double sq = 0; for (double v : r0) sq += v * v;
    const double rho[3] = {sq, 0.6, 0.08};
    double r[3] = {r0[0], r0[1], r0[2]}, J[6]; for (int i = 0; i < 6; ++i) J[i] = J0[i];
    Corr corr(sq, rho);
    corr.jac(3, 2, r, J); for (double& v : r) v *= corr.residual_scaling;
    bool ok = true;
    for (int a = 0; a < 2; ++a) {
      double g = 0, ge = 0; for (int k = 0; k < 3; ++k) { g += J[k * 2 + a] * r[k]; ge += rho[1] * J0[k * 2 + a] * r0[k]; }
      ok = ok && std::fabs(g - ge) <= 1e-10;
      for (int b = 0; b < 2; ++b) {
        double h = 0, he = 0, ja = 0, jb = 0;
        for (int k = 0; k < 3; ++k) { h += J[k * 2 + a] * J[k * 2 + b]; he += rho[1] * J0[k * 2 + a] * J0[k * 2 + b]; ja += J0[k * 2 + a] * r0[k]; jb += J0[k * 2 + b] * r0[k]; }
        he += 2 * rho[2] * ja * jb;
        ok = ok && std::fabs(h - he) <= 1e-10;
      }
    }
    report("MultidimensionalGaussNewtonApproximation", ok);
  }
  // ---- QuaternionParameterization (local_parameterization_test.cc:305-352)
  {
    double x[4] = {0.5, 0.5, 0.5, 0.5}, d[3] = {0, 0, 0}, qd[4] = {1, 0, 0, 0}, ref[4];
    ceres::QuaternionProduct(qd, x, ref);
    report("QuaternionZeroTest", quat_helper(x, d, ref));
    double y[4] = {0.52, 0.25, 0.15, 0.45}; double n = 0; for (double v : y) n += v * v; n = std::sqrt(n); for (double& v : y) v /= n;
    double d2[3] = {0.24e-14, 0.15e-14, 0.10e-14}, qd2[4] = {1.0, d2[0], d2[1], d2[2]};
    ceres::QuaternionProduct(qd2, y, ref);
    report("QuaternionNearZeroTest", quat_helper(y, d2, ref));
    double d3[3] = {0.24, 0.15, 0.10}; const double dn = std::sqrt(d3[0] * d3[0] + d3[1] * d3[1] + d3[2] * d3[2]);
    double qd3[4] = {std::cos(dn), std::sin(dn) / dn * d3[0], std::sin(dn) / dn * d3[1], std::sin(dn) / dn * d3[2]};
    ceres::QuaternionProduct(qd3, y, ref);
    report("QuaternionAwayFromZeroTest", quat_helper(y, d3, ref));
  }
  // ---- rotation.h identities (rotation_test.cc: QuaternionRotatePoint gives the rotation-matrix answer; products compose)
  {
    double q[4] = {0.3, -0.5, 0.7, 0.1}, p[3] = {1.5, -2.0, 0.25}, r1[3];
    double n = 0; for (double v : q) n += v * v; n = std::sqrt(n);
    double u[4] = {q[0] / n, q[1] / n, q[2] / n, q[3] / n};
    ceres::QuaternionRotatePoint(q, p, r1);                   // normalises internally
    const double w = u[0], x = u[1], y = u[2], z = u[3];
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                         2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
    bool ok = true;
    for (int i = 0; i < 3; ++i) ok = ok && std::fabs(r1[i] - (R[3 * i] * p[0] + R[3 * i + 1] * p[1] + R[3 * i + 2] * p[2])) <= 1e-14;
    report("QuaternionRotatePointGivesSameAnswerAsRotationMatrix", ok);
    double v[4] = {0.9, 0.1, -0.3, 0.2}; double m = 0; for (double a : v) m += a * a; m = std::sqrt(m); for (double& a : v) a /= m;
    double uv[4], t1[3], t2[3], t3[3];
    ceres::QuaternionProduct(u, v, uv); ceres::UnitQuaternionRotatePoint(v, p, t1); ceres::UnitQuaternionRotatePoint(u, t1, t2); ceres::UnitQuaternionRotatePoint(uv, p, t3);
    ok = true; for (int i = 0; i < 3; ++i) ok = ok && std::fabs(t2[i] - t3[i]) <= 1e-14;
    report("QuaternionProductComposesRotations", ok);
  }
  // ---- Jet: value and derivative of every function the functors may call (jet_test.cc compares with numeric differentiation)
  {
    typedef ceres::Jet<double, 1> J;
    report("Jet_sqrt", jet_fn([](J a) { return sqrt(a); }, 2.3, [](double x) { return std::sqrt(x); }));
    report("Jet_exp", jet_fn([](J a) { return exp(a); }, 0.7, [](double x) { return std::exp(x); }));
    report("Jet_log", jet_fn([](J a) { return log(a); }, 1.9, [](double x) { return std::log(x); }));
    report("Jet_sin", jet_fn([](J a) { return sin(a); }, 0.4, [](double x) { return std::sin(x); }));
    report("Jet_cos", jet_fn([](J a) { return cos(a); }, 0.4, [](double x) { return std::cos(x); }));
    report("Jet_tan", jet_fn([](J a) { return tan(a); }, 0.4, [](double x) { return std::tan(x); }));
    report("Jet_asin", jet_fn([](J a) { return asin(a); }, 0.3, [](double x) { return std::asin(x); }));
    report("Jet_acos", jet_fn([](J a) { return acos(a); }, 0.3, [](double x) { return std::acos(x); }));
    report("Jet_atan", jet_fn([](J a) { return atan(a); }, 1.3, [](double x) { return std::atan(x); }));
    report("Jet_abs", jet_fn([](J a) { return abs(a); }, -1.3, [](double x) { return std::fabs(x); }));
    report("Jet_pow_c", jet_fn([](J a) { return pow(a, 2.5); }, 1.7, [](double x) { return std::pow(x, 2.5); }));
    report("Jet_c_pow", jet_fn([](J a) { return pow(2.5, a); }, 1.7, [](double x) { return std::pow(2.5, x); }));
    report("Jet_pow_jj", jet_fn([](J a) { return pow(a, a); }, 1.7, [](double x) { return std::pow(x, x); }));
    report("Jet_atan2", jet_fn([](J a) { return atan2(a, J(0.8) + a * 0.5); }, 0.6, [](double x) { return std::atan2(x, 0.8 + 0.5 * x); }));
    report("Jet_div", jet_fn([](J a) { return (J(1.0) + a * a) / (a + 2.0) - 3.0 / a; }, 0.9, [](double x) { return (1 + x * x) / (x + 2) - 3.0 / x; }));
  }
  problem_evaluate_tests();
  autodiff_tests();
  return n_fail ? 1 : 0;
}
