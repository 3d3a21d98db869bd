// band_chol_test.cpp — the band Cholesky + solve of the host minimizer (glio::detail::cholesky_solve: portable path,
// AVX2/FMA panel path, run-time dispatcher) against a dense reference on random SPD band matrices, including the window
// shape (n = 300, hb = 29), the batch shapes (n = 2400, hb = 41; wide band), tiny and degenerate sizes, and a non-positive-definite
// input; and the AVX2 band matrix-vector product (glio::detail::symv_avx2) against the dense product on the same matrices.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../../glio_b200/csrc/solver.h"

using glio::BandMat;

static bool dense_solve(const BandMat& A, const std::vector<double>& b, std::vector<double>& x) {
  const int n = A.n; std::vector<double> M((size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) M[(size_t)i * n + j] = A.sym(i, j);
  for (int j = 0; j < n; ++j) {
    double d = M[(size_t)j * n + j]; for (int k = 0; k < j; ++k) d -= M[(size_t)j * n + k] * M[(size_t)j * n + k];
    if (!(d > 0)) return false;
    const double l = std::sqrt(d); M[(size_t)j * n + j] = l;
    for (int i = j + 1; i < n; ++i) { double s = M[(size_t)i * n + j]; for (int k = 0; k < j; ++k) s -= M[(size_t)i * n + k] * M[(size_t)j * n + k]; M[(size_t)i * n + j] = s / l; }
  }
  x = b;
  for (int i = 0; i < n; ++i) { double s = x[i]; for (int k = 0; k < i; ++k) s -= M[(size_t)i * n + k] * x[k]; x[i] = s / M[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < n; ++k) s -= M[(size_t)k * n + i] * x[k]; x[i] = s / M[(size_t)i * n + i]; }
  return true;
}

int main() {
  std::mt19937_64 rng(11); std::normal_distribution<double> N(0, 1);
  int bad = 0, cases = 0;
  const int shapes[][2] = {{300, 29}, {300, 29}, {2400, 41}, {1200, 83}, {90, 89}, {17, 3}, {9, 8}, {8, 4}, {7, 6}, {5, 0}, {3, 2}, {1, 0}, {64, 1}, {33, 31}, {150, 14}, {6, 5}};
  for (auto& sh : shapes) {
    const int n = sh[0], hb = std::min(sh[1], std::max(n - 1, 0));
    BandMat L0; L0.reset(n, hb);
    for (int i = 0; i < n; ++i) for (int j = std::max(0, i - hb); j <= i; ++j) L0.at(i, j) = (i == j) ? 1.5 + std::fabs(N(rng)) : 0.4 / std::sqrt(1.0 + hb / 8.0) * N(rng);   // keeps the wide bands well conditioned
    BandMat A; A.reset(n, hb);
    for (int i = 0; i < n; ++i) for (int j = std::max(0, i - hb); j <= i; ++j) { double s = 0; for (int k = std::max(0, i - hb); k <= j; ++k) if (k >= j - hb) s += L0.at(i, k) * L0.at(j, k); A.at(i, j) = s; }
    std::vector<double> b(n), xr, xs(n), xd(n), xa(n);
    for (double& v : b) v = N(rng);
    const bool okr = dense_solve(A, b, xr);
    if (!okr) { printf("shape %d %d FAIL dense reference rejected the matrix\n", n, hb); ++bad; ++cases; continue; }
    BandMat A1 = A, A2 = A, A3 = A;
    const bool oks = glio::detail::cholesky_solve_scalar(A1, b.data(), xs.data());
    const bool okd = glio::detail::cholesky_solve(A2, b.data(), xd.data());
    bool oka = true; bool have_avx2 = false;
#if defined(__x86_64__)
    have_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
    if (have_avx2) oka = glio::detail::cholesky_solve_avx2(A3, b.data(), xa.data()); else xa = xs;
#else
    xa = xs;
#endif
    // y = A x: AVX2 band product against the dense one
    double ey = 0, ny = 1e-300;
#if defined(__x86_64__)
    if (have_avx2) {
      std::vector<double> xv(n), yr(n, 0.0), ya(n, -7.0);
      for (double& v : xv) v = N(rng);
      for (int i = 0; i < n; ++i) { double s = 0; for (int j = std::max(0, i - hb); j <= std::min(n - 1, i + hb); ++j) s += A.sym(i, j) * xv[j]; yr[i] = s; }
      glio::detail::symv_avx2(A, xv.data(), ya.data());
      for (int i = 0; i < n; ++i) { ny = std::max(ny, std::fabs(yr[i])); ey = std::max(ey, std::fabs(ya[i] - yr[i])); }
    }
#endif
    double nx = 1e-300, es = 0, ed = 0, ea = 0;
    for (int i = 0; i < n; ++i) { nx = std::max(nx, std::fabs(xr[i])); es = std::max(es, std::fabs(xs[i] - xr[i])); ed = std::max(ed, std::fabs(xd[i] - xr[i])); ea = std::max(ea, std::fabs(xa[i] - xr[i])); }
    const bool ok = okr && oks && okd && oka && es / nx < 1e-9 && ed / nx < 1e-9 && ea / nx < 1e-9 && ey / ny < 1e-12;
    printf("shape %d %d %s scalar %.2e dispatch %.2e avx2 %.2e symv %.2e (avx2 available %d)\n", n, hb, ok ? "ok" : "FAIL", es / nx, ed / nx, ea / nx, ey / ny, (int)have_avx2);
    bad += !ok; ++cases;
  }
  {  // not positive definite: every path must report failure
    BandMat A; A.reset(40, 9);
    for (int i = 0; i < 40; ++i) { A.at(i, i) = 1.0; if (i > 0) A.at(i, i - 1) = 0.9; if (i > 1) A.at(i, i - 2) = 0.9; }
    A.at(20, 20) = -1.0;
    std::vector<double> b(40, 1.0), x(40);
    BandMat A1 = A, A2 = A, A3 = A;
    bool f = !glio::detail::cholesky_solve_scalar(A1, b.data(), x.data()) && !glio::detail::cholesky_solve(A2, b.data(), x.data());
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) f = f && !glio::detail::cholesky_solve_avx2(A3, b.data(), x.data());
#endif
    printf("indefinite %s\n", f ? "ok" : "FAIL"); bad += !f; ++cases;
  }
  printf("cases %d bad %d\n", cases, bad);
  return bad ? 1 : 0;
}
