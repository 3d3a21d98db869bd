// marginalize.cpp — K3: MarginalizationInfo::Marginalize (GLIO/src/MarginalizationFactor.cpp:128-202), host C++.
//
// The LiDAR part of the dense normal equations A, b comes from the device (K2 with jac_kind = 1: the ambient x,y,z
// quaternion columns of MarginalizationFactor.cpp:9-17, one 6x6 block + 6-vector per keyframe); the host factors (IMU
// 0->1, the previous prior) add theirs through a callback; this file does what is left:
//   Amm <- 0.5 (Amm + Amm^T); Amm^-1 = V diag(lambda > eps ? 1/lambda : 0) V^T          (:176-182)
//   A' = Arr - Arm Amm^-1 Amr ;  b' = brr - Arm Amm^-1 bmm                              (:184-190)
//   A' = V2 diag(S) V2^T, S = lambda > eps ? lambda : 0                                  (:192-196)
//   linearized_jacobians = diag(sqrt S) V2^T ; linearized_residuals = diag(sqrt(1/S)) V2^T b'   (:198-201)
// Sizes at W = 20: m = 15, n = 123 — the Schur product is 123 x 15 x 123 (0.23 M multiply-adds): far too small for tensor
// cores to matter, and no GEMM here shows up in any profile; it stays on the host.  What would be expensive is the
// eigen-decomposition of a dense 123 x 123 matrix (~10 n^3 flops).  But the window problem's A' is block diagonal — LiDAR
// factors are unary, the only coupling removed with KF0 is the IMU factor to KF1 — so the matrix is first split into the
// connected components of its non-zero pattern and every component (6 x 6, one 15 x 15) is decomposed on its own:
// identical eigenpairs, microseconds instead of milliseconds.  A genuinely dense A' falls through to one full-size
// decomposition (Householder tridiagonalisation + implicit QL).
// Eigenvectors are defined up to sign / order within equal eigenvalues, so J and r are too; J^T J = A' (restricted to
// eigenvalues > eps) and J^T r are the invariants (tests/test_marginalize.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/glio_b200.h"
#include "marg.h"

namespace {

// Symmetric eigenproblem, EISPACK tred2 + tql2 (Householder tridiagonalisation, implicit-shift QL); V row-major n x n:
// in = the matrix, out = eigenvectors in COLUMNS; d = eigenvalues ascending.
void eigh_ql(std::vector<double>& Vm, int n, std::vector<double>& d) {
  d.assign(n, 0.0);
  if (n == 0) return;
  std::vector<double> e(n, 0.0);
  double* V = Vm.data();
  auto at = [&](int i, int j) -> double& { return V[(size_t)i * n + j]; };
  for (int j = 0; j < n; ++j) d[j] = at(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = at(i - 1, j); at(i, j) = 0.0; at(j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1], g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g; h -= f * g; d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j]; at(j, i) = f; g = e[j] + at(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) { g += at(k, j) * d[k]; e[k] += at(k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; ++k) at(k, j) -= (f * e[k] + g * d[k]);
        d[j] = at(i - 1, j); at(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    at(n - 1, i) = at(i, i); at(i, i) = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = at(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += at(k, i + 1) * at(k, j);
        for (int k = 0; k <= i; ++k) at(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) at(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) { d[j] = at(n - 1, j); at(n - 1, j) = 0.0; }
  at(n - 1, n - 1) = 1.0; e[0] = 0.0;
  // tql2 on the transposed eigenvector array (contiguous column updates)
  std::vector<double> Vt((size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Vt[(size_t)j * n + i] = at(i, j);
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = std::pow(2.0, -52.0);
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; ++m; }
    if (m > l) {
      int iter = 0;
      do {
        ++iter;
        double g = d[l], p = (d[l + 1] - g) / (2.0 * e[l]), r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r); d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
        const double el1 = e[l + 1];
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i]; h = c * p; r = std::hypot(p, e[i]);
          e[i + 1] = s * r; s = e[i] / r; c = p / r; p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          double* __restrict v1 = &Vt[(size_t)(i + 1) * n];
          double* __restrict v0 = &Vt[(size_t)i * n];
          for (int k = 0; k < n; ++k) { const double hv = v1[k]; v1[k] = s * v0[k] + c * hv; v0[k] = c * v0[k] - s * hv; }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1; e[l] = s * p; d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] += f; e[l] = 0.0;
  }
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int x, int y) { return d[x] < d[y]; });
  std::vector<double> ds(n);
  for (int j = 0; j < n; ++j) { ds[j] = d[idx[j]]; for (int k = 0; k < n; ++k) at(k, j) = Vt[(size_t)idx[j] * n + k]; }
  d.swap(ds);
}

// Eigen-decomposition of a symmetric matrix by connected components of its non-zero pattern: a (n x n row-major) ->
// w (n eigenvalues), V (eigenvectors in columns, zero outside the component's rows).  Block-diagonal input = same
// eigenpairs as the full decomposition at a fraction of the cost.
void eigh_components(const std::vector<double>& a, int n, std::vector<double>& w, std::vector<double>& V) {
  std::vector<int> comp(n);
  std::iota(comp.begin(), comp.end(), 0);
  auto find = [&](int x) { while (comp[x] != x) { comp[x] = comp[comp[x]]; x = comp[x]; } return x; };
  for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) if (a[(size_t)i * n + j] != 0.0 || a[(size_t)j * n + i] != 0.0) { const int ri = find(i), rj = find(j); if (ri != rj) comp[ri] = rj; }
  w.assign(n, 0.0); V.assign((size_t)n * n, 0.0);
  std::vector<char> seen(n, 0);
  int col = 0;
  for (int r = 0; r < n; ++r) {
    const int root = find(r);
    if (seen[root]) continue;
    seen[root] = 1;
    std::vector<int> mem;
    for (int i = 0; i < n; ++i) if (find(i) == root) mem.push_back(i);
    const int k = (int)mem.size();
    std::vector<double> sub((size_t)k * k), d;
    for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) sub[(size_t)i * k + j] = a[(size_t)mem[i] * n + mem[j]];
    eigh_ql(sub, k, d);
    for (int j = 0; j < k; ++j, ++col) { w[col] = d[j]; for (int i = 0; i < k; ++i) V[(size_t)mem[i] * n + col] = sub[(size_t)i * k + j]; }
  }
}

}  // namespace

namespace glio {

int marginalize_dense(const double* A, const double* b, int n_total, int m, double eps, double* lin_jac, double* lin_res) {
  if (!A || !b || !lin_jac || !lin_res || m <= 0 || m >= n_total) return GLIO_ERR_ARG;
  const int N = n_total, n = N - m;
  // Amm^-1 through the eigen-decomposition of the symmetrised block
  std::vector<double> Amm((size_t)m * m), w, V;
  for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * N + j] + A[(size_t)j * N + i]);
  eigh_components(Amm, m, w, V);
  std::vector<double> Ainv((size_t)m * m, 0.0);
  for (int k = 0; k < m; ++k) {
    if (!(w[k] > eps)) continue;
    const double iw = 1.0 / w[k];
    for (int i = 0; i < m; ++i) { const double vi = V[(size_t)i * m + k]; if (vi == 0.0) continue; for (int j = 0; j < m; ++j) Ainv[(size_t)i * m + j] += vi * iw * V[(size_t)j * m + k]; }
  }
  // T = Arm Amm^-1 (n x m); rows of Arm that are entirely zero (keyframes not coupled to the dropped one) stay exactly zero
  std::vector<double> T((size_t)n * m, 0.0);
  for (int i = 0; i < n; ++i) for (int k = 0; k < m; ++k) {
    const double arm = A[(size_t)(m + i) * N + k];
    if (arm == 0.0) continue;
    for (int j = 0; j < m; ++j) T[(size_t)i * m + j] += arm * Ainv[(size_t)k * m + j];
  }
  std::vector<double> Ar((size_t)n * n), br(n);
  for (int i = 0; i < n; ++i) {
    // rows that are not coupled to the dropped states (all of T's row is zero) are copied as they are
    bool coupled = false;
    for (int k = 0; k < m; ++k) if (T[(size_t)i * m + k] != 0.0) { coupled = true; break; }
    const double* arow = A + (size_t)(m + i) * N + m;
    double* orow = &Ar[(size_t)i * n];
    std::memcpy(orow, arow, sizeof(double) * (size_t)n);
    double s = b[m + i];
    if (coupled) {
      for (int k = 0; k < m; ++k) {
        const double t = T[(size_t)i * m + k];
        if (t == 0.0) continue;
        s -= t * b[k];
        const double* amr = A + (size_t)k * N + m;
        for (int j = 0; j < n; ++j) orow[j] -= t * amr[j];
      }
    }
    br[i] = s;
  }
  // symmetrise before the second decomposition (SelfAdjointEigenSolver reads the lower triangle only)
  for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) Ar[(size_t)j * n + i] = Ar[(size_t)i * n + j];
  std::vector<double> w2, V2;
  eigh_components(Ar, n, w2, V2);
  for (int k = 0; k < n; ++k) {
    const double S = w2[k] > eps ? w2[k] : 0.0, Si = w2[k] > eps ? 1.0 / w2[k] : 0.0;
    const double sq = std::sqrt(S), sqi = std::sqrt(Si);
    double vb = 0;
    for (int j = 0; j < n; ++j) { const double v = V2[(size_t)j * n + k]; lin_jac[(size_t)k * n + j] = sq * v; vb += v * br[j]; }
    lin_res[k] = sqi * vb;
  }
  return GLIO_OK;
}

}  // namespace glio

extern "C" {

int glio_marginalize(const double* A, const double* b, int n_total, int m, double eps, double* lin_jac, double* lin_res) {
  return glio::marginalize_dense(A, b, n_total, m, eps, lin_jac, lin_res);
}

void glio_marg_prior_destroy(glio_marg_prior* p) { delete p; }

int glio_marg_prior_size(const glio_marg_prior* p, int* n, int* W) {
  if (!p) return GLIO_ERR_ARG;
  if (n) *n = p->n;
  if (W) *W = p->W;
  return GLIO_OK;
}

int glio_marg_prior_get(const glio_marg_prior* p, double* lin_jac, double* lin_res, double* x0_poses, double* x0_sb, double* A_info, double* b_info) {
  if (!p) return GLIO_ERR_ARG;
  const size_t n = (size_t)p->n;
  if (lin_jac) std::memcpy(lin_jac, p->lin_jac.data(), n * n * sizeof(double));
  if (lin_res) std::memcpy(lin_res, p->lin_res.data(), n * sizeof(double));
  if (x0_poses) std::memcpy(x0_poses, p->x0_pose.data(), p->x0_pose.size() * sizeof(double));
  if (x0_sb) std::memcpy(x0_sb, p->x0_sb, 9 * sizeof(double));
  if (A_info) std::memcpy(A_info, p->A_info.data(), n * n * sizeof(double));
  if (b_info) std::memcpy(b_info, p->b_info.data(), n * sizeof(double));
  return GLIO_OK;
}

// a prior from explicit linearized_jacobians / residuals (e.g. the ones a reference MarginalizationInfo produced)
glio_marg_prior* glio_marg_prior_create(int W, const double* lin_jac, const double* lin_res, const double* x0_poses, const double* x0_sb) {
  if (W < 2 || !lin_jac || !lin_res || !x0_poses) return nullptr;
  glio_marg_prior* p = new glio_marg_prior();
  p->W = W; p->n = 6 * W + 3;
  const int n = p->n;
  p->lin_jac.assign(lin_jac, lin_jac + (size_t)n * n);
  p->lin_res.assign(lin_res, lin_res + n);
  p->x0_pose.assign(x0_poses, x0_poses + (size_t)(W - 1) * 7);
  if (x0_sb) std::memcpy(p->x0_sb, x0_sb, 9 * sizeof(double));
  p->A_info.assign((size_t)n * n, 0.0); p->b_info.assign(n, 0.0); p->c0 = 0;
  for (int k = 0; k < n; ++k) {
    const double* row = &p->lin_jac[(size_t)k * n];
    p->c0 += 0.5 * lin_res[k] * lin_res[k];
    for (int i = 0; i < n; ++i) {
      if (row[i] == 0.0) continue;
      p->b_info[i] += row[i] * lin_res[k];
      for (int j = 0; j < n; ++j) p->A_info[(size_t)i * n + j] += row[i] * row[j];
    }
  }
  // half bandwidth in the solve ordering (15 tangent dims per keyframe): prior index -> solve index
  auto solve_index = [&](int pi) { if (pi < 15) return pi; const int k = 1 + (pi - 15) / 6; return 15 * k + (pi - 15) % 6; };
  int hb = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (p->A_info[(size_t)i * n + j] != 0.0) hb = std::max(hb, std::abs(solve_index(i) - solve_index(j)));
  p->half_bandwidth = hb;
  return p;
}

}  // extern "C"
