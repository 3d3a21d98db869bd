// marginalize.cpp — K3: the dense Schur complement of MarginalizationInfo::Marginalize
// (GLIO/src/MarginalizationFactor.cpp:176-201), host C++.  It is the only genuine dense contraction on the path
// (Arm * Amm^-1 * Amr with m = 15, n-m = 6(W-1)+9 = 123 at W = 20): 123 x 15 x 123 — far too small for tensor cores to
// matter, so it stays a host routine; the LiDAR contribution to A, b comes from the device
// (glio_eval_unary with jac_kind = 1, the ambient x,y,z quaternion columns of MarginalizationFactor.cpp:9-12).
//
//   Amm <- 0.5 (Amm + Amm^T); Amm^-1 = V diag(lambda > eps ? 1/lambda : 0) V^T          (:176-182)
//   A' = Arr - Arm Amm^-1 Amr ;  b' = brr - Arm Amm^-1 bmm                              (:184-190)
//   A' = V2 diag(S) V2^T, S = lambda > eps ? lambda : 0                                  (:192-196)
//   linearized_jacobians = diag(sqrt S) V2^T ; linearized_residuals = diag(sqrt(1/S)) V2^T b'   (:198-201)
// Eigenvectors are defined up to sign / order within equal eigenvalues, so J and r are too; J^T J = A' (restricted to
// eigenvalues > eps) and J^T r are the invariants (tests/test_marginalize.py checks them against numpy).
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/glio_b200.h"

namespace {

// cyclic Jacobi eigen-decomposition of a symmetric matrix (row-major n x n): A = V diag(w) V^T, eigenvalues ascending
void jacobi_eigh(std::vector<double> a, int n, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; ++i) { diag += a[(size_t)i * n + i] * a[(size_t)i * n + i]; for (int j = i + 1; j < n; ++j) off += a[(size_t)i * n + j] * a[(size_t)i * n + j]; }
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = a[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = a[(size_t)p * n + p], aqq = a[(size_t)q * n + q];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - s * akq; a[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - s * aqk; a[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq; V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int x, int y) { return a[(size_t)x * n + x] < a[(size_t)y * n + y]; });
  w.resize(n);
  std::vector<double> Vs((size_t)n * n);
  for (int j = 0; j < n; ++j) { w[j] = a[(size_t)idx[j] * n + idx[j]]; for (int k = 0; k < n; ++k) Vs[(size_t)k * n + j] = V[(size_t)k * n + idx[j]]; }
  V.swap(Vs);
}

}  // namespace

extern "C" int glio_marginalize(const double* A, const double* b, int n_total, int m, double eps, double* lin_jac, double* lin_res) {
  if (!A || !b || !lin_jac || !lin_res || m <= 0 || m >= n_total) return GLIO_ERR_ARG;
  const int N = n_total, n = N - m;
  // Amm^-1 through the eigen-decomposition of the symmetrised block
  std::vector<double> Amm((size_t)m * m), w, V;
  for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * N + j] + A[(size_t)j * N + i]);
  jacobi_eigh(Amm, m, w, V);
  std::vector<double> Ainv((size_t)m * m, 0.0);
  for (int k = 0; k < m; ++k) {
    if (!(w[k] > eps)) continue;
    const double iw = 1.0 / w[k];
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) Ainv[(size_t)i * m + j] += V[(size_t)i * m + k] * iw * V[(size_t)j * m + k];
  }
  // T = Arm Amm^-1 (n x m)
  std::vector<double> T((size_t)n * m, 0.0);
  for (int i = 0; i < n; ++i) for (int k = 0; k < m; ++k) {
    const double arm = A[(size_t)(m + i) * N + k];
    if (arm == 0.0) continue;
    for (int j = 0; j < m; ++j) T[(size_t)i * m + j] += arm * Ainv[(size_t)k * m + j];
  }
  std::vector<double> Ar((size_t)n * n), br(n);
  for (int i = 0; i < n; ++i) {
    double s = b[m + i];
    for (int k = 0; k < m; ++k) s -= T[(size_t)i * m + k] * b[k];
    br[i] = s;
    for (int j = 0; j < n; ++j) {
      double v = A[(size_t)(m + i) * N + m + j];
      for (int k = 0; k < m; ++k) v -= T[(size_t)i * m + k] * A[(size_t)k * N + m + j];
      Ar[(size_t)i * n + j] = v;
    }
  }
  // symmetrise before the second decomposition (SelfAdjointEigenSolver reads the lower triangle only)
  for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) Ar[(size_t)j * n + i] = Ar[(size_t)i * n + j];
  std::vector<double> w2, V2;
  jacobi_eigh(Ar, n, w2, V2);
  for (int k = 0; k < n; ++k) {
    const double S = w2[k] > eps ? w2[k] : 0.0, Si = w2[k] > eps ? 1.0 / w2[k] : 0.0;
    const double sq = std::sqrt(S), sqi = std::sqrt(Si);
    double vb = 0;
    for (int j = 0; j < n; ++j) { lin_jac[(size_t)k * n + j] = sq * V2[(size_t)j * n + k]; vb += V2[(size_t)j * n + k] * br[j]; }
    lin_res[k] = sqi * vb;
  }
  return GLIO_OK;
}
