// band_chol_avx2.cpp — AVX2/FMA fast path of detail::cholesky_solve (see solver.h): the same band Cholesky
// A = L L^T and the two triangular solves, restructured so that the hot loops vectorise:
//   * right-looking factorisation by panels of 4 columns.  The panel's band rows are copied once into four contiguous
//     columns, factored there (contiguous scale / in-panel updates), written back, and the same four columns then drive
//     ONE rank-4 update of the trailing rows inside the band (4 FMAs per loaded/stored element instead of 1; two rows
//     share every column load; the ragged row ends use masked stores instead of scalar tails),
//   * forward substitution with a 4-lane dot product (no scalar add-latency chain),
//   * backward substitution in axpy form (row i of L updates the right-hand side of rows i-hb..i-1: contiguous).
// Selected at run time when the CPU has AVX2 + FMA (solver.cpp); results differ from the scalar path only by
// rounding (fused multiply-adds, 4-term sums).  Built with -mavx2 -mfma for this file only.
#if defined(__x86_64__)
#include <immintrin.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "solver.h"

namespace glio {
namespace detail {

namespace {
alignas(32) const long long kMask[8] = {-1, -1, -1, -1, 0, 0, 0, 0};
inline __m256i tail_mask(int k) { return _mm256_loadu_si256(reinterpret_cast<const __m256i*>(kMask + 4 - k)); }   // first k of 4 lanes
}  // namespace

bool cholesky_solve_avx2(BandMat& A, const double* b, double* x) {
  const int n = A.n, hb = A.hb, w = hb + 1;
  double* a = A.a.data();
  constexpr int NB = 4;
  const int LC = hb + NB + 8;                      // column scratch length (zero padded, room for full-vector reads)
  static thread_local std::vector<double> colbuf;
  colbuf.assign((size_t)NB * LC, 0.0);
  double* col[NB];
  for (int c = 0; c < NB; ++c) col[c] = colbuf.data() + (size_t)c * LC;
  for (int j0 = 0; j0 < n; j0 += NB) {
    const int nb = std::min(NB, n - j0);
    const int last = std::min(j0 + nb - 1 + hb, n - 1);       // last row any panel column reaches
    const int np = last - j0 + 1;                             // panel rows j0 .. last; col[c][r - j0] = A(r, j0 + c)
    // ---- copy the panel out of the band (strided reads, once)
    for (int c = 0; c < NB; ++c) {
      double* cc = col[c];
      if (c < nb) {
        const int j = j0 + c, lim = std::min(j + hb, n - 1);
        for (int r = j0; r < j; ++r) cc[r - j0] = 0.0;                                        // above the diagonal: not part of L
        for (int r = j; r <= last; ++r) cc[r - j0] = r <= lim ? a[(size_t)r * w + hb - (r - j)] : 0.0;
      } else for (int r = 0; r < np; ++r) cc[r] = 0.0;
      for (int k = 0; k < 8; ++k) cc[np + k] = 0.0;
    }
    // ---- factor the panel in the contiguous copy
    for (int c = 0; c < nb; ++c) {
      double* cc = col[c];
      const double d = cc[c];
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      const double l = std::sqrt(d), linv = 1.0 / l;
      for (int r = c + 1; r < np; ++r) cc[r] *= linv;
      cc[c] = l;
      for (int c2 = c + 1; c2 < nb; ++c2) {
        const double f = cc[c2];                                                               // L(j0+c2, j0+c)
        double* c2p = col[c2];
        for (int r = c2; r < np; ++r) c2p[r] -= cc[r] * f;
      }
    }
    // ---- write the factored panel back (strided writes, once)
    for (int c = 0; c < nb; ++c) {
      const int j = j0 + c, lim = std::min(j + hb, n - 1);
      const double* cc = col[c];
      for (int r = j; r <= lim; ++r) a[(size_t)r * w + hb - (r - j)] = cc[r - j0];
    }
    const int s = j0 + nb;                                      // first trailing column
    if (s >= n) break;
    // ---- rank-nb update of the trailing triangle: A(i, s..i) -= sum_c L(i,c) * L(s..i, c); two rows per pass
    const double* c0 = col[0] + nb; const double* c1 = col[1] + nb; const double* c2 = col[2] + nb; const double* c3 = col[3] + nb;   // index r - s
    int i = s;
    for (; i + 1 <= last; i += 2) {
      double* ra = a + (size_t)i * w + hb - (i - s);                                           // A(i, s)
      double* rb = a + (size_t)(i + 1) * w + hb - (i + 1 - s);                                 // A(i+1, s)
      const int la = i - s + 1, lb = la + 1;
      const int ia = i - s, ib = ia + 1;
      const __m256d a0 = _mm256_set1_pd(c0[ia]), a1 = _mm256_set1_pd(c1[ia]), a2 = _mm256_set1_pd(c2[ia]), a3 = _mm256_set1_pd(c3[ia]);
      const __m256d b0 = _mm256_set1_pd(c0[ib]), b1 = _mm256_set1_pd(c1[ib]), b2 = _mm256_set1_pd(c2[ib]), b3 = _mm256_set1_pd(c3[ib]);
      int u = 0;
      for (; u + 4 <= la; u += 4) {
        const __m256d v0 = _mm256_loadu_pd(c0 + u), v1 = _mm256_loadu_pd(c1 + u), v2 = _mm256_loadu_pd(c2 + u), v3 = _mm256_loadu_pd(c3 + u);
        // four chained fnmadds per row vector (independent across vectors and rows: the out-of-order core overlaps them)
        __m256d xa = _mm256_loadu_pd(ra + u), xb = _mm256_loadu_pd(rb + u);
        xa = _mm256_fnmadd_pd(a0, v0, xa); xb = _mm256_fnmadd_pd(b0, v0, xb);
        xa = _mm256_fnmadd_pd(a1, v1, xa); xb = _mm256_fnmadd_pd(b1, v1, xb);
        xa = _mm256_fnmadd_pd(a2, v2, xa); xb = _mm256_fnmadd_pd(b2, v2, xb);
        xa = _mm256_fnmadd_pd(a3, v3, xa); xb = _mm256_fnmadd_pd(b3, v3, xb);
        _mm256_storeu_pd(ra + u, xa); _mm256_storeu_pd(rb + u, xb);
      }
      {  // ragged end: la - u in [0,3] elements of row a, lb - u in [1,4] of row b
        const __m256d v0 = _mm256_loadu_pd(c0 + u), v1 = _mm256_loadu_pd(c1 + u), v2 = _mm256_loadu_pd(c2 + u), v3 = _mm256_loadu_pd(c3 + u);
        const __m256i ma = tail_mask(la - u), mb = tail_mask(lb - u);
        __m256d pa = _mm256_fmadd_pd(a1, v1, _mm256_mul_pd(a0, v0)), qa = _mm256_fmadd_pd(a3, v3, _mm256_mul_pd(a2, v2));
        __m256d pb = _mm256_fmadd_pd(b1, v1, _mm256_mul_pd(b0, v0)), qb = _mm256_fmadd_pd(b3, v3, _mm256_mul_pd(b2, v2));
        _mm256_maskstore_pd(ra + u, ma, _mm256_sub_pd(_mm256_maskload_pd(ra + u, ma), _mm256_add_pd(pa, qa)));
        _mm256_maskstore_pd(rb + u, mb, _mm256_sub_pd(_mm256_maskload_pd(rb + u, mb), _mm256_add_pd(pb, qb)));
      }
    }
    if (i <= last) {                                                                          // odd row left
      double* ra = a + (size_t)i * w + hb - (i - s);
      const int la = i - s + 1, ia = i - s;
      const __m256d a0 = _mm256_set1_pd(c0[ia]), a1 = _mm256_set1_pd(c1[ia]), a2 = _mm256_set1_pd(c2[ia]), a3 = _mm256_set1_pd(c3[ia]);
      for (int u = 0; u < la; u += 4) {
        const __m256i ma = tail_mask(std::min(4, la - u));
        const __m256d v0 = _mm256_loadu_pd(c0 + u), v1 = _mm256_loadu_pd(c1 + u), v2 = _mm256_loadu_pd(c2 + u), v3 = _mm256_loadu_pd(c3 + u);
        __m256d pa = _mm256_fmadd_pd(a1, v1, _mm256_mul_pd(a0, v0)), qa = _mm256_fmadd_pd(a3, v3, _mm256_mul_pd(a2, v2));
        _mm256_maskstore_pd(ra + u, ma, _mm256_sub_pd(_mm256_maskload_pd(ra + u, ma), _mm256_add_pd(pa, qa)));
      }
    }
  }
  // ---- forward: L y = b  (row-wise dot products, 4 lanes)
  for (int i = 0; i < n; ++i) {
    const int k0 = std::max(0, i - hb);
    const double* ri = a + (size_t)i * w + hb - i;                                            // ri[k] = L(i,k)
    __m256d acc = _mm256_setzero_pd();
    int k = k0;
    for (; k + 4 <= i; k += 4) acc = _mm256_fmadd_pd(_mm256_loadu_pd(ri + k), _mm256_loadu_pd(x + k), acc);
    double t[4]; _mm256_storeu_pd(t, acc);
    double sdot = (t[0] + t[1]) + (t[2] + t[3]);
    for (; k < i; ++k) sdot += ri[k] * x[k];
    x[i] = (b[i] - sdot) / ri[i];
  }
  // ---- backward: L^T x = y in axpy form (row i of L updates rows i-hb .. i-1)
  for (int i = n - 1; i >= 0; --i) {
    const double* ri = a + (size_t)i * w + hb - i;
    const double xi = x[i] / ri[i];
    x[i] = xi;
    const int k0 = std::max(0, i - hb);
    const __m256d vx = _mm256_set1_pd(xi);
    int k = k0;
    for (; k + 4 <= i; k += 4) _mm256_storeu_pd(x + k, _mm256_fnmadd_pd(_mm256_loadu_pd(ri + k), vx, _mm256_loadu_pd(x + k)));
    for (; k < i; ++k) x[k] -= ri[k] * xi;
  }
  for (int i = 0; i < n; ++i) if (!std::isfinite(x[i])) return false;
  return true;
}

// y = A x for a symmetric matrix in lower-band storage: one pass over the stored rows; row i contributes the dot product
// A(i, k0..i-1) . x to y[i] and the axpy x[i] * A(i, k0..i-1) to y[k0..i-1] from the same loads (4 lanes each).
void symv_avx2(const BandMat& A, const double* x, double* y) {
  const int n = A.n, hb = A.hb, w = hb + 1;
  const double* a = A.a.data();
  for (int i = 0; i < n; ++i) y[i] = 0.0;
  for (int i = 0; i < n; ++i) {
    const int k0 = std::max(0, i - hb);
    const double* ri = a + (size_t)i * w + hb - i;
    const double xi = x[i];
    const __m256d vx = _mm256_set1_pd(xi);
    __m256d acc = _mm256_setzero_pd();
    int k = k0;
    for (; k + 4 <= i; k += 4) {
      const __m256d r = _mm256_loadu_pd(ri + k);
      acc = _mm256_fmadd_pd(r, _mm256_loadu_pd(x + k), acc);
      _mm256_storeu_pd(y + k, _mm256_fmadd_pd(r, vx, _mm256_loadu_pd(y + k)));
    }
    double t[4]; _mm256_storeu_pd(t, acc);
    double s = (t[0] + t[1]) + (t[2] + t[3]);
    for (; k < i; ++k) { s += ri[k] * x[k]; y[k] += ri[k] * xi; }
    y[i] += s + ri[i] * xi;
  }
}

}  // namespace detail
}  // namespace glio
#endif
