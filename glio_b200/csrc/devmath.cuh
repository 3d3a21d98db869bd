// devmath.cuh — device math shared by the grid / association kernels.
//
// Everything here that decides a float rounding or a gate (point transform, kNN distance, plane fit,
// validity, weight) is written with explicit round-to-nearest intrinsics (__dmul_rn/__dadd_rn/__fmul_rn/...)
// so that nvcc never contracts a*b+c into an FMA: the reference is a baseline x86-64 build (no FMA), and
// kNN indices / valid masks must be bit-exact (north_star).  Division and sqrt are IEEE in CUDA for double,
// and for float without -use_fast_math.
#pragma once
#include "common.cuh"

namespace glio {

struct PoseD {
  double t[3];
  double q[4];   // w, x, y, z
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// Eigen cross product: (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0)
__device__ __forceinline__ void cross3_rn(const double a[3], const double b[3], double o[3]) {
  double o0 = dsub(dmul(a[1], b[2]), dmul(a[2], b[1]));
  double o1 = dsub(dmul(a[2], b[0]), dmul(a[0], b[2]));
  double o2 = dsub(dmul(a[0], b[1]), dmul(a[1], b[0]));
  o[0] = o0; o[1] = o1; o[2] = o2;
}

// Eigen 3.3 QuaternionBase::_transformVector:  uv = u x v; uv += uv; return v + w*uv + u x uv
__device__ __forceinline__ void qrot_rn(const double q[4], const double v[3], double o[3]) {
  const double u[3] = {q[1], q[2], q[3]};
  double uv[3]; cross3_rn(u, v, uv);
  uv[0] = dadd(uv[0], uv[0]); uv[1] = dadd(uv[1], uv[1]); uv[2] = dadd(uv[2], uv[2]);
  double c[3]; cross3_rn(u, uv, c);
  o[0] = dadd(dadd(v[0], dmul(q[0], uv[0])), c[0]);
  o[1] = dadd(dadd(v[1], dmul(q[0], uv[1])), c[1]);
  o[2] = dadd(dadd(v[2], dmul(q[0], uv[2])), c[2]);
}

// po = float(q * double(pi) + t)   (GLIO/src/Estimator.cpp:1490-1498)
__device__ __forceinline__ void transform_point_f(const PoseD& P, const float pi[3], float po[3]) {
  double v[3] = {(double)pi[0], (double)pi[1], (double)pi[2]}, o[3];
  qrot_rn(P.q, v, o);
  float r0 = (float)dadd(o[0], P.t[0]);
  float r1 = (float)dadd(o[1], P.t[1]);
  float r2 = (float)dadd(o[2], P.t[2]);
  po[0] = r0; po[1] = r1; po[2] = r2;
}

// FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz in float, sequential, no FMA
__device__ __forceinline__ float l2_simple(float qx, float qy, float qz, float px, float py, float pz) {
  float d0 = fsub(qx, px), d1 = fsub(qy, py), d2 = fsub(qz, pz);
  float r = fmul(d0, d0);
  r = fadd(r, fmul(d1, d1));
  r = fadd(r, fmul(d2, d2));
  return r;
}

// unclamped cell coordinate of a value along one axis
__device__ __forceinline__ int cell_coord(float v, float o, float inv_cell) {
  return (int)floorf((v - o) * inv_cell);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int cell_of_clamped(const GridDesc& g, float x, float y, float z) {
  int cx = clampi(cell_coord(x, g.ox, g.inv_cell), 0, g.nx - 1);
  int cy = clampi(cell_coord(y, g.oy, g.inv_cell), 0, g.ny - 1);
  int cz = clampi(cell_coord(z, g.oz, g.inv_cell), 0, g.nz - 1);
  return (cz * g.ny + cy) * g.nx + cx;
}

// -----------------------------------------------------------------------------------------------
// x = ColPivHouseholderQR(A).solve(b) for a 5x3 A, b = -1 (Estimator.cpp:3649-3661), the algorithm of
// Eigen 3.3 (ColPivHouseholderQR.h computeInPlace/_solve_impl, Householder.h), sequential reductions.
// a is column-major-by-registers: a[c][r].  Fully unrolled so everything stays in registers.
// -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void swap_d(double& x, double& y) { double t = x; x = y; y = t; }

__device__ __forceinline__ int plane_solve5(double (&a)[3][5], double x[3]) {
  const double eps = 2.220446049250313e-16;
  double nupd[3], ndir[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) s = dadd(s, dmul(a[k][i], a[k][i]));
    ndir[k] = sqrt(s); nupd[k] = ndir[k];
  }
  double maxn = fmax(nupd[0], fmax(nupd[1], nupd[2]));
  const double me = dmul(maxn, eps);
  const double threshold_helper = dmul(me, me) / 5.0;
  const double downdate_thr = 1.4901161193847656e-08;  // sqrt(eps)
  int nonzero = 3;
  double hc[3];
  int ptr[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int big = k; double bn = nupd[k];
#pragma unroll
    for (int j = k + 1; j < 3; ++j) if (nupd[j] > bn) { bn = nupd[j]; big = j; }
    const double big_sq = dmul(bn, bn);
    if (nonzero == 3 && big_sq < dmul(threshold_helper, (double)(5 - k))) nonzero = k;
    ptr[k] = big;
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (big == j) {
#pragma unroll
        for (int i = 0; i < 5; ++i) swap_d(a[k][i], a[j][i]);
        swap_d(nupd[k], nupd[j]); swap_d(ndir[k], ndir[j]);
      }
    }
    // makeHouseholderInPlace on rows k..4 of column k
    double tailSq = 0.0;
#pragma unroll
    for (int i = k + 1; i < 5; ++i) tailSq = dadd(tailSq, dmul(a[k][i], a[k][i]));
    const double c0 = a[k][k];
    double beta, tau;
    if (tailSq <= 2.2250738585072014e-308) {
      tau = 0.0; beta = c0;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) a[k][i] = 0.0;
    } else {
      beta = sqrt(dadd(dmul(c0, c0), tailSq));
      if (c0 >= 0.0) beta = -beta;
      const double den = dsub(c0, beta);
#pragma unroll
      for (int i = k + 1; i < 5; ++i) a[k][i] = a[k][i] / den;
      tau = dsub(beta, c0) / beta;
    }
    a[k][k] = beta; hc[k] = tau;
    if (tau != 0.0) {
#pragma unroll
      for (int j = k + 1; j < 3; ++j) {
        double tmp = 0.0;
#pragma unroll
        for (int i = k + 1; i < 5; ++i) tmp = dadd(tmp, dmul(a[k][i], a[j][i]));
        tmp = dadd(tmp, a[j][k]);
        a[j][k] = dsub(a[j][k], dmul(tau, tmp));
#pragma unroll
        for (int i = k + 1; i < 5; ++i) a[j][i] = dsub(a[j][i], dmul(dmul(tau, a[k][i]), tmp));
      }
    }
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (nupd[j] != 0.0) {
        double temp = fabs(a[j][k]) / nupd[j];
        temp = dmul(dadd(1.0, temp), dsub(1.0, temp));
        temp = temp < 0.0 ? 0.0 : temp;
        const double rr = nupd[j] / ndir[j];
        const double temp2 = dmul(temp, dmul(rr, rr));
        if (temp2 <= downdate_thr) {
          double s = 0.0;
#pragma unroll
          for (int i = k + 1; i < 5; ++i) s = dadd(s, dmul(a[j][i], a[j][i]));
          ndir[j] = sqrt(s); nupd[j] = ndir[j];
        } else {
          nupd[j] = dmul(nupd[j], sqrt(temp));
        }
      }
    }
  }
  x[0] = x[1] = x[2] = 0.0;
  if (nonzero == 0) return 0;
  double c[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < nonzero) {
      const double tau = hc[k];
      if (tau != 0.0) {
        double tmp = 0.0;
#pragma unroll
        for (int i = k + 1; i < 5; ++i) tmp = dadd(tmp, dmul(a[k][i], c[i]));
        tmp = dadd(tmp, c[k]);
        c[k] = dsub(c[k], dmul(tau, tmp));
#pragma unroll
        for (int i = k + 1; i < 5; ++i) c[i] = dsub(c[i], dmul(dmul(tau, a[k][i]), tmp));
      }
    }
  }
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    if (i < nonzero) {
      c[i] = c[i] / a[i][i];
#pragma unroll
      for (int r = 0; r < i; ++r) c[r] = dsub(c[r], dmul(c[i], a[i][r]));
    }
  }
  // permutation: indices = identity with the transpositions applied in order
  int p0 = 0, p1 = 1, p2 = 2;
  // k = 0: swap(pidx[0], pidx[ptr0])
  if (ptr[0] == 1) { int t = p0; p0 = p1; p1 = t; } else if (ptr[0] == 2) { int t = p0; p0 = p2; p2 = t; }
  if (ptr[1] == 2) { int t = p1; p1 = p2; p2 = t; }
  double r0 = 0.0, r1 = 0.0, r2 = 0.0;
  const double c0v = nonzero > 0 ? c[0] : 0.0, c1v = nonzero > 1 ? c[1] : 0.0, c2v = nonzero > 2 ? c[2] : 0.0;
  if (p0 == 0) r0 = c0v; else if (p0 == 1) r1 = c0v; else r2 = c0v;
  if (p1 == 0) r0 = c1v; else if (p1 == 1) r1 = c1v; else r2 = c1v;
  if (p2 == 0) r0 = c2v; else if (p2 == 1) r1 = c2v; else r2 = c2v;
  x[0] = r0; x[1] = r1; x[2] = r2;
  return nonzero;
}

// Estimator.cpp:3662-3663  normInverse = 1/norm.norm(); norm.normalize()
__device__ __forceinline__ void plane_from_solution(const double x[3], double n[3], double& d) {
  const double z = dadd(dadd(dmul(x[0], x[0]), dmul(x[1], x[1])), dmul(x[2], x[2]));
  const double nn = sqrt(z);
  d = 1.0 / nn;
  if (z > 0.0) { n[0] = x[0] / nn; n[1] = x[1] / nn; n[2] = x[2] / nn; }
  else { n[0] = x[0]; n[1] = x[1]; n[2] = x[2]; }
}

// Estimator.cpp:3678-3679 (see SURVEY appendix "C++ overload subtlety")
__device__ __forceinline__ float weight_of(const double n[3], double d, float px, float py, float pz) {
  const double s = dadd(dadd(dadd(dmul(n[0], (double)px), dmul(n[1], (double)py)), dmul(n[2], (double)pz)), d);
  const float pd = (float)s;
  const float r2 = fadd(fadd(fmul(px, px), fmul(py, py)), fmul(pz, pz));
  const float rr = sqrtf(sqrtf(r2));
  return (float)dsub(1.0, dmul(0.9, (double)fabsf(pd)) / (double)rr);
}

}  // namespace glio
