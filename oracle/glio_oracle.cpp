// glio_oracle.cpp — CPU restatement of GLIO's LiDAR association + factor evaluation.
// TEST INFRASTRUCTURE ONLY (see glio_oracle.h).  PARITY UNPINNED (no reference tests exist).
// Build: g++ -O3 -ffp-contract=off -fopenmp -shared -fPIC   (oracle/Makefile)
#include "glio_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------------
// Jet<N>: first-order dual numbers, the arithmetic of ceres::Jet (ceres.tgz::include/ceres/jet.h).
// ------------------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1.0; }
};
template <int N> Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // jet.h: h = f/g ; dh = (df - h dg)/g
  Jet<N> h; const double gi = 1.0 / g.a; const double fg = f.a * gi; h.a = fg;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
  return h; }
template <int N> Jet<N> jsqrt(const Jet<N>& f) { Jet<N> h; const double t = std::sqrt(f.a); h.a = t; const double s = 1.0 / (2.0 * t); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
inline double jsqrt(double x) { return std::sqrt(x); }
template <int N> Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }

// ------------------------------------------------------------------------------------------------
// Eigen 3.3 geometry, restated (Eigen is not vendored; formulas from Eigen/src/Geometry/Quaternion.h).
// Quaternions are (w,x,y,z).
// ------------------------------------------------------------------------------------------------
template <class T> inline void cross3(const T a[3], const T b[3], T o[3]) {
  // Eigen cross: (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0)
  T o0 = a[1] * b[2] - a[2] * b[1];
  T o1 = a[2] * b[0] - a[0] * b[2];
  T o2 = a[0] * b[1] - a[1] * b[0];
  o[0] = o0; o[1] = o1; o[2] = o2;
}
// QuaternionBase::_transformVector:  uv = u x v; uv += uv; return v + w*uv + u x uv
template <class T> inline void qrot(const T q[4], const T v[3], T o[3]) {
  const T u[3] = {q[1], q[2], q[3]};
  T uv[3]; cross3(u, v, uv);
  uv[0] = uv[0] + uv[0]; uv[1] = uv[1] + uv[1]; uv[2] = uv[2] + uv[2];
  T c[3]; cross3(u, uv, c);
  T r0 = v[0] + q[0] * uv[0] + c[0];
  T r1 = v[1] + q[0] * uv[1] + c[1];
  T r2 = v[2] + q[0] * uv[2] + c[2];
  o[0] = r0; o[1] = r1; o[2] = r2;
}
// Eigen quat product a*b
template <class T> inline void qmul(const T a[4], const T b[4], T o[4]) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  T z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
// QuaternionBase::inverse(): conjugate / squaredNorm (if > 0)
template <class T> inline void qinv(const T q[4], T o[4]) {
  T n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  o[0] = q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = -q[3] / n2;
}

// ------------------------------------------------------------------------------------------------
// FLANN L2_Simple<float>: result += diff*diff sequentially, in float, no FMA (flann/algorithms/dist.h).
// ------------------------------------------------------------------------------------------------
inline float l2_simple(const float* a, const float* b) {
  float result = 0.0f;
  float d0 = a[0] - b[0]; result += d0 * d0;
  float d1 = a[1] - b[1]; result += d1 * d1;
  float d2 = a[2] - b[2]; result += d2 * d2;
  return result;
}

struct Top5 {
  float d[6];
  int32_t i[6];
  int n;
  Top5() : n(0) { for (int k = 0; k < 6; ++k) { d[k] = std::numeric_limits<float>::infinity(); i[k] = -1; } }
  inline float worst() const { return d[4]; }
  // keeps the 6 best by (dist, index) so ties at the 5/6 boundary are visible
  inline void push(float dist, int32_t idx) {
    if (!(dist < d[5] || (dist == d[5] && idx < i[5]) || i[5] < 0)) return;
    int k = 5;
    while (k > 0 && (i[k - 1] < 0 || dist < d[k - 1] || (dist == d[k - 1] && idx < i[k - 1]))) {
      d[k] = d[k - 1]; i[k] = i[k - 1]; --k;
    }
    d[k] = dist; i[k] = idx;
  }
};

// ------------------------------------------------------------------------------------------------
// kd-tree in the style of FLANN KDTreeSingleIndex (leaf_max_size 15 as PCL's KdTreeFLANN passes,
// points reordered, bounding-box split, exact search with per-dimension distance bookkeeping).
// Results are independent of the tree shape (exact search, ties broken by index).
// ------------------------------------------------------------------------------------------------
struct KdTree {
  struct Node { int32_t left, right; int32_t lo, hi; int dim; float divlow, divhigh; };
  std::vector<Node> nodes;
  std::vector<float> pts;       // reordered xyz
  std::vector<int32_t> ids;     // original index of reordered point
  float bbmin[3], bbmax[3];
  int64_t M = 0;
  static const int kLeaf = 15;

  int build(int32_t lo, int32_t hi, std::vector<int32_t>& order, const float* xyz) {
    Node nd; nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1; nd.dim = -1; nd.divlow = nd.divhigh = 0;
    int me = (int)nodes.size(); nodes.push_back(nd);
    if (hi - lo <= kLeaf) return me;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int32_t k = lo; k < hi; ++k) for (int d = 0; d < 3; ++d) {
      float v = xyz[3 * (int64_t)order[k] + d]; mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v); }
    int dim = 0; float span = mx[0] - mn[0];
    for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > span) { span = mx[d] - mn[d]; dim = d; }
    if (span <= 0) return me;  // all identical: keep as a (big) leaf
    int32_t mid = lo + (hi - lo) / 2;
    std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                     [&](int32_t a, int32_t b) { float va = xyz[3 * (int64_t)a + dim], vb = xyz[3 * (int64_t)b + dim]; return va < vb || (va == vb && a < b); });
    float dl = -1e30f, dh = 1e30f;
    for (int32_t k = lo; k < mid; ++k) dl = std::max(dl, xyz[3 * (int64_t)order[k] + dim]);
    for (int32_t k = mid; k < hi; ++k) dh = std::min(dh, xyz[3 * (int64_t)order[k] + dim]);
    nodes[me].dim = dim; nodes[me].divlow = dl; nodes[me].divhigh = dh;
    int l = build(lo, mid, order, xyz); nodes[me].left = l;
    int r = build(mid, hi, order, xyz); nodes[me].right = r;
    return me;
  }
  void init(const float* xyz, int64_t m) {
    M = m; std::vector<int32_t> order(m); for (int64_t i = 0; i < m; ++i) order[i] = (int32_t)i;
    for (int d = 0; d < 3; ++d) { bbmin[d] = 1e30f; bbmax[d] = -1e30f; }
    for (int64_t i = 0; i < m; ++i) for (int d = 0; d < 3; ++d) { bbmin[d] = std::min(bbmin[d], xyz[3 * i + d]); bbmax[d] = std::max(bbmax[d], xyz[3 * i + d]); }
    nodes.reserve((size_t)(2 * m / kLeaf + 16));
    if (m > 0) build(0, (int32_t)m, order, xyz);
    pts.resize(3 * (size_t)m); ids.resize((size_t)m);
    for (int64_t i = 0; i < m; ++i) { ids[i] = order[i]; for (int d = 0; d < 3; ++d) pts[3 * i + d] = xyz[3 * (int64_t)order[i] + d]; }
  }
  void search(int node, const float* q, float mindist, float dists[3], Top5& res) const {
    const Node& nd = nodes[node];
    if (nd.dim < 0) {
      for (int32_t k = nd.lo; k < nd.hi; ++k) res.push(l2_simple(q, &pts[3 * (size_t)k]), ids[k]);
      return;
    }
    int d = nd.dim; float val = q[d]; float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
    int best, other; float cut;
    if (diff1 + diff2 < 0) { best = nd.left; other = nd.right; cut = diff2 * diff2; }
    else { best = nd.right; other = nd.left; cut = diff1 * diff1; }
    search(best, q, mindist, dists, res);
    float dst = dists[d];
    float md = mindist + cut - dst;
    dists[d] = cut;
    // '<=' (not '<') so equal-distance candidates in the other branch are still seen: ties are resolved by index
    // 1e-4 relative slack: md is a bound in exact arithmetic, candidates are compared by their float l2_simple value
    if (md * 0.9999f <= res.d[5] || res.i[5] < 0) search(other, q, md, dists, res);
    dists[d] = dst;
  }
  void knn5(const float* q, int32_t* idx5, float* sqd5) const {
    Top5 res; float dists[3] = {0, 0, 0}; float dsq = 0;
    for (int d = 0; d < 3; ++d) {
      if (q[d] < bbmin[d]) { dists[d] = (q[d] - bbmin[d]) * (q[d] - bbmin[d]); dsq += dists[d]; }
      if (q[d] > bbmax[d]) { dists[d] = (q[d] - bbmax[d]) * (q[d] - bbmax[d]); dsq += dists[d]; }
    }
    if (M > 0) search(0, q, dsq, dists, res);
    for (int k = 0; k < 5; ++k) { idx5[k] = res.i[k]; sqd5[k] = res.d[k]; }
  }
};

// ------------------------------------------------------------------------------------------------
// Eigen 3.3 ColPivHouseholderQR<Matrix<double,5,3>>::solve, restated (Eigen/src/QR/ColPivHouseholderQR.h
// computeInPlace + _solve_impl; Eigen/src/Householder/Householder.h makeHouseholder /
// applyHouseholderOnTheLeft).  Reductions are accumulated sequentially (Eigen may use SSE2 packets for
// the dynamic-size tails; that changes results at the 1-ulp level only — see DESIGN.md).
// A is row-major 5x3 on input.
// ------------------------------------------------------------------------------------------------
int colpiv_qr_solve_5x3(const double Ain[15], const double b_in[5], double x[3]) {
  const int rows = 5, cols = 3;
  double a[5][3];
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 3; ++j) a[i][j] = Ain[3 * i + j];
  double hc[3]; int perm_tr[3];
  double nrm_upd[3], nrm_dir[3];
  for (int k = 0; k < cols; ++k) {
    double s = 0; for (int i = 0; i < rows; ++i) s += a[i][k] * a[i][k];
    nrm_dir[k] = std::sqrt(s); nrm_upd[k] = nrm_dir[k];
  }
  const double eps = std::numeric_limits<double>::epsilon();
  double maxn = std::max(nrm_upd[0], std::max(nrm_upd[1], nrm_upd[2]));
  const double threshold_helper = (maxn * eps) * (maxn * eps) / double(rows);
  const double norm_downdate_threshold = std::sqrt(eps);
  int nonzero_pivots = cols;
  for (int k = 0; k < cols; ++k) {
    int big = k; double bn = nrm_upd[k];
    for (int j = k + 1; j < cols; ++j) if (nrm_upd[j] > bn) { bn = nrm_upd[j]; big = j; }
    double big_sq = bn * bn;
    if (nonzero_pivots == cols && big_sq < threshold_helper * double(rows - k)) nonzero_pivots = k;
    perm_tr[k] = big;
    if (k != big) {
      for (int i = 0; i < rows; ++i) std::swap(a[i][k], a[i][big]);
      std::swap(nrm_upd[k], nrm_upd[big]); std::swap(nrm_dir[k], nrm_dir[big]);
    }
    // makeHouseholderInPlace on a[k..rows-1][k]
    double tailSq = 0; for (int i = k + 1; i < rows; ++i) tailSq += a[i][k] * a[i][k];
    double c0 = a[k][k]; double beta, tau;
    const double tol = std::numeric_limits<double>::min();
    if (tailSq <= tol) { tau = 0; beta = c0; for (int i = k + 1; i < rows; ++i) a[i][k] = 0; }
    else {
      beta = std::sqrt(c0 * c0 + tailSq); if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < rows; ++i) a[i][k] = a[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    a[k][k] = beta; hc[k] = tau;
    // apply H to the trailing columns
    if (tau != 0) {
      for (int j = k + 1; j < cols; ++j) {
        double tmp = 0; for (int i = k + 1; i < rows; ++i) tmp += a[i][k] * a[i][j];
        tmp += a[k][j];
        a[k][j] -= tau * tmp;
        for (int i = k + 1; i < rows; ++i) a[i][j] -= tau * a[i][k] * tmp;
      }
    }
    // norm downdate (LAPACK-style, Eigen 3.3)
    for (int j = k + 1; j < cols; ++j) {
      if (nrm_upd[j] != 0) {
        double temp = std::fabs(a[k][j]) / nrm_upd[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0 ? 0 : temp;
        double rr = nrm_upd[j] / nrm_dir[j];
        double temp2 = temp * (rr * rr);
        if (temp2 <= norm_downdate_threshold) {
          double s = 0; for (int i = k + 1; i < rows; ++i) s += a[i][j] * a[i][j];
          nrm_dir[j] = std::sqrt(s); nrm_upd[j] = nrm_dir[j];
        } else {
          nrm_upd[j] *= std::sqrt(temp);
        }
      }
    }
  }
  // _solve_impl
  x[0] = x[1] = x[2] = 0;
  if (nonzero_pivots == 0) return 0;
  double c[5]; for (int i = 0; i < 5; ++i) c[i] = b_in[i];
  for (int k = 0; k < nonzero_pivots; ++k) {
    double tau = hc[k];
    if (tau != 0) {
      double tmp = 0; for (int i = k + 1; i < rows; ++i) tmp += a[i][k] * c[i];
      tmp += c[k];
      c[k] -= tau * tmp;
      for (int i = k + 1; i < rows; ++i) c[i] -= tau * a[i][k] * tmp;
    }
  }
  // upper-triangular solve, column-oriented back substitution
  for (int i = nonzero_pivots - 1; i >= 0; --i) {
    c[i] /= a[i][i];
    for (int r = 0; r < i; ++r) c[r] -= c[i] * a[r][i];
  }
  // column permutation: P = T_0 T_1 ... ; indices built by applying transpositions in order
  int pidx[3] = {0, 1, 2};
  for (int k = 0; k < cols; ++k) std::swap(pidx[k], pidx[perm_tr[k]]);
  for (int i = 0; i < nonzero_pivots; ++i) x[pidx[i]] = c[i];
  return nonzero_pivots;
}

struct PlaneFit { double n[3]; double d; };
// Estimator.cpp:3661-3663: norm = solve; normInverse = 1/norm.norm(); norm.normalize()
inline void plane_from_solution(const double x[3], PlaneFit* pf) {
  double z = (x[0] * x[0] + x[1] * x[1]) + x[2] * x[2];
  double nn = std::sqrt(z);
  pf->d = 1.0 / nn;
  if (z > 0) { pf->n[0] = x[0] / nn; pf->n[1] = x[1] / nn; pf->n[2] = x[2] / nn; }
  else { pf->n[0] = x[0]; pf->n[1] = x[1]; pf->n[2] = x[2]; }
}

// Estimator.cpp:3678-3679 with the float/double overloads resolved as in SURVEY appendix:
//   float pd = n.x*pm.x + n.y*pm.y + n.z*pm.z + d      (double arithmetic, rounded to float)
//   float weight = 1 - 0.9*fabs(pd)/sqrt(sqrt(pm.x*pm.x + pm.y*pm.y + pm.z*pm.z))
//   (inner sum and both sqrt in float; 0.9*|pd| / ... and 1 - ... in double; rounded to float)
inline float weight_of(const PlaneFit& pf, const float pm[3]) {
  float pd = (float)(pf.n[0] * (double)pm[0] + pf.n[1] * (double)pm[1] + pf.n[2] * (double)pm[2] + pf.d);
  float r2 = pm[0] * pm[0] + pm[1] * pm[1] + pm[2] * pm[2];
  float rr = sqrtf(sqrtf(r2));
  float w = (float)(1.0 - 0.9 * (double)fabsf(pd) / (double)rr);
  return w;
}

template <class T> inline void transform_pt(const double q[4], const double t[3], const T* in, float out[3]) {
  double v[3] = {(double)in[0], (double)in[1], (double)in[2]}, o[3];
  qrot<double>(q, v, o);
  out[0] = (float)(o[0] + t[0]); out[1] = (float)(o[1] + t[1]); out[2] = (float)(o[2] + t[2]);
}

inline void knn_dispatch(const KdTree* tree, const float* map, int64_t M, const float q[3], int32_t idx5[5], float sqd5[5]) {
  if (tree) { tree->knn5(q, idx5, sqd5); return; }
  Top5 res;
  for (int64_t m = 0; m < M; ++m) res.push(l2_simple(q, map + 3 * m), (int32_t)m);
  for (int k = 0; k < 5; ++k) { idx5[k] = res.i[k]; sqd5[k] = res.d[k]; }
}

}  // namespace

// =================================================================================================
extern "C" {

void go_transform_points(const float* in_xyz, int64_t n, const double t[3], const double q[4], float* out_xyz) {
  for (int64_t i = 0; i < n; ++i) transform_pt<float>(q, t, in_xyz + 3 * i, out_xyz + 3 * i);
}

int64_t go_voxel_filter(const float* xyz, int64_t n, float leaf, int order_mode, float* out_xyz, int32_t* out_idx) {
  if (n <= 0) return 0;
  // setLeafSize: inverse_leaf_size_ = Eigen::Array4f::Ones() / leaf_size_.array()  (float division)
  const float inv = 1.0f / leaf;
  // getMinMax3D: float min / max over the points
  float mn[3] = {xyz[0], xyz[1], xyz[2]}, mx[3] = {xyz[0], xyz[1], xyz[2]};
  for (int64_t i = 1; i < n; ++i) for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], xyz[3 * i + d]); mx[d] = std::max(mx[d], xyz[3 * i + d]); }
  // leaf-size sanity check of applyFilter
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) return -1;
  int min_b[3], max_b[3], div_b[3];
  for (int d = 0; d < 3; ++d) { min_b[d] = (int)std::floor(mn[d] * inv); max_b[d] = (int)std::floor(mx[d] * inv); div_b[d] = max_b[d] - min_b[d] + 1; }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  struct Cpi { unsigned int idx; unsigned int cloud_point_index; bool operator<(const Cpi& o) const { return idx < o.idx; } };
  std::vector<Cpi> iv; iv.reserve((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    const int ijk0 = (int)(std::floor(xyz[3 * i] * inv) - (float)min_b[0]);
    const int ijk1 = (int)(std::floor(xyz[3 * i + 1] * inv) - (float)min_b[1]);
    const int ijk2 = (int)(std::floor(xyz[3 * i + 2] * inv) - (float)min_b[2]);
    iv.push_back(Cpi{(unsigned int)(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]), (unsigned int)i});
  }
  if (order_mode == 0) std::sort(iv.begin(), iv.end());
  else std::stable_sort(iv.begin(), iv.end());
  int64_t m = 0;
  for (size_t a = 0; a < iv.size();) {
    size_t b = a + 1;
    while (b < iv.size() && iv[b].idx == iv[a].idx) ++b;
    // CentroidPoint<PointXYZI>: AccumulatorXYZ sums Eigen::Vector3f, get() divides by the count
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (size_t k = a; k < b; ++k) { const float* q = xyz + 3 * (size_t)iv[k].cloud_point_index; sx += q[0]; sy += q[1]; sz += q[2]; }
    const float cnt = (float)(b - a);
    out_xyz[3 * m] = sx / cnt; out_xyz[3 * m + 1] = sy / cnt; out_xyz[3 * m + 2] = sz / cnt;
    if (out_idx) out_idx[m] = (int32_t)iv[a].idx;
    ++m; a = b;
  }
  return m;
}

void go_knn5_brute(const float* map_xyz, int64_t M, const float* qry_xyz, int64_t Q, int32_t* idx5, float* sqd5, uint8_t* tie) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < Q; ++i) {
    Top5 res; const float* q = qry_xyz + 3 * i;
    for (int64_t m = 0; m < M; ++m) res.push(l2_simple(q, map_xyz + 3 * m), (int32_t)m);
    for (int k = 0; k < 5; ++k) { idx5[5 * i + k] = res.i[k]; sqd5[5 * i + k] = res.d[k]; }
    if (tie) { uint8_t t = 0; for (int k = 0; k < 5; ++k) if (res.i[k + 1] >= 0 && res.d[k] == res.d[k + 1]) t = 1; tie[i] = t; }
  }
}

void* go_kdtree_build(const float* xyz, int64_t M) { KdTree* t = new KdTree(); t->init(xyz, M); return t; }
void go_kdtree_free(void* tree) { delete (KdTree*)tree; }
void go_kdtree_knn5(const void* tree, const float* qry_xyz, int64_t Q, int32_t* idx5, float* sqd5) {
  const KdTree* t = (const KdTree*)tree;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < Q; ++i) t->knn5(qry_xyz + 3 * i, idx5 + 5 * i, sqd5 + 5 * i);
}

int go_plane_solve5(const double A[15], double x[3]) {
  const double b[5] = {-1, -1, -1, -1, -1};
  return colpiv_qr_solve_5x3(A, b, x);
}

int64_t go_assoc_scan_to_map(const go_assoc_params* prm, const float* map_xyz, int64_t M, const void* kdtree,
                             const float* scan_xyz, int64_t Q, const double t[3], const double q[4],
                             uint8_t* status, int32_t* idx5, float* sqd5, float* pm_out, double* plane,
                             float* nsd, float* weight, double* score, int nthreads) {
  const KdTree* tree = (const KdTree*)kdtree;
  int64_t nvalid = 0;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) reduction(+ : nvalid)
  for (int64_t i = 0; i < Q; ++i) {
    float pm[3]; transform_pt<float>(q, t, scan_xyz + 3 * i, pm);                  // :3645 transformPoint
    int32_t id[5]; float sd[5]; knn_dispatch(tree, map_xyz, M, pm, id, sd);         // :3647 nearestKSearch
    uint8_t st = GO_VALID; PlaneFit pf = {{0, 0, 0}, 0}; float w = 0;
    if (M >= 5 && (double)sd[4] < prm->kd_max_radius) {                             // :3651 (Q1)
      double A[15];
      for (int j = 0; j < 5; ++j) for (int c = 0; c < 3; ++c) A[3 * j + c] = (double)map_xyz[3 * (int64_t)id[j] + c];
      double x[3]; go_plane_solve5(A, x);                                           // :3661
      plane_from_solution(x, &pf);                                                  // :3662-3663
      bool ok = true;
      for (int j = 0; j < 5; ++j) {                                                 // :3667-3674
        double v = pf.n[0] * A[3 * j] + pf.n[1] * A[3 * j + 1] + pf.n[2] * A[3 * j + 2] + pf.d;
        if (std::fabs(v) > prm->surf_dist_thres) { ok = false; break; }
      }
      if (ok) {
        w = weight_of(pf, pm);                                                      // :3678-3679
        if ((double)w > prm->weight_min) st = GO_VALID; else st = GO_FAIL_WEIGHT;   // :3681
      } else st = GO_FAIL_PLANE;
    } else st = GO_FAIL_RADIUS;
    if (status) status[i] = st;
    if (idx5) for (int k = 0; k < 5; ++k) idx5[5 * i + k] = id[k];
    if (sqd5) for (int k = 0; k < 5; ++k) sqd5[5 * i + k] = sd[k];
    if (pm_out) for (int k = 0; k < 3; ++k) pm_out[3 * i + k] = pm[k];
    if (plane) { plane[4 * i] = pf.n[0]; plane[4 * i + 1] = pf.n[1]; plane[4 * i + 2] = pf.n[2]; plane[4 * i + 3] = pf.d; }
    if (st == GO_VALID) {
      ++nvalid;
      if (nsd) {                                                                    // :3683-3687 (float stores)
        nsd[4 * i + 0] = (float)((double)w * pf.n[0]); nsd[4 * i + 1] = (float)((double)w * pf.n[1]);
        nsd[4 * i + 2] = (float)((double)w * pf.n[2]); nsd[4 * i + 3] = (float)((double)w * pf.d);
      }
      if (weight) weight[i] = w;
      if (score) score[i] = prm->lidar_const * (double)w;                           // :3692
    } else {
      if (nsd) nsd[4 * i] = nsd[4 * i + 1] = nsd[4 * i + 2] = nsd[4 * i + 3] = 0;
      if (weight) weight[i] = (st == GO_FAIL_WEIGHT) ? w : 0;
      if (score) score[i] = 0;
    }
  }
  return nvalid;
}

int64_t go_assoc_pair(const go_assoc_params* prm, const float* cur_xyz, int64_t Qc, const double t_c[3],
                      const double q_c[4], const float* oth_xyz, int64_t Qo, const double t_o[3],
                      const double q_o[4], int use_kdtree, uint8_t* status, int32_t* idx5, float* sqd5,
                      float* weight, double* score, double* normal_cent, int nthreads) {
  // :3716,3725 world clouds of both frames (float)
  std::vector<float> cw(3 * (size_t)Qc), ow(3 * (size_t)Qo);
  go_transform_points(cur_xyz, Qc, t_c, q_c, cw.data());
  go_transform_points(oth_xyz, Qo, t_o, q_o, ow.data());
  KdTree* tree = nullptr;
  if (use_kdtree) { tree = new KdTree(); tree->init(ow.data(), Qo); }               // :3729-3731
  int64_t nvalid = 0;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) reduction(+ : nvalid)
  for (int64_t i = 0; i < Qc; ++i) {
    const float* pm = &cw[3 * (size_t)i];
    int32_t id[5]; float sd[5]; knn_dispatch(tree, ow.data(), Qo, pm, id, sd);      // :3746
    uint8_t st; float w = 0; double nc[6] = {0, 0, 0, 0, 0, 0};
    if (Qo >= 5 && (double)sd[4] < prm->batch_max_radius) {                         // :3751
      double A[15], Al[15]; double cx = 0, cy = 0, cz = 0;
      for (int j = 0; j < 5; ++j) {                                                 // :3753-3763
        for (int c = 0; c < 3; ++c) { A[3 * j + c] = (double)ow[3 * (size_t)id[j] + c]; Al[3 * j + c] = (double)oth_xyz[3 * (int64_t)id[j] + c]; }
        cx += Al[3 * j]; cy += Al[3 * j + 1]; cz += Al[3 * j + 2];
      }
      nc[3] = cx / 5.; nc[4] = cy / 5.; nc[5] = cz / 5.;                            // :3764-3766
      double x[3]; go_plane_solve5(A, x); PlaneFit pf; plane_from_solution(x, &pf); // :3768-3770
      double xl[3]; go_plane_solve5(Al, xl); PlaneFit pl; plane_from_solution(xl, &pl); // :3771-3772
      bool ok = true;
      for (int j = 0; j < 5; ++j) {                                                 // :3775-3782
        double v = pf.n[0] * A[3 * j] + pf.n[1] * A[3 * j + 1] + pf.n[2] * A[3 * j + 2] + pf.d;
        if (std::fabs(v) > prm->batch_dist_thres) { ok = false; break; }
      }
      if (ok) {
        w = weight_of(pf, pm);                                                      // :3785-3786
        if ((double)w > prm->weight_min) { st = GO_VALID; nc[0] = pl.n[0]; nc[1] = pl.n[1]; nc[2] = pl.n[2]; }
        else st = GO_FAIL_WEIGHT;
      } else st = GO_FAIL_PLANE;
    } else st = GO_FAIL_RADIUS;
    if (status) status[i] = st;
    if (idx5) for (int k = 0; k < 5; ++k) idx5[5 * i + k] = id[k];
    if (sqd5) for (int k = 0; k < 5; ++k) sqd5[5 * i + k] = sd[k];
    if (st == GO_VALID) {
      ++nvalid;
      if (weight) weight[i] = w;
      if (score) score[i] = prm->batch_score * (double)w;                           // :3798
      if (normal_cent) for (int k = 0; k < 6; ++k) normal_cent[6 * i + k] = nc[k];
    } else {
      if (weight) weight[i] = (st == GO_FAIL_WEIGHT) ? w : 0;
      if (score) score[i] = 0;
      if (normal_cent) for (int k = 0; k < 6; ++k) normal_cent[6 * i + k] = 0;
    }
  }
  delete tree;
  return nvalid;
}

// ---------------------------------------------------------------------------------------------
// Ceres pieces
// ---------------------------------------------------------------------------------------------
void go_huber(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

void go_corrector(double sq_norm, const double rho[3], double out[3]) {
  double sqrt_rho1 = std::sqrt(rho[1]);
  if ((sq_norm == 0.0) || (rho[2] <= 0.0)) { out[0] = sqrt_rho1; out[1] = sqrt_rho1; out[2] = 0.0; return; }
  const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
  const double alpha = 1.0 - std::sqrt(D);
  out[0] = sqrt_rho1; out[1] = sqrt_rho1 / (1 - alpha); out[2] = alpha / sq_norm;
}

void go_quat_plus(const double x[4], const double delta[3], double out[4]) {
  const double nd = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    const double s = std::sin(nd) / nd;
    double z[4] = {std::cos(nd), s * delta[0], s * delta[1], s * delta[2]};
    // ceres::QuaternionProduct(z, x, out)   (ceres.tgz::include/ceres/rotation.h)
    out[0] = z[0] * x[0] - z[1] * x[1] - z[2] * x[2] - z[3] * x[3];
    out[1] = z[0] * x[1] + z[1] * x[0] + z[2] * x[3] - z[3] * x[2];
    out[2] = z[0] * x[2] - z[1] * x[3] + z[2] * x[0] + z[3] * x[1];
    out[3] = z[0] * x[3] + z[1] * x[2] - z[2] * x[1] + z[3] * x[0];
  } else for (int i = 0; i < 4; ++i) out[i] = x[i];
}

void go_quat_plus_jacobian(const double x[4], double J[12]) {
  J[0] = -x[1]; J[1] = -x[2]; J[2] = -x[3];
  J[3] = x[0];  J[4] = x[3];  J[5] = -x[2];
  J[6] = -x[3]; J[7] = x[0];  J[8] = x[1];
  J[9] = x[2];  J[10] = -x[1]; J[11] = x[0];
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Functors, transcribed from LidarKeyframeFactor.h (templated on the scalar like the originals).
// ---------------------------------------------------------------------------------------------
namespace {

template <class T> inline T tsqrt(const T& x) { return jsqrt(x); }

// LidarPlaneNormFactor::operator()  LidarKeyframeFactor.h:88-103
template <class T>
void plane_norm_functor(const double cp_[3], const double nrm_[3], const double qlb_[4], const double tlb_[3],
                        double negOA, double score, const T* t, const T* q, T* residual) {
  T cp[3] = {T(cp_[0]), T(cp_[1]), T(cp_[2])};
  T qlb[4] = {T(qlb_[0]), T(qlb_[1]), T(qlb_[2]), T(qlb_[3])};
  T tlb[3] = {T(tlb_[0]), T(tlb_[1]), T(tlb_[2])};
  T qinvlb[4]; qinv<T>(qlb, qinvlb);
  T d[3] = {cp[0] - tlb[0], cp[1] - tlb[1], cp[2] - tlb[2]};
  T pb[3]; qrot<T>(qinvlb, d, pb);
  T pw[3]; qrot<T>(q, pb, pw);
  pw[0] = pw[0] + t[0]; pw[1] = pw[1] + t[1]; pw[2] = pw[2] + t[2];
  T n[3] = {T(nrm_[0]), T(nrm_[1]), T(nrm_[2])};
  T dot = n[0] * pw[0] + n[1] * pw[1] + n[2] * pw[2];
  residual[0] = T(score) * (dot + T(negOA));
}

// BinaryLidarPlaneNormFactor::operator()  LidarKeyframeFactor.h:131-150
template <class T>
void binary_plane_functor(const double cp_[3], const double pnc_[6], double score, const T* t1, const T* q1,
                          const T* t2, const T* q2, T* residual) {
  T cp[3] = {T(cp_[0]), T(cp_[1]), T(cp_[2])};
  T nl[3] = {T(pnc_[0]), T(pnc_[1]), T(pnc_[2])};
  T cl[3] = {T(pnc_[3]), T(pnc_[4]), T(pnc_[5])};
  T pw[3]; qrot<T>(q1, cp, pw); pw[0] = pw[0] + t1[0]; pw[1] = pw[1] + t1[1]; pw[2] = pw[2] + t1[2];
  T no[3]; qrot<T>(q2, nl, no);
  T co[3]; qrot<T>(q2, cl, co); co[0] = co[0] + t2[0]; co[1] = co[1] + t2[1]; co[2] = co[2] + t2[2];
  T dd[3] = {pw[0] - co[0], pw[1] - co[1], pw[2] - co[2]};
  residual[0] = T(score) * (no[0] * dd[0] + no[1] * dd[1] + no[2] * dd[2]);
}

// LidarEdgeFactor::operator()  LidarKeyframeFactor.h:27-53
template <class T>
void edge_functor(const double cp_[3], const double a_[3], const double b_[3], const double qlb_[4],
                  const double tlb_[3], double s, const T* t, const T* q, T* residual) {
  T cp[3] = {T(cp_[0]), T(cp_[1]), T(cp_[2])};
  T lpa[3] = {T(a_[0]), T(a_[1]), T(a_[2])}, lpb[3] = {T(b_[0]), T(b_[1]), T(b_[2])};
  T qlb[4] = {T(qlb_[0]), T(qlb_[1]), T(qlb_[2]), T(qlb_[3])};
  T tlb[3] = {T(tlb_[0]), T(tlb_[1]), T(tlb_[2])};
  T qi[4]; qinv<T>(qlb, qi);
  T d[3] = {cp[0] - tlb[0], cp[1] - tlb[1], cp[2] - tlb[2]};
  T lp0[3]; qrot<T>(qi, d, lp0);
  T lp[3]; qrot<T>(q, lp0, lp); lp[0] = lp[0] + t[0]; lp[1] = lp[1] + t[1]; lp[2] = lp[2] + t[2];
  T u[3] = {lp[0] - lpa[0], lp[1] - lpa[1], lp[2] - lpa[2]};
  T v[3] = {lp[0] - lpb[0], lp[1] - lpb[1], lp[2] - lpb[2]};
  T nu[3]; cross3<T>(u, v, nu);
  T de[3] = {lpa[0] - lpb[0], lpa[1] - lpb[1], lpa[2] - lpb[2]};
  T nn = tsqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
  T dn = tsqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
  residual[0] = nn / dn;
  residual[0] = residual[0] * T(s);
}

// Ceres residual block post-processing for a scalar residual with nb parameter blocks:
//  jt[b] 1x3 (translation blocks, untouched), jq[b] ambient 1x4 -> tangent 1x3 (jac_kind 0) or x,y,z cols (1).
// then loss / corrector.  Output J is [t(3) | rot(3)] per block.
inline void finish_block(int nb, const double* const* q, double r_raw, const double jt[][3], const double jq[][4],
                         int jac_kind, double huber_delta, double* r_out, double* J_out, double* cost_out) {
  for (int b = 0; b < nb; ++b) {
    double* Jb = J_out + 6 * b;
    Jb[0] = jt[b][0]; Jb[1] = jt[b][1]; Jb[2] = jt[b][2];
    if (jac_kind == 0) {
      double P[12]; go_quat_plus_jacobian(q[b], P);
      // residual_block.cc:150-160  J_local(1x3) = J_global(1x4) * P(4x3)
      for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 4; ++k) s += jq[b][k] * P[3 * k + c]; Jb[3 + c] = s; }
    } else {
      Jb[3] = jq[b][1]; Jb[4] = jq[b][2]; Jb[5] = jq[b][3];   // MarginalizationFactor.cpp:9-12 rightCols(3)
    }
  }
  double sq = r_raw * r_raw, r = r_raw, cost;
  if (huber_delta > 0) {
    double rho[3]; go_huber(huber_delta, sq, rho);
    cost = 0.5 * rho[0];
    double cr[3]; go_corrector(sq, rho, cr);
    // corrector.cc:118-150 for num_rows = 1
    for (int c = 0; c < 6 * nb; ++c) {
      if (cr[2] == 0.0) J_out[c] *= cr[0];
      else { double rtj = J_out[c] * r_raw; J_out[c] = cr[0] * (J_out[c] - cr[2] * r_raw * rtj); }
    }
    r = r_raw * cr[1];
  } else cost = 0.5 * sq;
  *r_out = r; *cost_out = cost;
}

inline void qnorm_rot(const double q[4], const double v[3], double o[3]) { qrot<double>(q, v, o); }

}  // namespace

extern "C" {

void go_eval_unary(int mode, int jac_kind, int W, const double* poses, const double q_lb[4], const double t_lb[3],
                   double huber_delta, int64_t N, const int32_t* kf, const float* cp, const float* nsd,
                   const double* score, double* r_out, double* J_out, double* cost_out, double* H, double* g,
                   double* cost_total) {
  const int n = 6 * W;
  double ctot = 0;
  for (int64_t i = 0; i < N; ++i) {
    const int k = kf[i];
    const double* t = poses + 7 * k; const double* q = poses + 7 * k + 3;
    // Estimator.cpp:2227-2233: floats promoted to double
    double cpd[3] = {(double)cp[3 * i], (double)cp[3 * i + 1], (double)cp[3 * i + 2]};
    double nrm[3] = {(double)nsd[4 * i], (double)nsd[4 * i + 1], (double)nsd[4 * i + 2]};
    double negOA = (double)nsd[4 * i + 3];
    double jt[1][3], jq[1][4], r_raw;
    if (mode == 0) {
      typedef Jet<7> JT;
      JT tj[3] = {JT(t[0], 0), JT(t[1], 1), JT(t[2], 2)};
      JT qj[4] = {JT(q[0], 3), JT(q[1], 4), JT(q[2], 5), JT(q[3], 6)};
      JT res; plane_norm_functor<JT>(cpd, nrm, q_lb, t_lb, negOA, score[i], tj, qj, &res);
      r_raw = res.a; for (int c = 0; c < 3; ++c) jt[0][c] = res.v[c]; for (int c = 0; c < 4; ++c) jq[0][c] = res.v[3 + c];
    } else {
      // closed form: p_b = q_lb^-1 (cp - t_lb); a = R(q) p_b; r = s (n.(a+t) + d)
      double qi[4]; qinv<double>(q_lb, qi);
      double d[3] = {cpd[0] - t_lb[0], cpd[1] - t_lb[1], cpd[2] - t_lb[2]}, pb[3]; qrot<double>(qi, d, pb);
      double a[3]; qrot<double>(q, pb, a);
      r_raw = score[i] * (nrm[0] * (a[0] + t[0]) + nrm[1] * (a[1] + t[1]) + nrm[2] * (a[2] + t[2]) + negOA);
      for (int c = 0; c < 3; ++c) jt[0][c] = score[i] * nrm[c];
      // ambient quaternion Jacobian of n.(R(q)p):  d/dw = 2 n.(u x p);  d/du = 2( (u.p) n + (n.u... ) see below
      const double w = q[0], u[3] = {q[1], q[2], q[3]};
      double uxp[3]; cross3<double>(u, pb, uxp);
      double udp = u[0] * pb[0] + u[1] * pb[1] + u[2] * pb[2];
      double ndu = nrm[0] * u[0] + nrm[1] * u[1] + nrm[2] * u[2];
      double ndp = nrm[0] * pb[0] + nrm[1] * pb[1] + nrm[2] * pb[2];
      double pxn[3]; cross3<double>(pb, nrm, pxn);
      // R(q)p = p + 2w(u x p) + 2 u x (u x p) = p + 2w(u x p) + 2( u (u.p) - p (u.u) )
      // d/dw [n.Rp] = 2 n.(u x p)
      // d/du [n.Rp] = 2w (p x n) + 2( (u.p) n + (n.u) p - 2 (n.p) u )
      jq[0][0] = score[i] * 2.0 * (nrm[0] * uxp[0] + nrm[1] * uxp[1] + nrm[2] * uxp[2]);
      for (int c = 0; c < 3; ++c) jq[0][1 + c] = score[i] * (2.0 * w * pxn[c] + 2.0 * (udp * nrm[c] + ndu * pb[c] - 2.0 * ndp * u[c]));
    }
    const double* qs[1] = {q};
    double r, J[6], c;
    finish_block(1, qs, r_raw, jt, jq, jac_kind, huber_delta, &r, J, &c);
    ctot += c;
    if (r_out) r_out[i] = r;
    if (cost_out) cost_out[i] = c;
    if (J_out) for (int a = 0; a < 6; ++a) J_out[6 * i + a] = J[a];
    if (H) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) H[(size_t)(6 * k + a) * n + 6 * k + b] += J[a] * J[b];
    if (g) for (int a = 0; a < 6; ++a) g[6 * k + a] += J[a] * r;
  }
  if (cost_total) *cost_total += ctot;
}

void go_eval_binary(int mode, int K, const double* poses, double huber_delta, int64_t N, const int32_t* kf_c,
                    const int32_t* kf_o, const float* cp, const double* normal_cent, const double* score,
                    double* r_out, double* J_out, double* cost_out, double* H, double* g, double* cost_total) {
  const int n = 6 * K;
  double ctot = 0;
  for (int64_t i = 0; i < N; ++i) {
    const int kc = kf_c[i], ko = kf_o[i];
    const double* t1 = poses + 7 * kc; const double* q1 = t1 + 3;
    const double* t2 = poses + 7 * ko; const double* q2 = t2 + 3;
    double cpd[3] = {(double)cp[3 * i], (double)cp[3 * i + 1], (double)cp[3 * i + 2]};
    const double* pnc = normal_cent + 6 * i;
    double jt[2][3], jq[2][4], r_raw;
    if (mode == 0) {
      typedef Jet<14> JT;
      JT t1j[3] = {JT(t1[0], 0), JT(t1[1], 1), JT(t1[2], 2)};
      JT q1j[4] = {JT(q1[0], 3), JT(q1[1], 4), JT(q1[2], 5), JT(q1[3], 6)};
      JT t2j[3] = {JT(t2[0], 7), JT(t2[1], 8), JT(t2[2], 9)};
      JT q2j[4] = {JT(q2[0], 10), JT(q2[1], 11), JT(q2[2], 12), JT(q2[3], 13)};
      JT res; binary_plane_functor<JT>(cpd, pnc, score[i], t1j, q1j, t2j, q2j, &res);
      r_raw = res.a;
      for (int c = 0; c < 3; ++c) { jt[0][c] = res.v[c]; jt[1][c] = res.v[7 + c]; }
      for (int c = 0; c < 4; ++c) { jq[0][c] = res.v[3 + c]; jq[1][c] = res.v[10 + c]; }
    } else {
      // closed-form TANGENT Jacobians (SURVEY a-5); expressed as ambient via the minimum-norm lift
      // J_ambient = 0.5 * J_tangent * P^T? -- not needed: we fill the tangent result directly below.
      double pw[3]; qrot<double>(q1, cpd, pw); double a1[3] = {pw[0], pw[1], pw[2]};
      pw[0] += t1[0]; pw[1] += t1[1]; pw[2] += t1[2];
      double Nw[3]; qrot<double>(q2, pnc, Nw);
      double co[3]; qrot<double>(q2, pnc + 3, co); co[0] += t2[0]; co[1] += t2[1]; co[2] += t2[2];
      double dd[3] = {pw[0] - co[0], pw[1] - co[1], pw[2] - co[2]};
      r_raw = score[i] * (Nw[0] * dd[0] + Nw[1] * dd[1] + Nw[2] * dd[2]);
      double c1[3]; cross3<double>(a1, Nw, c1);                       // R(q_c)cp x N
      double pmt[3] = {pw[0] - t2[0], pw[1] - t2[1], pw[2] - t2[2]};
      double c2[3]; cross3<double>(Nw, pmt, c2);                      // N x (p_w - t_o)
      double Jtan[12];
      for (int c = 0; c < 3; ++c) { Jtan[c] = score[i] * Nw[c]; Jtan[3 + c] = 2.0 * score[i] * c1[c]; Jtan[6 + c] = -score[i] * Nw[c]; Jtan[9 + c] = 2.0 * score[i] * c2[c]; }
      // loss + outputs handled here for the closed-form branch
      double sq = r_raw * r_raw, r = r_raw, cst;
      if (huber_delta > 0) { double rho[3]; go_huber(huber_delta, sq, rho); cst = 0.5 * rho[0]; double cr[3]; go_corrector(sq, rho, cr);
        for (int c = 0; c < 12; ++c) { if (cr[2] == 0.0) Jtan[c] *= cr[0]; else { double rtj = Jtan[c] * r_raw; Jtan[c] = cr[0] * (Jtan[c] - cr[2] * r_raw * rtj); } }
        r = r_raw * cr[1]; } else cst = 0.5 * sq;
      ctot += cst;
      if (r_out) r_out[i] = r;
      if (cost_out) cost_out[i] = cst;
      if (J_out) for (int a = 0; a < 12; ++a) J_out[12 * i + a] = Jtan[a];
      const int off[2] = {6 * kc, 6 * ko};
      if (H) for (int bi = 0; bi < 2; ++bi) for (int bj = 0; bj < 2; ++bj) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b)
        H[(size_t)(off[bi] + a) * n + off[bj] + b] += Jtan[6 * bi + a] * Jtan[6 * bj + b];
      if (g) for (int bi = 0; bi < 2; ++bi) for (int a = 0; a < 6; ++a) g[off[bi] + a] += Jtan[6 * bi + a] * r;
      continue;
    }
    const double* qs[2] = {q1, q2};
    double r, J[12], c;
    finish_block(2, qs, r_raw, jt, jq, 0, huber_delta, &r, J, &c);
    ctot += c;
    if (r_out) r_out[i] = r;
    if (cost_out) cost_out[i] = c;
    if (J_out) for (int a = 0; a < 12; ++a) J_out[12 * i + a] = J[a];
    const int off[2] = {6 * kc, 6 * ko};
    if (H) for (int bi = 0; bi < 2; ++bi) for (int bj = 0; bj < 2; ++bj) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b)
      H[(size_t)(off[bi] + a) * n + off[bj] + b] += J[6 * bi + a] * J[6 * bj + b];
    if (g) for (int bi = 0; bi < 2; ++bi) for (int a = 0; a < 6; ++a) g[off[bi] + a] += J[6 * bi + a] * r;
  }
  if (cost_total) *cost_total += ctot;
}

void go_eval_edge(int mode, int W, const double* poses, const double q_lb[4], const double t_lb[3],
                  double huber_delta, int64_t N, const int32_t* kf, const float* cp, const float* pa,
                  const float* pb, const double* s, double* r_out, double* J_out, double* cost_out, double* H,
                  double* g, double* cost_total) {
  const int n = 6 * W;
  double ctot = 0;
  for (int64_t i = 0; i < N; ++i) {
    const int k = kf[i];
    const double* t = poses + 7 * k; const double* q = t + 3;
    double cpd[3] = {(double)cp[3 * i], (double)cp[3 * i + 1], (double)cp[3 * i + 2]};
    double ad[3] = {(double)pa[3 * i], (double)pa[3 * i + 1], (double)pa[3 * i + 2]};
    double bd[3] = {(double)pb[3 * i], (double)pb[3 * i + 1], (double)pb[3 * i + 2]};
    double r, J[6], c;
    if (mode == 0) {
      typedef Jet<7> JT;
      JT tj[3] = {JT(t[0], 0), JT(t[1], 1), JT(t[2], 2)};
      JT qj[4] = {JT(q[0], 3), JT(q[1], 4), JT(q[2], 5), JT(q[3], 6)};
      JT res; edge_functor<JT>(cpd, ad, bd, q_lb, t_lb, s[i], tj, qj, &res);
      double jt[1][3], jq[1][4];
      for (int cc = 0; cc < 3; ++cc) jt[0][cc] = res.v[cc];
      for (int cc = 0; cc < 4; ++cc) jq[0][cc] = res.v[3 + cc];
      const double* qs[1] = {q};
      finish_block(1, qs, res.a, jt, jq, 0, huber_delta, &r, J, &c);
    } else {
      // closed form (SURVEY a-6): e = (lp-a)x(lp-b); gvec = s/|a-b| * ((a-b) x e/|e|); dr/dt = gvec; dr/ddelta = 2 (R p_b x gvec)
      double qi[4]; qinv<double>(q_lb, qi);
      double d[3] = {cpd[0] - t_lb[0], cpd[1] - t_lb[1], cpd[2] - t_lb[2]}, p0[3]; qrot<double>(qi, d, p0);
      double a1[3]; qrot<double>(q, p0, a1);
      double lp[3] = {a1[0] + t[0], a1[1] + t[1], a1[2] + t[2]};
      double u[3] = {lp[0] - ad[0], lp[1] - ad[1], lp[2] - ad[2]}, v[3] = {lp[0] - bd[0], lp[1] - bd[1], lp[2] - bd[2]};
      double e[3]; cross3<double>(u, v, e);
      double de[3] = {ad[0] - bd[0], ad[1] - bd[1], ad[2] - bd[2]};
      double en = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
      double dn = std::sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
      double r_raw = en / dn * s[i];
      double eh[3] = {e[0] / en, e[1] / en, e[2] / en};
      double gv[3]; cross3<double>(de, eh, gv);
      for (int cc = 0; cc < 3; ++cc) gv[cc] *= s[i] / dn;
      double cr[3]; cross3<double>(a1, gv, cr);
      double Jt[6] = {gv[0], gv[1], gv[2], 2.0 * cr[0], 2.0 * cr[1], 2.0 * cr[2]};
      double sq = r_raw * r_raw; r = r_raw;
      if (huber_delta > 0) { double rho[3]; go_huber(huber_delta, sq, rho); c = 0.5 * rho[0]; double co[3]; go_corrector(sq, rho, co);
        for (int cc = 0; cc < 6; ++cc) { if (co[2] == 0.0) Jt[cc] *= co[0]; else { double rtj = Jt[cc] * r_raw; Jt[cc] = co[0] * (Jt[cc] - co[2] * r_raw * rtj); } }
        r = r_raw * co[1]; } else c = 0.5 * sq;
      for (int cc = 0; cc < 6; ++cc) J[cc] = Jt[cc];
    }
    ctot += c;
    if (r_out) r_out[i] = r;
    if (cost_out) cost_out[i] = c;
    if (J_out) for (int a = 0; a < 6; ++a) J_out[6 * i + a] = J[a];
    if (H) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) H[(size_t)(6 * k + a) * n + 6 * k + b] += J[a] * J[b];
    if (g) for (int a = 0; a < 6; ++a) g[6 * k + a] += J[a] * r;
  }
  if (cost_total) *cost_total += ctot;
}

}  // extern "C"
