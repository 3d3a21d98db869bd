// marg.h — the marginalisation prior produced by glio_window_marginalize and consumed by the next window's solve /
// marginalisation (MarginalizationInfo + MarginalizationFactor, GLIO/src/MarginalizationFactor.cpp:128-330).
//
// Index conventions (W keyframes, states per keyframe: t(3), q(4, tangent / ambient-x,y,z 3), speed_bias(9)):
//   marginalisation ordering of a window, N = 6W + 18:   KF0: t 0, q 3, sb 6 (the m = 15 states to drop, first)
//                                                         KF1: t 15, q 18, sb 21;   KF k >= 2: t 30 + 6(k-2), q 33 + 6(k-2)
//     (the reference orders its blocks by an unordered_map of addresses, MarginalizationFactor.cpp:129-145; any order gives
//      the same prior up to a permutation, and its consumer indexes through keep_block_idx)
//   prior ordering (what survives, n = 6W + 3), expressed in the NEXT window's numbering (addr_shift, Estimator.cpp:2583-2597):
//                                                         KF0: t 0, q 3, sb 6;   KF k >= 1 (k <= W-2): t 15 + 6(k-1), q 18 + 6(k-1)
#pragma once
#include <vector>

struct glio_marg_prior {
  int W = 0;                       // window size the prior was built from (= the window size it applies to)
  int n = 0;                       // 6W + 3
  std::vector<double> lin_jac;     // n x n row-major   linearized_jacobians = diag(sqrt S) V^T        (:198-199)
  std::vector<double> lin_res;     // n                 linearized_residuals = diag(sqrt(1/S)) V^T b'  (:200-201)
  std::vector<double> A_info;      // n x n             lin_jac^T lin_jac  (= the Schur complement restricted to eigenvalues > eps)
  std::vector<double> b_info;      // n                 lin_jac^T lin_res
  double c0 = 0;                   // 0.5 |lin_res|^2
  std::vector<double> x0_pose;     // (W-1) x 7         keep_block_data: the kept poses at linearisation, next-window numbering
  double x0_sb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // speed/bias of the kept keyframe (next window's KF0)
  int half_bandwidth = 0;          // of A_info in the SOLVE ordering (15 tangent dims per keyframe): what the band solve must cover
};

namespace glio {
inline int marg_index_t(int k) { return k == 0 ? 0 : (k == 1 ? 15 : 30 + 6 * (k - 2)); }       // marginalisation ordering
inline int prior_index_t(int k) { return k == 0 ? 0 : 15 + 6 * (k - 1); }                       // prior ordering (next window's KF k)
// Schur complement + eigen-decomposition (marginalize.cpp): A (N x N), b (N), first m states dropped
int marginalize_dense(const double* A, const double* b, int n_total, int m, double eps, double* lin_jac, double* lin_res);
}  // namespace glio
