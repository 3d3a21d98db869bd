/*
 * glio_oracle.h — CPU restatement ("oracle") of GLIO's per-scan LiDAR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under glio_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker / the reported CPU baseline.
 *
 * PARITY STATUS: the reference (XikunLiu-huskit/GLIO @ 332d19ff) ships no tests, golden vectors or fixtures for this
 * path, and it cannot be compiled in this image (no Eigen / Ceres install / PCL / ROS), so there is no oracle/_ref.
 * The oracle is authored from the reference *source* and from the published algorithms of its absent dependencies
 * (Eigen 3.3.4 ColPivHouseholderQR / Quaternion, FLANN 1.9 L2_Simple<float> KDTreeSingleIndex, PCL VoxelGrid, Ceres
 * 2.0.0 whose source IS vendored as a tarball).  PINNED (tests/test_oracle_pins.py, test_oracle_known_answers.py,
 * the C++ tests under tests/cpp): the kNN bit for bit to OpenCV's bundled FLANN (KDTREE_SINGLE and LINEAR, also at M = 1 M), the 5x3 plane
 * solve to LAPACK's pivoted QR, the Ceres pieces (corrector, loss, quaternion parameterization, dogleg, Levenberg-Marquardt,
 * Powell, polynomial roots) to Ceres' own known-answer tests.  PARITY UNPINNED for what no library in the image can check:
 * pcl::VoxelGrid, Eigen's packet-order reductions, std::sort's tie order (VoxelGrid, feature extraction).
 *
 * Every function cites the reference file:line it follows.  Paths are relative to
 * /root/reference; "ceres.tgz::" means support_files/ceres-solver.tar.gz → ceres-solver/.
 *
 * Floating point: compiled with -ffp-contract=off so that no FMA contraction happens,
 * matching the reference's baseline x86-64 (SSE2) build (GLIO/CMakeLists.txt:5, -O3, no -march).
 */
#ifndef GLIO_ORACLE_H
#define GLIO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Parameters of the association (GLIO/config/config_urban_hk.yaml:60-104, Estimator.cpp:70,852-877). */
typedef struct {
  double kd_max_radius;   /* 1.5; compared against a SQUARED distance (quirk Q1, Estimator.cpp:3651) */
  double surf_dist_thres; /* 0.18  (Estimator.cpp:3671) */
  double lidar_const;     /* 7.5   (Estimator.cpp:3690) */
  double weight_min;      /* 0.3   (Estimator.cpp:3681) */
  double batch_max_radius;  /* 1.5 hard-coded (Estimator.cpp:3751) */
  double batch_dist_thres;  /* 0.18 hard-coded (Estimator.cpp:3778) */
  double batch_score;       /* 2.5 hard-coded (Estimator.cpp:3798) */
} go_assoc_params;

/* status codes per query */
enum { GO_VALID = 0, GO_FAIL_RADIUS = 1, GO_FAIL_PLANE = 2, GO_FAIL_WEIGHT = 3 };

/* p_out = float(q * double(p_in) + t)    (Estimator.cpp:1490-1498, Eigen Quaternion::_transformVector) */
void go_transform_points(const float* in_xyz, int64_t n, const double t[3], const double q_wxyz[4],
                         float* out_xyz);

/* exact 5-NN, brute force, FLANN L2_Simple<float> accumulation order, ties broken by index.
 * tie[i] = 1 if the 5th/6th (or any adjacent pair among the first 6) distances are equal. */
/* Local map down-sampling: pcl::VoxelGrid<PointXYZI>::applyFilter as GLIO uses it (ds_filter_surf_map, leaf 0.4 m,
 * Estimator.cpp:854, :3617-3618; downsample_all_data = true, min_points_per_voxel = 0), xyz only.
 * PCL is a dependency that is NOT in /root/reference (ROS's libpcl 1.8/1.10): this restates its published algorithm
 * (filters/include/pcl/filters/impl/voxel_grid.hpp applyFilter: float min/max, min_b = floor(min * inv_leaf),
 * idx = ijk0 + ijk1*div0 + ijk2*div0*div1 with ijk = int(floor(p*inv_leaf) - float(min_b)), std::sort by idx, per-voxel
 * float sums divided by the float count, output in ascending voxel index).  order_mode 0: literal std::sort (the order of
 * points inside a voxel - hence the last bits of the float sums - is whatever introsort leaves); order_mode 1: stable
 * (points of a voxel summed in input order), the variant the CUDA path reproduces bit-for-bit.
 * Returns the number of output points (<= n); out_xyz needs room for n points; out_idx (optional) receives each output
 * point's voxel index.  If the grid would overflow int32 PCL copies the input unchanged: returns -1. */
int64_t go_voxel_filter(const float* xyz, int64_t n, float leaf, int order_mode, float* out_xyz, int32_t* out_idx);

void go_knn5_brute(const float* map_xyz, int64_t M, const float* qry_xyz, int64_t Q,
                   int32_t* idx5, float* sqd5, uint8_t* tie);

/* kd-tree (KDTreeSingleIndex-style: leaf size 15, reordered points, exact search) */
void* go_kdtree_build(const float* xyz, int64_t M);
void go_kdtree_free(void* tree);
void go_kdtree_knn5(const void* tree, const float* qry_xyz, int64_t Q, int32_t* idx5, float* sqd5);

/* x = colPivHouseholderQr(A).solve(-1)   for a 5x3 row-major A (Estimator.cpp:3649-3661).
 * returns the number of non-zero pivots. */
int go_plane_solve5(const double A[15], double x[3]);

/* Scan-to-map association (Estimator.cpp:3633-3708).  kdtree may be NULL (brute force).
 * Per-query outputs (all length Q unless noted; any may be NULL):
 *   status[Q], idx5[5Q], sqd5[5Q], pm[3Q] (transformed query, float), plane[4Q] (unit n, d; double),
 *   nsd[4Q] (float: weight*n, weight*d), weight[Q] (float), score[Q] (double = lidar_const*weight).
 * returns number of valid matches. */
int64_t go_assoc_scan_to_map(const go_assoc_params* prm, const float* map_xyz, int64_t M, const void* kdtree,
                             const float* scan_xyz, int64_t Q, const double t[3], const double q[4],
                             uint8_t* status, int32_t* idx5, float* sqd5, float* pm, double* plane,
                             float* nsd, float* weight, double* score, int nthreads);

/* Scan-to-multiscan pair association (Estimator.cpp:3710-3806 / 3808-3892): frame "cur" against frame "oth".
 * Poses are the body poses applied directly to lidar-frame points (quirk Q7).
 * Outputs per cur point: status, idx5, sqd5, weight (float), score (=batch_score*weight, double),
 * normal_cent[6Q] (double: local-frame unit normal, local-frame centroid). */
int64_t go_assoc_pair(const go_assoc_params* prm, const float* cur_xyz, int64_t Qc, const double t_c[3],
                      const double q_c[4], const float* oth_xyz, int64_t Qo, const double t_o[3],
                      const double q_o[4], int use_kdtree, uint8_t* status, int32_t* idx5, float* sqd5,
                      float* weight, double* score, double* normal_cent, int nthreads);

/* ---- factor evaluation (LidarKeyframeFactor.h:12-164 through Ceres' ResidualBlock::Evaluate,
 *      ceres.tgz::internal/ceres/residual_block.cc:70-197) ----
 * mode: 0 = Jet autodiff of the functor exactly as written, ambient J x QuaternionParameterization J,
 *       1 = closed-form tangent Jacobians.
 * jac_kind: 0 = Ceres tangent (solve path), 1 = ambient x,y,z quaternion columns (marginalisation path,
 *           MarginalizationFactor.cpp:9-17, quirk Q12).
 * huber_delta <= 0 : no loss function.
 * Per-residual outputs r[N], J[6N or 12N], cost[N] may be NULL.  H is a dense (6W x 6W) row-major matrix,
 * g is 6W; both are ACCUMULATED INTO (caller zeroes).  cost_total is accumulated into as well. */
void go_eval_unary(int mode, int jac_kind, int W, const double* poses /*W*7: t, q(wxyz)*/,
                   const double q_lb[4], const double t_lb[3], double huber_delta, int64_t N,
                   const int32_t* kf, const float* cp /*3N*/, const float* nsd /*4N*/, const double* score,
                   double* r, double* J, double* cost, double* H, double* g, double* cost_total);

void go_eval_binary(int mode, int K, const double* poses /*K*7*/, double huber_delta, int64_t N,
                    const int32_t* kf_c, const int32_t* kf_o, const float* cp /*3N*/,
                    const double* normal_cent /*6N*/, const double* score, double* r, double* J /*12N*/,
                    double* cost, double* H /*6K x 6K*/, double* g, double* cost_total);

void go_eval_edge(int mode, int W, const double* poses, const double q_lb[4], const double t_lb[3],
                  double huber_delta, int64_t N, const int32_t* kf, const float* cp, const float* pa,
                  const float* pb, const double* s, double* r, double* J, double* cost, double* H,
                  double* g, double* cost_total);

/* Ceres pieces exposed for the known-answer tests */
void go_huber(double a, double s, double rho[3]);                          /* loss_function.cc:48-62 */
void go_corrector(double sq_norm, const double rho[3], double out[3]);     /* corrector.cc:41-110: sqrt_rho1, residual_scaling, alpha_sq_norm */
void go_quat_plus(const double x[4], const double delta[3], double out[4]);/* local_parameterization.cc:163-182 */
void go_quat_plus_jacobian(const double x[4], double J12[12]);             /* local_parameterization.cc:184-191 */

#ifdef __cplusplus
}
#endif
#endif
