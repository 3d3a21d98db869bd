/*
 * glio_b200.h — C ABI of the B200-native LiDAR hot path of GLIO.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / Eigen / PCL / Ceres / ROS types.
 * Every entry point names the reference interface it replaces (paths relative to the GLIO repository
 * root, reference commit 332d19ff).  The Ceres-compatible C++ shim (glio_b200/shim/ceres/ headers) and the
 * Python mirror (glio_b200/api.py) are thin layers over exactly these functions.
 *
 * Conventions
 *   - points: float32 xyz with a caller-given stride in floats (3 = packed xyz, 8 = pcl::PointXYZI as
 *     stored in pcl::PointCloud<PointType>::points, GLIO/include/utils/common.h:87-89).
 *   - poses: double[7] = t(x,y,z), q(w,x,y,z) — the reference's tmpTrans[k] / tmpQuat[k] parameter
 *     blocks (GLIO/src/Estimator.cpp:345-347).
 *   - every function returns 0 on success or a negative glio_status; glio_last_error() gives the text.
 *     No exception crosses the boundary and nothing aborts.
 *   - mem: GLIO_HOST pointers are copied for the duration of the call only; GLIO_DEVICE pointers are
 *     device addresses on the context's GPU (used by bench "value" runs with inputs resident in HBM).
 *   - a glio_ctx is single-threaded; create one per concurrent caller (window solve / batch solve run on
 *     different threads in the reference, Estimator.cpp:5352-5368).  Each context owns one CUDA stream.
 *   - there is NO CPU fallback: without a CUDA device glio_create fails with GLIO_ERR_NO_DEVICE.
 */
#ifndef GLIO_B200_H
#define GLIO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct glio_ctx glio_ctx;

typedef enum {
  GLIO_OK = 0,
  GLIO_ERR_NO_DEVICE = -1,
  GLIO_ERR_CUDA = -2,
  GLIO_ERR_ARG = -3,
  GLIO_ERR_STATE = -4,
  GLIO_ERR_NCCL = -5,
  GLIO_ERR_NUMERIC = -6
} glio_status;

enum { GLIO_HOST = 0, GLIO_DEVICE = 1 };

/* per-query association status (mirrors the reference's three failure counters, Estimator.cpp:3641-3643) */
enum { GLIO_MATCH_VALID = 0, GLIO_MATCH_FAIL_RADIUS = 1, GLIO_MATCH_FAIL_PLANE = 2, GLIO_MATCH_FAIL_WEIGHT = 3 };

/* All tunables of the path in one POD (SURVEY.md §5 "Config / flags"). */
typedef struct {
  double kd_max_radius;    /* Estimator/kd_max_radius 1.5 — compared with a SQUARED distance (Estimator.cpp:3651) */
  double surf_dist_thres;  /* Estimator/surf_dist_thres 0.18 (Estimator.cpp:3671) */
  double lidar_const;      /* Estimator/lidar_const 7.5 (Estimator.cpp:3690) */
  double weight_min;       /* 0.3 hard-coded (Estimator.cpp:3681) */
  double huber_delta;      /* lossKernel 1.0 (Estimator.cpp:70,2092); <= 0 disables the loss */
  double q_lb[4];          /* extrinsic, wxyz (Estimator.cpp:270-274,873-877) */
  double t_lb[3];
  double batch_max_radius; /* 1.5 hard-coded (Estimator.cpp:3751) */
  double batch_dist_thres; /* 0.18 hard-coded (Estimator.cpp:3778) */
  double batch_score;      /* 2.5 hard-coded (Estimator.cpp:3798) */
  float cell_size;         /* uniform-grid cell edge in metres; 0 = choose from point density */
  int32_t keep_debug;      /* 1: keep idx5/sqd5/plane/pm per query for glio_get_assoc_debug (parity tests) */
  int32_t unit_score;      /* 0: residual score = lidar_const*weight (LidarPlaneNormFactor, Estimator.cpp:3690);
                            * 1: score = lidar_const for every match - with lidar_const = 1, identity extrinsic, kd_max_radius 1.0,
                            *    surf_dist_thres 0.06, weight_min 0.4, huber_delta 0.1 this is the front end's scan matcher
                            *    (LidarOdometry.cpp:343-404, :499-521; LidarPlaneNormIncreFactor, LidarKeyframeFactor.h:222-257) */
} glio_params;

void glio_default_params(glio_params* p);

/* lifecycle */
int glio_create(int device, const glio_params* params, glio_ctx** out);
void glio_destroy(glio_ctx* ctx);
const char* glio_last_error(const glio_ctx* ctx);   /* ctx may be NULL: last global error */
int glio_synchronize(glio_ctx* ctx);
/* the context's CUDA stream (cudaStream_t as void*) so callers can time with events on it */
void* glio_stream(glio_ctx* ctx);
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
int64_t glio_launch_count(const glio_ctx* ctx);

/* per-kernel device time: when enabled every launch is bracketed by CUDA events on the context's stream;
 * glio_profile_get sums the elapsed time of all launches of the named kernel since profiling was enabled
 * (names: K0 k_init_bounds, k_load_bounds, k_cell_hist, k_cell_scatter, k_make_pairs, k_scan_block, k_scan_add;
 * K1 k_transform_hist, k_order_scatter, k_knn_box [GLIO_KNN_MODE 2, default] | k_knn_thread [1] | k_knn_search + k_knn_deferred [0]
 * | k_knn_tile + k_defer_scatter + k_knn_tile2 + k_knn_team [4] | k_knn_box_start + k_knn_grow [5] | k_knn_box_far + k_knn_far [6]
 * | k_knn_box_cells [7], k_plane_fit, k_plane_fit_pair, k_flags, k_compact;
 * K2 k_eval_unary, k_eval_unary_cost, k_eval_binary, k_eval_binary_cost, k_bin_assemble; local map k_lm_transform, k_vox_*;
 * front end k_feat_curvature, k_feat_select, k_feat_voxel, k_feat_offsets, k_feat_gather). */
int glio_profile_enable(glio_ctx* ctx, int on);
int glio_profile_get(glio_ctx* ctx, const char* kernel_name, double* ms_total, int64_t* launches);
/* statistics: number of queries a first search pass handed to its second pass (GLIO_KNN_MODE 0 and 4; 0 in the default mode) */
int glio_get_stats(glio_ctx* ctx, int64_t* knn_fallback_queries, int reset);
/* vec_surf_res_cnt for slots 0..W-1 (all matches) and the number currently active (after glio_select) */
int glio_get_match_counts(glio_ctx* ctx, int W, int64_t* n_match, int64_t* n_active);

/* lidar->map pose of a keyframe:  Q2 = Q * q_lb^-1 ; T2 = T - Q2 * t_lb   (Estimator.cpp:2216-2217) */
void glio_lidar_pose(const glio_params* prm, const double pose_body[7], double t2[3], double q2[4]);

/* ---- K0: local-map upload + uniform-grid build.
 * Replaces kd_tree_surf_local_map->setInputCloud(surf_local_map_ds)  (Estimator.cpp:2056). */
int glio_set_map(glio_ctx* ctx, const float* xyz, int64_t M, int stride_floats, int mem);
/* start the upload of the NEXT map from (pinned) host memory on the copy stream and return; the following glio_set_map with the same
 * pointer / count / stride uses the staged copy.  The poses a map is built from are final when the solve returns, so a caller can
 * hand the rebuilt map over while the window is still being marginalised. */
int glio_map_prefetch(glio_ctx* ctx, const float* xyz, int64_t M, int stride_floats);

/* ---- Local map maintenance on the device.
 * Replaces Estimator::buildLocalMapWithLandMark (Estimator.cpp:3529-3610: the deque `recent_surf_keyframes` of
 * world-frame keyframe clouds and their concatenation `surf_local_map`) and the map half of downSampleCloud
 * (:3615-3618: ds_filter_surf_map, pcl::VoxelGrid leaf 0.4 m set at :854) followed by
 * kd_tree_surf_local_map->setInputCloud (:2056).  The keyframe clouds stay resident on the GPU; only a new keyframe's
 * cloud crosses PCIe.
 *   glio_localmap_push       recent_surf_keyframes.push_front / push_back(transformCloud(cloud, {t, q}))   (:3574, :3600)
 *                            with t = q_po*t_bl + t_po, q = q_po*q_bl formed by the caller as at :3563-3564
 *   glio_localmap_pop_front  recent_surf_keyframes.pop_front()                                               (:3580)
 *   glio_localmap_clear      recent_surf_keyframes.clear()                                                   (:3549)
 *   glio_localmap_build      concatenate (:3605-3608), VoxelGrid(leaf) (:3617-3618), make it the searchable map; leaf <= 0
 *                            skips the filter.  PCL sums the points of a voxel in the order std::sort leaves them; this
 *                            sums them in input order (the deterministic member of that family).
 *   glio_get_map             the points of the current map in their original order (filtered map: ascending voxel index). */
int glio_localmap_clear(glio_ctx* ctx);
int glio_localmap_push(glio_ctx* ctx, int at_front, const float* cloud_xyz, int64_t n, int stride_floats, int mem,
                       const double t[3], const double q[4]);
int glio_localmap_pop_front(glio_ctx* ctx);
int glio_localmap_size(glio_ctx* ctx, int* n_frames, int64_t* n_points);
int glio_localmap_build(glio_ctx* ctx, float leaf, int64_t* n_map);
int glio_get_map(glio_ctx* ctx, int64_t capacity, float* xyz, int64_t* n_map);

/* ---- K1: scan-to-map surf association for ONE keyframe slot.
 * Replaces Estimator::findCorrespondingSurfFeatures(idx, q, t)  (Estimator.cpp:3633-3708):
 * q,t is the lidar->map pose (Q2,T2 of Estimator.cpp:2216-2217).  The scan stays resident in the slot.
 * *n_match receives vec_surf_res_cnt[idVec].  Matches are kept on the device in scan order (the order of the
 * reference's push_back), as (point, weight*normal, weight*d, weight). */
int glio_assoc_scan_to_map(glio_ctx* ctx, int slot, const float* scan_xyz, int64_t Q, int stride_floats, int mem,
                           const double t[3], const double q[4], int64_t* n_match);

/* One keyframe's scan into one slot (host buffers travel on the copy stream; the next association waits for them), and the
 * slide of the window by one keyframe (slideWindow in the reference: slot k+1 -> slot k, matches included; no copies).  A
 * sliding window hands over ONE new scan per keyframe: the other W-1 stay resident. */
int glio_window_set_scan(glio_ctx* ctx, int slot, const float* scan_xyz, int64_t Q, int stride_floats, int mem);
int glio_window_slide(glio_ctx* ctx, int W);

/* All W keyframes of the window in one launch (same results as W calls of glio_assoc_scan_to_map).
 * poses_body[W*7] are the keyframe (IMU-body) poses tmpTrans/tmpQuat; the lidar->map pose is formed on the
 * device-side host code exactly as Estimator.cpp:2216-2217 with params.q_lb/t_lb.
 * glio_window_set_scans with GLIO_HOST buffers is ASYNCHRONOUS: the copies run on the context's copy stream and the next
 * association waits for them, so the host buffers must stay valid until that association (or glio_synchronize) returns.
 * It does not touch the matches of the previous association: the scans of the NEXT window may be handed over while the
 * current window is still being solved (upload overlaps the solve). */
int glio_window_set_scans(glio_ctx* ctx, int W, const float* const* scans, const int64_t* Q, int stride_floats, int mem);
int glio_window_associate(glio_ctx* ctx, int W, const double* poses_body, int64_t* n_match /*W*/);

/* direct upload of a slot's match list when the association was done elsewhere (e.g. by an unmodified Estimator through
 * PCL): cp[3n], nsd[4n] = (weight*n, weight*d), weight[n]; score = lidar_const*weight.  Used by the Ceres shim. */
int glio_set_matches(glio_ctx* ctx, int slot, const float* cp, const float* nsd, const float* weight, int64_t n);
int glio_get_params(const glio_ctx* ctx, glio_params* out);

/* copy-out of one slot's matches (parity tests, and the Ceres shim's host view).  Any pointer may be NULL.
 *   cp[3n] float (scan-frame point), nsd[4n] float (weight*n, weight*d), weight[n] float, src[n] int32
 *   (index of the match's scan point).  score = lidar_const * (double)weight. */
int glio_get_matches(glio_ctx* ctx, int slot, int64_t capacity, float* cp, float* nsd, float* weight, int32_t* src,
                     int64_t* n_match);
/* per-query debug view (needs params.keep_debug): status[Q] u8, idx5[5Q], sqd5[5Q], pm[3Q], plane[4Q] double */
int glio_get_assoc_debug(glio_ctx* ctx, int slot, int64_t Q, uint8_t* status, int32_t* idx5, float* sqd5, float* pm,
                         double* plane);

/* ---- feature selection as an INPUT (Estimator.cpp:3894-3992 is an RNG sub-sampling).
 * keep[n] are indices into the slot's match list; n < 0 clears the selection (all matches active). */
int glio_select(glio_ctx* ctx, int slot, const int32_t* keep, int64_t n);

/* ---- K2: evaluate all active unary plane residuals at the given body poses.
 * Replaces ceres::ResidualBlock::Evaluate over LidarPlaneNormFactor (LidarKeyframeFactor.h:73-122) + HuberLoss +
 * QuaternionParameterization and the J^T J / J^T r accumulation for those blocks.
 *   jac_kind 0: Ceres tangent Jacobian (solve path); 1: ambient x,y,z quaternion columns
 *   (marginalisation path, MarginalizationFactor.cpp:9-17).
 *   H[W*36] row-major 6x6 per keyframe (t | rotation), g[W*6] = J^T r, cost[W] = sum 0.5*rho(r^2).
 *   H and g may be NULL (cost-only evaluation at a candidate point). */
int glio_eval_unary(glio_ctx* ctx, int W, const double* poses_body, int jac_kind, double* H, double* g, double* cost);

/* per-residual view for the Ceres-API path: r[n], J[6n] (corrected, tangent) for one slot */
int glio_eval_unary_residuals(glio_ctx* ctx, int slot, const double pose_body[7], int jac_kind, int64_t capacity,
                              double* r, double* J, int64_t* n);

/* ---- the minimizer iteration around K2 (replaces ceres::Solve for the window problem, Estimator.cpp:2424-2433):
 * Ceres 2.0.0 TrustRegionMinimizer + DoglegStrategy + Jacobi scaling + normal-equation Cholesky semantics
 * (ceres.tgz::internal/ceres/{trust_region_minimizer,dogleg_strategy,trust_region_step_evaluator}.cc).
 * LiDAR blocks come from the device; all other factors of the problem (IMU, marginalisation prior, GNSS — host C++
 * by design) are added by the host_factors callback into the same dense tangent-space normal equations. */
typedef struct {
  int32_t max_num_iterations;            /* 15 (Estimator.cpp:2427); batch: max_num_iter */
  int32_t dogleg_type;                   /* 0 TRADITIONAL_DOGLEG (window), 1 SUBSPACE_DOGLEG (batch, Estimator.cpp:3280) */
  int32_t use_nonmonotonic_steps;        /* 0 window (Estimator.cpp:2430) / 1 batch (:3281) */
  int32_t max_consecutive_nonmonotonic_steps; /* 5  (ceres solver.h:263) */
  double initial_trust_region_radius;    /* 1e4  (solver.h:276) */
  double max_trust_region_radius;        /* 1e16 */
  double min_trust_region_radius;        /* 1e-32 */
  double min_relative_decrease;          /* 1e-3 */
  double min_lm_diagonal;                /* 1e-6 */
  double max_lm_diagonal;                /* 1e32 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t jacobi_scaling;                /* 1 */
  double function_tolerance;             /* 1e-6 */
  double gradient_tolerance;             /* 1e-10 */
  double parameter_tolerance;            /* 1e-8 */
  int32_t fuse_candidate_jacobian;       /* 1: evaluate J with the candidate cost and reuse it on acceptance (same numbers, one pass less) */
  int32_t trust_region_strategy;         /* 0 DOGLEG (Estimator.cpp:2427, :3278), 1 LEVENBERG_MARQUARDT (Ceres' default: the front end's scan
                                            matcher, LidarOdometry.cpp:521-530; ceres.tgz::internal/ceres/levenberg_marquardt_strategy.cc) */
} glio_solver_options;
void glio_default_solver_options(glio_solver_options* o);

typedef struct {
  int32_t iteration, step_is_valid, step_is_successful, reserved;
  double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease, trust_region_radius, mu;
} glio_iteration;

typedef struct {
  int32_t termination;                   /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE */
  int32_t num_iterations;                /* records written to the iteration log (iteration 0 included) */
  int32_t num_successful_steps, num_unsuccessful_steps;
  int32_t num_evaluations, num_jacobian_evaluations, num_linear_solves, num_valid_steps;
  double initial_cost, final_cost;
  double eval_seconds;                   /* wall clock inside residual/Jacobian evaluation (device kernel + copies + host factors) */
  double linear_solver_seconds;          /* wall clock inside the band Cholesky solves */
  double total_seconds;
  char message[128];
} glio_solver_summary;

/* Adds the host-side factors at the given state.  n = W*(speed_bias ? 15 : 6), tangent order per keyframe
 * [t(3), rotation(3), speed_bias(9)].  H (n*n row-major, full symmetric) and g (n) are ACCUMULATED INTO when
 * want_jac != 0 (they already hold the LiDAR blocks); *cost is accumulated into.  Return 0 on success. */
typedef int (*glio_host_factors_fn)(void* user, int W, const double* poses, const double* speed_bias, int want_jac,
                                    double* H, double* g, double* cost);

/* poses[W*7] (and speed_bias[W*9], may be NULL) are in/out: on return the lowest-cost accepted state, like the
 * parameter blocks after ceres::Solve.  iter_log[iter_cap] and step_log[step_cap doubles: one n-vector per valid
 * step, the tangent update Delta_k of SURVEY B.4] may be NULL. */
int glio_window_solve(glio_ctx* ctx, int W, double* poses, double* speed_bias, glio_host_factors_fn host_factors, void* user,
                      const glio_solver_options* options, glio_solver_summary* summary, glio_iteration* iter_log, int iter_cap,
                      double* step_log, int64_t step_cap);

/* Same solve with the host factors accumulating straight into LOWER-BAND storage (entry (i,j), i-hb <= j <= i, at
 * Hband[i*(hb+1) + (j-i+hb)]): no dense n x n scratch.  half_bandwidth must cover every coupling the host factors create
 * (window: prior + IMU chain -> 2*nt-1 = 29 with speed-bias states); < 5 means dense. */
typedef int (*glio_host_factors_band_fn)(void* user, int K, const double* poses, const double* speed_bias, int want_jac,
                                         double* Hband, int hb, double* g, double* cost);
int glio_window_solve_band(glio_ctx* ctx, int W, double* poses, double* speed_bias, glio_host_factors_band_fn host_factors,
                           int half_bandwidth, void* user, const glio_solver_options* options, glio_solver_summary* summary,
                           glio_iteration* iter_log, int iter_cap, double* step_log, int64_t step_cap);

/* ---- stand-in host factors (CPU C++, analytic) with the block structure of the reference's non-LiDAR factors:
 * prior 15x[t,q,sb] (marginalisation-prior-like), between 15x[t,q,sb|t,q,sb] (IMU-chain-like), range 1x[t,q]
 * (pseudorange-like).  glio_hf_evaluate has the glio_host_factors_fn signature (user = the set).  They exist so the
 * plumbing of host factors + device LiDAR blocks can be tested and benchmarked without ROS / GNSS data; the real
 * ImuFactor / MarginalizationFactor / dd_psr_factor plug in through the same callback (or the Ceres shim). */
typedef struct glio_host_factor_set glio_host_factor_set;
glio_host_factor_set* glio_hf_create(void);
void glio_hf_destroy(glio_host_factor_set* s);
void glio_hf_add_prior(glio_host_factor_set* s, int kf, const double t0[3], const double q0[4], const double* sb0 /*9 or NULL*/,
                       const double sqrt_w[15]);
void glio_hf_add_between(glio_host_factor_set* s, int i, int j, const double dp[3], const double dq[4], const double dv[3], double dt,
                         const double sqrt_w[15]);
void glio_hf_add_range(glio_host_factor_set* s, int kf, const double lever[3], const double sat[3], double rho, double w);
int glio_hf_evaluate(void* user, int W, const double* poses, const double* speed_bias, int want_jac, double* H, double* g, double* cost);

/* ---- front-end feature extraction: Preprocessing::cloudHandler, GLIO/src/Preprocessing.cpp:529-655.
 * cloud_xyzi = `laserCloud` of the reference: the scan lines concatenated ring after ring, x,y,z,intensity at stride_floats
 * (>= 4) with the intensity at float intensity_offset (packed x,y,z,i: stride 4, offset 3; pcl::PointXYZI: stride 8, offset 4),
 * scan_start / scan_end = scanStartInd / scanEndInd per ring (:529-534; every range inside
 * [5, n-6]).  Outputs, any may be NULL: cloudCurvature[n], cloudLabel[n] (2 sharp, 1 less sharp, -1 flat, 0 other), and the
 * four feature sets in the reference's push_back order as indices into the cloud (cornerPointsSharp, cornerPointsLessSharp,
 * surfPointsFlat, the less-flat points BEFORE the voxel filter) plus surfPointsLessFlat itself (x,y,z,intensity after the
 * per-ring pcl::VoxelGrid of leaf ds_v).  Equal curvatures are ordered by index (std::sort leaves it open). */
int glio_extract_features(glio_ctx* ctx, const float* cloud_xyzi, int64_t n, int stride_floats, int intensity_offset, int mem, int n_scans, const int32_t* scan_start,
                          const int32_t* scan_end, int ds_rate, double edge_thres, double surf_thres, float ds_v, float* curvature, int8_t* label,
                          int32_t* sharp, int64_t* n_sharp, int32_t* less_sharp, int64_t* n_less_sharp, int32_t* flat, int64_t* n_flat,
                          int32_t* less_flat, int64_t* n_less_flat, float* less_flat_ds, int64_t* n_less_flat_ds);

/* ---- K3: marginalisation of the oldest keyframe (MarginalizationInfo, GLIO/src/MarginalizationFactor.cpp:82-202; call site
 * Estimator.cpp:2462-2608).  The reference re-evaluates EVERY LiDAR factor of the window (:2538-2576, one virtual Evaluate per
 * residual) plus the IMU factor KF0->KF1 and the previous prior, accumulates the dense A = J^T J, b = J^T r with the ambient
 * x,y,z quaternion columns (:9-17), removes KF0's 15 states by a Schur complement and factors the rest by an
 * eigen-decomposition (:176-201).  glio_window_marginalize does the LiDAR part on the device (K2 with jac_kind = 1: one 6x6
 * block + 6-vector per keyframe), asks the host callback for the other factors, and does the small dense algebra on the host.
 * Orderings (see glio_b200/csrc/marg.h): marginalisation ordering N = 6W+18 = [KF0: t,q,sb | KF1: t,q,sb | KF k>=2: t,q];
 * the prior keeps n = 6W+3 states, numbered as the NEXT window numbers them (addr_shift, Estimator.cpp:2583-2597).
 * host_marg(user, W, poses, speed_bias, A[N*N], b[N]) ACCUMULATES the non-LiDAR factors (may be NULL). */
typedef int (*glio_host_marg_fn)(void* user, int W, const double* poses, const double* speed_bias, double* A, double* b);
typedef struct glio_marg_prior glio_marg_prior;
int glio_window_marginalize(glio_ctx* ctx, int W, const double* poses, const double* speed_bias, glio_host_marg_fn host_marg, void* user,
                            double eps /* 1e-8 in the reference */, glio_marg_prior** prior_out);
/* The same pass with its host half (A/b assembly, Schur complement, decomposition, prior: ~0.15 ms) on a worker thread of the
 * context: the call returns when the device half is queued and host_marg has run, so the caller can hand the NEXT window's map and
 * association to the GPU meanwhile (the reference does these strictly one after the other, Estimator.cpp:2462-2608 then the next
 * processImage).  One job at a time per context.  glio_marg_job_wait joins the job, frees it and returns the prior (identical to
 * glio_window_marginalize's); call it before the next solve needs the prior.  glio_destroy joins AND frees an abandoned job: a job
 * handle must not be used after its context is gone.  A context is otherwise single-threaded: calls on one context must not
 * overlap, the worker is the library's own. */
typedef struct glio_marg_job glio_marg_job;
int glio_window_marginalize_async(glio_ctx* ctx, int W, const double* poses, const double* speed_bias, glio_host_marg_fn host_marg,
                                  void* user, double eps, glio_marg_job** job_out);
int glio_marg_job_wait(glio_marg_job* job, glio_marg_prior** prior_out);
glio_marg_prior* glio_marg_prior_create(int W, const double* lin_jac, const double* lin_res, const double* x0_poses, const double* x0_sb);
void glio_marg_prior_destroy(glio_marg_prior* p);
int glio_marg_prior_size(const glio_marg_prior* p, int* n, int* W);
/* linearized_jacobians[n*n] row-major, linearized_residuals[n], keep_block_data (x0_poses[(W-1)*7], x0_sb[9]), and the
 * information form A_info = J^T J [n*n], b_info = J^T r [n]; any pointer may be NULL */
int glio_marg_prior_get(const glio_marg_prior* p, double* lin_jac, double* lin_res, double* x0_poses, double* x0_sb, double* A_info, double* b_info);
/* stand-in host factors: attach the prior to the next window's factor set (it then takes part in glio_hf_evaluate[_band] as a
 * MarginalizationFactor, MarginalizationFactor.cpp:232-330) and glio_hf_marg_evaluate = the glio_host_marg_fn of the set
 * (previous prior + priors on KF0 + the between factor KF0->KF1). */
int glio_hf_set_marg_prior(glio_host_factor_set* s, const glio_marg_prior* p /* copied; NULL removes */);
int glio_hf_marg_evaluate(void* user, int W, const double* poses, const double* speed_bias, double* A, double* b);
int glio_hf_marg_half_bandwidth(const glio_host_factor_set* s);

/* ---- K1b: scan-to-multiscan (batch) association.
 * Replaces findGlobalCorrespondingSurfFeatures_Batch / ...Add_Batch (Estimator.cpp:3710-3892) and their driver
 * batchFeatureAssociation (:3413-3432).  Frames (one keyframe's surf scan, stored as received + its pose from
 * pose_info_keyframe, quirks Q7/Q8) are registered once; a pair (cur, oth) associates every point of frame `cur`
 * against the world cloud of frame `oth`.  The uniform grid of a frame is built once and reused by every pair
 * that searches it (the reference rebuilds a kd-tree per pair, Estimator.cpp:3729-3731: same neighbours). */
int glio_batch_set_frame(glio_ctx* ctx, int frame, const float* scan_xyz, int64_t Q, int stride_floats, int mem,
                         const double pose[7]);
int glio_batch_set_pose(glio_ctx* ctx, int frame, const double pose[7]);
/* associate `cur` against each of oth[n_oth]; n_match[n_oth] receives gl_vec_surf_res_cnt[cur][oth]. */
int glio_batch_associate(glio_ctx* ctx, int cur, const int32_t* oth, int n_oth, int64_t* n_match);
/* many pairs at once (grouped by searched frame internally): pairs_cur[n], pairs_oth[n] */
int glio_batch_associate_pairs(glio_ctx* ctx, const int32_t* pairs_cur, const int32_t* pairs_oth, int64_t n_pairs,
                               int64_t* n_match);
/* matches of one pair: cp[3n] float (point of cur, local), weight[n] float (score = batch_score*weight),
 * normal_cent[6n] double (unit normal and centroid of the 5 neighbours in oth's LOCAL frame), src[n] */
int glio_batch_get_matches(glio_ctx* ctx, int cur, int oth, int64_t capacity, float* cp, float* weight,
                           double* normal_cent, int32_t* src, int64_t* n_match);
/* globalFeatureSelection_Batch (Estimator.cpp:3994-4116) as an input index list; n < 0 clears */
int glio_batch_select(glio_ctx* ctx, int cur, int oth, const int32_t* keep, int64_t n);
int glio_batch_pair_list(glio_ctx* ctx, int64_t capacity, int32_t* cur, int32_t* oth, int64_t* n_pairs);
int glio_batch_clear(glio_ctx* ctx);

/* Creates (empty) pair entries in the given order without associating them: on a keyframe-sharded multi-GPU run
 * every rank declares the SAME global pair list and then associates only the pairs it owns, so the block
 * buffers have one layout everywhere. */
int glio_batch_declare_pairs(glio_ctx* ctx, const int32_t* pairs_cur, const int32_t* pairs_oth, int64_t n_pairs);

/* Upload the matches of pair (cur, oth) instead of associating them (creates the pair when it does not exist yet) - the
 * path the Ceres-API shim takes for BinaryLidarPlaneNormFactor residual blocks, which carry their own data
 * (LidarKeyframeFactor.h:161-163: curr_point, planet_norm_cent, score = batch_score*weight). */
int glio_batch_set_pair_matches(glio_ctx* ctx, int cur, int oth, const float* cp, const double* normal_cent, const float* weight, int64_t n);

/* ---- K2b: evaluate all active binary plane residuals at poses[K*7].
 * Replaces ResidualBlock::Evaluate over BinaryLidarPlaneNormFactor (LidarKeyframeFactor.h:124-164; no loss function,
 * Estimator.cpp:2768) and the J^T J / J^T r accumulation.  Block-sparse output: Hdiag[K*36], Hoff[n_pairs*36]
 * (block (cur,oth): row = cur tangent, col = oth tangent, pairs in glio_batch_pair_list order), g[K*6], cost. */
int glio_eval_binary(glio_ctx* ctx, int K, const double* poses, double* Hdiag, double* Hoff, double* g, double* cost);

/* ---- K3: the marginalisation Schur step of MarginalizationInfo::Marginalize (GLIO/src/MarginalizationFactor.cpp:176-201),
 * host C++ (123 x 15 x 123 at W = 20: too small for the GPU to matter).  A[n_total x n_total] row-major and b[n_total]
 * are the dense normal equations with the m states to drop first (LiDAR part: glio_eval_unary with jac_kind = 1);
 * outputs linearized_jacobians[(n_total-m)^2] row-major and linearized_residuals[n_total-m].  eps = 1e-8 in the reference. */
int glio_marginalize(const double* A, const double* b, int n_total, int m, double eps, double* lin_jac, double* lin_res);

/* ---- K2e: point-to-edge residuals (LidarEdgeFactor, LidarKeyframeFactor.h:12-70).  The reference defines this factor
 * but never instantiates it and has no edge association, so the correspondences are an input:
 * cp[3n] scan point, pa/pb[3n] two points of the map line, s[n] weight.  Huber(huber_delta) as for the plane factors.
 * glio_eval_edge returns the same per-keyframe 6x6 / 6 / cost blocks as glio_eval_unary (add them to the plane blocks). */
int glio_set_edges(glio_ctx* ctx, int slot, const float* cp, const float* pa, const float* pb, const double* s, int64_t n);
int glio_eval_edge(glio_ctx* ctx, int W, const double* poses_body, double* H, double* g, double* cost);

/* ---- batch minimizer (replaces ceres::Solve of optimizeBatchWithLandMark, Estimator.cpp:3275-3284: SUBSPACE_DOGLEG,
 * nonmonotonic steps, max_num_iter iterations; options == NULL selects exactly those).  The normal equations are
 * block-banded (half bandwidth (max|cur-oth|+1)*nt - 1) and stored/factored in band form.  Host factors (IMU chain,
 * delta_q, DD pseudorange: host C++ by design) accumulate into the same LOWER-BAND storage:
 * entry (i,j), i-hb <= j <= i, lives at Hband[i*(hb+1) + (j-i+hb)]. */
int glio_hf_evaluate_band(void* user, int K, const double* poses, const double* speed_bias, int want_jac, double* Hband, int hb,
                          double* g, double* cost);
int glio_batch_solve(glio_ctx* ctx, int K, double* poses, double* speed_bias, glio_host_factors_band_fn host_factors, void* user,
                     const glio_solver_options* options, glio_solver_summary* summary, glio_iteration* iter_log, int iter_cap,
                     double* step_log, int64_t step_cap);

/* ---- multi-GPU (batch path only; the window path is "replicas only").
 * Keyframe-sharded ranks sum their pose-block buffers once per evaluation.  The library does not link NCCL: the
 * caller installs the reduction (in-place sum over ranks of count doubles at device pointer d_buf, enqueued on
 * `cuda_stream`).  glio_b200/libglio_nccl.so provides glio_nccl_allreduce for an ncclComm_t; Python callers may
 * install a torch.distributed-based hook instead. */
typedef int (*glio_allreduce_fn)(void* user, double* d_buf, int64_t count, void* cuda_stream);
int glio_set_allreduce(glio_ctx* ctx, glio_allreduce_fn fn, void* user);

#ifdef __cplusplus
}
#endif
#endif
