// eval.cu — K2 / K2b / K2e: residual + Jacobian + Huber + block J^T J / J^T r reduction.
//
// Replaces, for the LiDAR factors, ceres::ResidualBlock::Evaluate (ceres.tgz::internal/ceres/residual_block.cc:70-197:
// autodiff cost -> local parameterization -> Corrector) over LidarPlaneNormFactor / BinaryLidarPlaneNormFactor /
// LidarEdgeFactor (GLIO/include/factors/LidarKeyframeFactor.h:12-164) and the normal-equation accumulation that
// follows it (sparse J^T J in the solve, dense A/b in MarginalizationFactor.cpp:3-29).
//
// HBM traffic per unary residual: 2 x 16 B loads (cp+weight, weight*n + weight*d) = 32 B; outputs are W x 28 doubles.
// Arithmetic is fp64 (tolerance 1e-8 rad per iteration), inputs fp32 as the reference stores them (quirk Q5).
// Reduction: per-thread register accumulators over a run of one keyframe's residuals -> warp shuffle -> shared ->
// one partial per work item; the last block to finish sums the partials per keyframe in a fixed order
// (deterministic, no floating-point atomics).
#include "common.cuh"

namespace glio {

// closed-form rotation matrix of Eigen's Quaternion * v for a (not necessarily unit) quaternion, as the
// polynomial the functor evaluates:  R v = v + 2w(u x v) + 2 u x (u x v)
__device__ __forceinline__ void quat_to_mat(const double q[4], double R[9]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1.0 - 2.0 * (y * y + z * z); R[1] = 2.0 * (x * y - w * z);       R[2] = 2.0 * (x * z + w * y);
  R[3] = 2.0 * (x * y + w * z);       R[4] = 1.0 - 2.0 * (x * x + z * z); R[5] = 2.0 * (y * z - w * x);
  R[6] = 2.0 * (x * z - w * y);       R[7] = 2.0 * (y * z + w * x);       R[8] = 1.0 - 2.0 * (x * x + y * y);
}
__device__ __forceinline__ void mat_mul3(const double A[9], const double B[9], double C[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void mat_vec3(const double A[9], double x, double y, double z, double o[3]) {
  o[0] = A[0] * x + A[1] * y + A[2] * z;
  o[1] = A[3] * x + A[4] * y + A[5] * z;
  o[2] = A[6] * x + A[7] * y + A[8] * z;
}

// Huber + Corrector for a scalar residual (loss_function.cc:48-62, corrector.cc:41-155; rho'' <= 0 always, so the
// simple branch: r <- sqrt(rho') r, J <- sqrt(rho') J).  Returns the block cost 0.5 * rho(s).
__device__ __forceinline__ double huber_scale(double r, double delta, double& scale) {
  const double s = r * r;
  if (delta > 0.0 && s > delta * delta) {
    const double a = sqrt(s);
    double rho1 = delta / a;
    rho1 = rho1 < 2.2250738585072014e-308 ? 2.2250738585072014e-308 : rho1;
    scale = sqrt(rho1);
    return 0.5 * (2.0 * delta * a - delta * delta);
  }
  scale = 1.0;
  return 0.5 * s;
}

constexpr int EV_T = 256;
constexpr int NACC = 28;   // 21 upper-triangular H + 6 g + 1 cost

struct KfFrame {           // per keyframe, prepared in shared memory
  double M[9];             // R(q) * R(q_lb)^-1   (maps cp - t_lb to the rotated body point a)
  double Rlbi[9];          // R(q_lb)^-1
  double t[3];
  double q[4];
};

// per-residual math of the unary plane factor; M, t (and R_lb^-1, q for the marginalisation variant) live in registers
template <bool WANT_JAC, int JAC_KIND>
__device__ __forceinline__ void unary_accumulate(const float4 c4, const float4 n4, const double (&M)[9], const double (&t)[3],
                                                 const KfFrame& F, const EvalParams& ep, double (&acc)[NACC]) {
  const double s = ep.unit_score ? ep.lidar_const : ep.lidar_const * (double)c4.w;   // Estimator.cpp:3692 score = lidar_const*weight; front end: no score
  const double dx = (double)c4.x - ep.t_lb[0], dy = (double)c4.y - ep.t_lb[1], dz = (double)c4.z - ep.t_lb[2];
  double a[3]; mat_vec3(M, dx, dy, dz, a);                   // a = R(q) q_lb^-1 (cp - t_lb)
  const double nx = (double)n4.x, ny = (double)n4.y, nz = (double)n4.z;
  double r = s * (nx * (a[0] + t[0]) + ny * (a[1] + t[1]) + nz * (a[2] + t[2]) + (double)n4.w);
  double scale;
  acc[27] += huber_scale(r, ep.huber_delta, scale);
  if (WANT_JAC) {
    double J[6];
    const double ss = s * scale;
    J[0] = ss * nx; J[1] = ss * ny; J[2] = ss * nz;
    if (JAC_KIND == 0) {
      // tangent (Ceres QuaternionParameterization):  2 s (a x n)
      const double s2 = 2.0 * ss;
      J[3] = s2 * (a[1] * nz - a[2] * ny);
      J[4] = s2 * (a[2] * nx - a[0] * nz);
      J[5] = s2 * (a[0] * ny - a[1] * nx);
    } else {
      // ambient x,y,z columns (MarginalizationFactor.cpp:9-12):
      //   s [ 2w (p_b x n) + 2( (u.p_b) n + (n.u) p_b - 2 (n.p_b) u ) ]
      double pb[3]; mat_vec3(F.Rlbi, dx, dy, dz, pb);
      const double w = F.q[0], ux = F.q[1], uy = F.q[2], uz = F.q[3];
      const double udp = ux * pb[0] + uy * pb[1] + uz * pb[2];
      const double ndu = nx * ux + ny * uy + nz * uz;
      const double ndp = nx * pb[0] + ny * pb[1] + nz * pb[2];
      J[3] = ss * (2.0 * w * (pb[1] * nz - pb[2] * ny) + 2.0 * (udp * nx + ndu * pb[0] - 2.0 * ndp * ux));
      J[4] = ss * (2.0 * w * (pb[2] * nx - pb[0] * nz) + 2.0 * (udp * ny + ndu * pb[1] - 2.0 * ndp * uy));
      J[5] = ss * (2.0 * w * (pb[0] * ny - pb[1] * nx) + 2.0 * (udp * nz + ndu * pb[2] - 2.0 * ndp * uz));
    }
    r *= scale;
    int k = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int c = p; c < 6; ++c) acc[k++] += J[p] * J[c];
#pragma unroll
    for (int p = 0; p < 6; ++p) acc[21 + p] += J[p] * r;
  }
}

// One block per work item (a run of one keyframe's residuals, sized so the whole launch is about one wave of
// 2 blocks/SM).  Loads are issued two iterations ahead of their use (4 x 16 B in flight per thread).
// poses of up to EV_MAXW keyframes travel in the kernel parameters (no host-to-device copy on the per-iteration path)
struct PoseArgs { double p[EV_MAXW * 7]; };

#ifndef GLIO_EVAL_MINBLOCKS
#define GLIO_EVAL_MINBLOCKS 2
#endif
template <bool WANT_JAC, int JAC_KIND, bool POSE_ARGS>
__global__ void __launch_bounds__(EV_T, GLIO_EVAL_MINBLOCKS) k_eval_unary(const EvalItem* __restrict__ items, int nitems, int W,
                                                        const double* __restrict__ poses, const __grid_constant__ PoseArgs pa, EvalParams ep,
                                                        double* __restrict__ partials, double* __restrict__ out,
                                                        const int* __restrict__ kf_item_start, unsigned int* __restrict__ ticket,
                                                        unsigned int* done_flag, unsigned int epoch) {
  __shared__ KfFrame F;
  __shared__ double red[EV_T / 32][NACC];
  __shared__ bool is_last;
  const EvalItem it = items[blockIdx.x];
  if (threadIdx.x == 0) {
    const double* P = POSE_ARGS ? pa.p + 7 * it.kf : poses + 7 * it.kf;
    double qn[4] = {P[3], P[4], P[5], P[6]};
    double R[9]; quat_to_mat(qn, R);
    // q_lb^-1 = conj / |q|^2 (Eigen inverse()); rotation by it as the polynomial in its coefficients
    const double n2 = ep.q_lb[0] * ep.q_lb[0] + ep.q_lb[1] * ep.q_lb[1] + ep.q_lb[2] * ep.q_lb[2] + ep.q_lb[3] * ep.q_lb[3];
    double qi[4] = {ep.q_lb[0] / n2, -ep.q_lb[1] / n2, -ep.q_lb[2] / n2, -ep.q_lb[3] / n2};
    quat_to_mat(qi, F.Rlbi);
    mat_mul3(R, F.Rlbi, F.M);
    F.t[0] = P[0]; F.t[1] = P[1]; F.t[2] = P[2];
    F.q[0] = qn[0]; F.q[1] = qn[1]; F.q[2] = qn[2]; F.q[3] = qn[3];
  }
  __syncthreads();
  double M[9], t[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = F.M[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = F.t[k];

  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;

  const int n = it.count;
  int i = threadIdx.x;
  // software pipeline, depth 2
  float4 c0, n0, c1, n1;
  if (i < n) { c0 = __ldg(&it.cpw[i]); n0 = __ldg(&it.nsd[i]); }
  if (i + EV_T < n) { c1 = __ldg(&it.cpw[i + EV_T]); n1 = __ldg(&it.nsd[i + EV_T]); }
  for (; i < n; i += 2 * EV_T) {
    const float4 ca = c0, na = n0, cb = c1, nb = n1;
    const bool has_b = i + EV_T < n;
    if (i + 2 * EV_T < n) { c0 = __ldg(&it.cpw[i + 2 * EV_T]); n0 = __ldg(&it.nsd[i + 2 * EV_T]); }
    if (i + 3 * EV_T < n) { c1 = __ldg(&it.cpw[i + 3 * EV_T]); n1 = __ldg(&it.nsd[i + 3 * EV_T]); }
    unary_accumulate<WANT_JAC, JAC_KIND>(ca, na, M, t, F, ep, acc);
    if (has_b) unary_accumulate<WANT_JAC, JAC_KIND>(cb, nb, M, t, F, ep, acc);
  }

  // block reduction
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = WANT_JAC ? 0 : 27; k < NACC; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double v = 0.0;
    if (WANT_JAC || threadIdx.x == 27) {
#pragma unroll
      for (int w8 = 0; w8 < EV_T / 32; ++w8) v += red[w8][threadIdx.x];
    }
    partials[(size_t)blockIdx.x * NACC + threadIdx.x] = v;
  }
  // last block: sum the partials per keyframe, in item order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int tk = atomicAdd(ticket, 1u);
    is_last = (tk == (unsigned)nitems - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int o = threadIdx.x; o < W * NACC; o += EV_T) {
      const int kf = o / NACC, k = o - kf * NACC;
      // same left-to-right order as a plain loop, but eight independent L2 loads in flight at a time: this tail is a serial
      // chain of ~15 dependent loads per output otherwise (the last block runs alone on the GPU)
      double v = 0.0;
      int b = kf_item_start[kf];
      const int b1 = kf_item_start[kf + 1];
      for (; b + 8 <= b1; b += 8) {
        double p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) p[u] = __ldcg(&partials[(size_t)(b + u) * NACC + k]);
#pragma unroll
        for (int u = 0; u < 8; ++u) v += p[u];
      }
      for (; b < b1; ++b) v += __ldcg(&partials[(size_t)b * NACC + k]);
      out[o] = v;
    }
    // `out` may be host-mapped pinned memory: publish the W x NACC doubles system-wide, then raise the epoch flag the
    // host spins on (no device-to-host copy, no stream synchronisation on the per-iteration path)
    if (done_flag) { __threadfence_system(); __syncthreads(); }
    if (threadIdx.x == 0) { *ticket = 0u; if (done_flag) *(volatile unsigned int*)done_flag = epoch; }
  }
}

void eval_unary_run(const EvalItem* d_items, int nitems, int W, const double* d_poses, const EvalParams& ep, int jac_kind,
                    bool want_jac, double* d_partials, double* d_out, const int* d_kf_item_start, unsigned int* d_ticket,
                    cudaStream_t st, LaunchCounter& lc, unsigned int* done_flag, unsigned int epoch, const double* h_poses) {
  if (nitems <= 0) {
    GLIO_CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)W * NACC * sizeof(double), st));
    return;
  }
  lc.begin(want_jac ? "k_eval_unary" : "k_eval_unary_cost", st);
  if (h_poses && W <= EV_MAXW) {
    PoseArgs pa;
    memcpy(pa.p, h_poses, (size_t)W * 7 * sizeof(double));
    if (!want_jac) k_eval_unary<false, 0, true><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, nullptr, pa, ep, d_partials, d_out, d_kf_item_start, d_ticket, done_flag, epoch);
    else if (jac_kind == 0) k_eval_unary<true, 0, true><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, nullptr, pa, ep, d_partials, d_out, d_kf_item_start, d_ticket, done_flag, epoch);
    else k_eval_unary<true, 1, true><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, nullptr, pa, ep, d_partials, d_out, d_kf_item_start, d_ticket, done_flag, epoch);
  } else {
    static const PoseArgs none{};
    if (!want_jac) k_eval_unary<false, 0, false><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, d_poses, none, ep, d_partials, d_out, d_kf_item_start, d_ticket, done_flag, epoch);
    else if (jac_kind == 0) k_eval_unary<true, 0, false><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, d_poses, none, ep, d_partials, d_out, d_kf_item_start, d_ticket, done_flag, epoch);
    else k_eval_unary<true, 1, false><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, d_poses, none, ep, d_partials, d_out, d_kf_item_start, d_ticket, done_flag, epoch);
  }
  lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}

// per-residual r, J (the Ceres-API view of one keyframe's residual blocks)
__global__ void __launch_bounds__(256) k_unary_residuals(const float4* __restrict__ cpw, const float4* __restrict__ nsd, int64_t n,
                                                         const double* __restrict__ pose, EvalParams ep, int jac_kind,
                                                         double* __restrict__ r_out, double* __restrict__ J_out) {
  __shared__ KfFrame F;
  if (threadIdx.x == 0) {
    double qn[4] = {pose[3], pose[4], pose[5], pose[6]};
    double R[9]; quat_to_mat(qn, R);
    const double n2 = ep.q_lb[0] * ep.q_lb[0] + ep.q_lb[1] * ep.q_lb[1] + ep.q_lb[2] * ep.q_lb[2] + ep.q_lb[3] * ep.q_lb[3];
    double qi[4] = {ep.q_lb[0] / n2, -ep.q_lb[1] / n2, -ep.q_lb[2] / n2, -ep.q_lb[3] / n2};
    quat_to_mat(qi, F.Rlbi);
    mat_mul3(R, F.Rlbi, F.M);
    F.t[0] = pose[0]; F.t[1] = pose[1]; F.t[2] = pose[2];
    F.q[0] = qn[0]; F.q[1] = qn[1]; F.q[2] = qn[2]; F.q[3] = qn[3];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 c4 = cpw[i];
  const float4 n4 = nsd[i];
  const double s = ep.unit_score ? ep.lidar_const : ep.lidar_const * (double)c4.w;
  const double dx = (double)c4.x - ep.t_lb[0], dy = (double)c4.y - ep.t_lb[1], dz = (double)c4.z - ep.t_lb[2];
  double a[3]; mat_vec3(F.M, dx, dy, dz, a);
  const double nx = (double)n4.x, ny = (double)n4.y, nz = (double)n4.z;
  double r = s * (nx * (a[0] + F.t[0]) + ny * (a[1] + F.t[1]) + nz * (a[2] + F.t[2]) + (double)n4.w);
  double scale;
  huber_scale(r, ep.huber_delta, scale);
  const double ss = s * scale;
  double J[6];
  J[0] = ss * nx; J[1] = ss * ny; J[2] = ss * nz;
  if (jac_kind == 0) {
    J[3] = 2.0 * ss * (a[1] * nz - a[2] * ny);
    J[4] = 2.0 * ss * (a[2] * nx - a[0] * nz);
    J[5] = 2.0 * ss * (a[0] * ny - a[1] * nx);
  } else {
    double pb[3]; mat_vec3(F.Rlbi, dx, dy, dz, pb);
    const double w = F.q[0], ux = F.q[1], uy = F.q[2], uz = F.q[3];
    const double udp = ux * pb[0] + uy * pb[1] + uz * pb[2];
    const double ndu = nx * ux + ny * uy + nz * uz;
    const double ndp = nx * pb[0] + ny * pb[1] + nz * pb[2];
    J[3] = ss * (2.0 * w * (pb[1] * nz - pb[2] * ny) + 2.0 * (udp * nx + ndu * pb[0] - 2.0 * ndp * ux));
    J[4] = ss * (2.0 * w * (pb[2] * nx - pb[0] * nz) + 2.0 * (udp * ny + ndu * pb[1] - 2.0 * ndp * uy));
    J[5] = ss * (2.0 * w * (pb[0] * ny - pb[1] * nx) + 2.0 * (udp * nz + ndu * pb[2] - 2.0 * ndp * uz));
  }
  r_out[i] = r * scale;
#pragma unroll
  for (int k = 0; k < 6; ++k) J_out[6 * i + k] = J[k];
}

void eval_unary_residuals_run(const float4* cpw, const float4* nsd, int64_t n, const double* d_pose, const EvalParams& ep,
                              int jac_kind, double* d_r, double* d_J, cudaStream_t st, LaunchCounter& lc) {
  if (n <= 0) return;
  lc.begin("k_unary_residuals", st); k_unary_residuals<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cpw, nsd, n, d_pose, ep, jac_kind, d_r, d_J); lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}



// ---------------------------------------------------------------------------------------------------------------
// K2e: point-to-edge residuals, LidarEdgeFactor (LidarKeyframeFactor.h:12-70; defined by the reference but never
// instantiated):  lp = R(q) q_lb^-1 (cp - t_lb) + t ;  r = s |(lp-a) x (lp-b)| / |a-b|
// closed form:  e = (lp-a) x (lp-b), gv = s/|a-b| ((a-b) x e/|e|), dr/dt = gv, dr/ddelta = 2 (R p_b x gv).
// Inputs per residual: cp, a, b (float3 each) + s -> 3 x 16 B loads (s rides in cp.w).
// ---------------------------------------------------------------------------------------------------------------
template <bool WANT_JAC>
__global__ void __launch_bounds__(EV_T) k_eval_edge(const EdgeItem* __restrict__ items, int nitems, int W, const double* __restrict__ poses,
                                                    EvalParams ep, double* __restrict__ partials, double* __restrict__ out,
                                                    const int* __restrict__ kf_item_start, unsigned int* __restrict__ ticket) {
  __shared__ KfFrame F;
  __shared__ double red[EV_T / 32][NACC];
  __shared__ bool is_last;
  const EdgeItem it = items[blockIdx.x];
  if (threadIdx.x == 0) {
    const double* P = poses + 7 * it.kf;
    double qn[4] = {P[3], P[4], P[5], P[6]};
    double R[9]; quat_to_mat(qn, R);
    const double n2 = ep.q_lb[0] * ep.q_lb[0] + ep.q_lb[1] * ep.q_lb[1] + ep.q_lb[2] * ep.q_lb[2] + ep.q_lb[3] * ep.q_lb[3];
    double qi[4] = {ep.q_lb[0] / n2, -ep.q_lb[1] / n2, -ep.q_lb[2] / n2, -ep.q_lb[3] / n2};
    quat_to_mat(qi, F.Rlbi);
    mat_mul3(R, F.Rlbi, F.M);
    F.t[0] = P[0]; F.t[1] = P[1]; F.t[2] = P[2];
  }
  __syncthreads();
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  for (int i = threadIdx.x; i < it.count; i += EV_T) {
    const float4 c4 = __ldg(&it.cps[i]);
    const float4 a4 = __ldg(&it.pa[i]);
    const float4 b4 = __ldg(&it.pb[i]);
    const double s = (double)c4.w;
    const double dx = (double)c4.x - ep.t_lb[0], dy = (double)c4.y - ep.t_lb[1], dz = (double)c4.z - ep.t_lb[2];
    double a1[3]; mat_vec3(F.M, dx, dy, dz, a1);
    const double lp[3] = {a1[0] + F.t[0], a1[1] + F.t[1], a1[2] + F.t[2]};
    const double u[3] = {lp[0] - (double)a4.x, lp[1] - (double)a4.y, lp[2] - (double)a4.z};
    const double v[3] = {lp[0] - (double)b4.x, lp[1] - (double)b4.y, lp[2] - (double)b4.z};
    const double e[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    const double de[3] = {(double)a4.x - (double)b4.x, (double)a4.y - (double)b4.y, (double)a4.z - (double)b4.z};
    const double en = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double dn = sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    double r = en / dn * s;
    double scale;
    acc[27] += huber_scale(r, ep.huber_delta, scale);
    if (WANT_JAC) {
      const double k = s / dn * scale / en;
      const double gv[3] = {k * (de[1] * e[2] - de[2] * e[1]), k * (de[2] * e[0] - de[0] * e[2]), k * (de[0] * e[1] - de[1] * e[0])};
      double J[6] = {gv[0], gv[1], gv[2], 2.0 * (a1[1] * gv[2] - a1[2] * gv[1]), 2.0 * (a1[2] * gv[0] - a1[0] * gv[2]), 2.0 * (a1[0] * gv[1] - a1[1] * gv[0])};
      r *= scale;
      int kk = 0;
#pragma unroll
      for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int c = p; c < 6; ++c) acc[kk++] += J[p] * J[c];
#pragma unroll
      for (int p = 0; p < 6; ++p) acc[21 + p] += J[p] * r;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = WANT_JAC ? 0 : 27; k < NACC; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double v = 0.0;
    if (WANT_JAC || threadIdx.x == 27) {
#pragma unroll
      for (int w8 = 0; w8 < EV_T / 32; ++w8) v += red[w8][threadIdx.x];
    }
    partials[(size_t)blockIdx.x * NACC + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { const unsigned int tk = atomicAdd(ticket, 1u); is_last = (tk == (unsigned)nitems - 1); }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int o = threadIdx.x; o < W * NACC; o += EV_T) {
      const int kf = o / NACC, k = o - kf * NACC;
      // same left-to-right order as a plain loop, but eight independent L2 loads in flight at a time: this tail is a serial
      // chain of ~15 dependent loads per output otherwise (the last block runs alone on the GPU)
      double v = 0.0;
      int b = kf_item_start[kf];
      const int b1 = kf_item_start[kf + 1];
      for (; b + 8 <= b1; b += 8) {
        double p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) p[u] = __ldcg(&partials[(size_t)(b + u) * NACC + k]);
#pragma unroll
        for (int u = 0; u < 8; ++u) v += p[u];
      }
      for (; b < b1; ++b) v += __ldcg(&partials[(size_t)b * NACC + k]);
      out[o] = v;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

void eval_edge_run(const EdgeItem* d_items, int nitems, int W, const double* d_poses, const EvalParams& ep, bool want_jac,
                   double* d_partials, double* d_out, const int* d_kf_item_start, unsigned int* d_ticket, cudaStream_t st, LaunchCounter& lc) {
  if (nitems <= 0) { GLIO_CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)W * NACC * sizeof(double), st)); return; }
  lc.begin(want_jac ? "k_eval_edge" : "k_eval_edge_cost", st);
  if (want_jac) k_eval_edge<true><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, d_poses, ep, d_partials, d_out, d_kf_item_start, d_ticket);
  else k_eval_edge<false><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, d_poses, ep, d_partials, d_out, d_kf_item_start, d_ticket);
  lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------
// K2b: binary (scan-to-multiscan) plane factors, BinaryLidarPlaneNormFactor (LidarKeyframeFactor.h:124-164):
//   p_w = R(q_c) cp + t_c ;  N = R(q_o) n_l ;  c_w = R(q_o) c_l + t_o ;  r = s N.(p_w - c_w),  s = batch_score * weight
// tangent Jacobians:  J_c = [u | v],  J_o = [-u | z]  with  u = s N,  v = 2 s (R(q_c)cp x N),  z = 2 s (N x (p_w - t_o)).
// Because J_o's translation part is -u, the 12x12 outer product needs only 45 distinct sums (+9 for J^T r, +1 cost).
// HBM traffic per residual: 16 B (cp, weight) + 48 B (n_l, c_l as the reference stores them: double) = 64 B.
// ---------------------------------------------------------------------------------------------------------------
constexpr int NB = GLIO_NACC_BIN;
constexpr int BIN_T = 128;

struct PairFrame { double R1[9], t1[3], R2[9], t2[3]; };

template <bool WANT_JAC>
__global__ void __launch_bounds__(BIN_T) k_eval_binary(const BinItem* __restrict__ items, const double* __restrict__ poses, double score_scale,
                                                       double huber_delta, double* __restrict__ partials) {
  __shared__ PairFrame F;
  __shared__ double red[BIN_T / 32][NB];
  const BinItem it = items[blockIdx.x];
  if (threadIdx.x == 0) {
    const double* P1 = poses + 7 * it.kf_c; const double* P2 = poses + 7 * it.kf_o;
    double q1[4] = {P1[3], P1[4], P1[5], P1[6]}, q2[4] = {P2[3], P2[4], P2[5], P2[6]};
    quat_to_mat(q1, F.R1); quat_to_mat(q2, F.R2);
    F.t1[0] = P1[0]; F.t1[1] = P1[1]; F.t1[2] = P1[2]; F.t2[0] = P2[0]; F.t2[1] = P2[1]; F.t2[2] = P2[2];
  }
  __syncthreads();
  double acc[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) acc[k] = 0.0;
  for (int i = threadIdx.x; i < it.count; i += BIN_T) {
    const float4 c4 = __ldg(&it.cpw[i]);
    const double2* ncp = reinterpret_cast<const double2*>(it.nc + 6 * (int64_t)i);
    const double2 n01 = __ldg(ncp), n2c0 = __ldg(ncp + 1), c12 = __ldg(ncp + 2);
    const double s = score_scale * (double)c4.w;
    double a1[3]; mat_vec3(F.R1, (double)c4.x, (double)c4.y, (double)c4.z, a1);
    const double pw[3] = {a1[0] + F.t1[0], a1[1] + F.t1[1], a1[2] + F.t1[2]};
    double N[3]; mat_vec3(F.R2, n01.x, n01.y, n2c0.x, N);
    double co[3]; mat_vec3(F.R2, n2c0.y, c12.x, c12.y, co);
    const double pm[3] = {pw[0] - F.t2[0], pw[1] - F.t2[1], pw[2] - F.t2[2]};     // p_w - t_o
    const double dd[3] = {pm[0] - co[0], pm[1] - co[1], pm[2] - co[2]};           // p_w - c_w
    double r = s * (N[0] * dd[0] + N[1] * dd[1] + N[2] * dd[2]);
    double scale;
    acc[54] += huber_scale(r, huber_delta, scale);
    if (WANT_JAC) {
      const double ss = s * scale;
      const double u[3] = {ss * N[0], ss * N[1], ss * N[2]};
      const double v[3] = {2.0 * ss * (a1[1] * N[2] - a1[2] * N[1]), 2.0 * ss * (a1[2] * N[0] - a1[0] * N[2]), 2.0 * ss * (a1[0] * N[1] - a1[1] * N[0])};
      const double z[3] = {2.0 * ss * (N[1] * pm[2] - N[2] * pm[1]), 2.0 * ss * (N[2] * pm[0] - N[0] * pm[2]), 2.0 * ss * (N[0] * pm[1] - N[1] * pm[0])};
      r *= scale;
      int k = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) acc[k++] += u[a] * u[b];          // uu  [0,6)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[k++] += u[a] * v[b];          // uv  [6,15)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) acc[k++] += v[a] * v[b];          // vv  [15,21)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[k++] += u[a] * z[b];          // uz  [21,30)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[k++] += v[a] * z[b];          // vz  [30,39)
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) acc[k++] += z[a] * z[b];          // zz  [39,45)
#pragma unroll
      for (int a = 0; a < 3; ++a) { acc[45 + a] += r * u[a]; acc[48 + a] += r * v[a]; acc[51 + a] += r * z[a]; }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = WANT_JAC ? 0 : 54; k < NB; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NB) {
    double v = 0.0;
    if (WANT_JAC || threadIdx.x == 54) {
#pragma unroll
      for (int w4 = 0; w4 < BIN_T / 32; ++w4) v += red[w4][threadIdx.x];
    }
    partials[(size_t)blockIdx.x * NB + threadIdx.x] = v;
  }
}

// per pair: sum of its items' partials, in item order (deterministic)
__global__ void __launch_bounds__(64) k_bin_pair_sums(const double* __restrict__ partials, const int* __restrict__ pair_item_start,
                                                      double* __restrict__ pair_sums) {
  const int p = blockIdx.x, k = threadIdx.x;
  if (k >= NB) return;
  double v = 0.0;
  for (int b = pair_item_start[p]; b < pair_item_start[p + 1]; ++b) v += partials[(size_t)b * NB + k];
  pair_sums[(size_t)p * NB + k] = v;
}

__device__ __forceinline__ int sym3(int a, int b) { return a <= b ? (a * 3 - a * (a - 1) / 2 + (b - a)) : (b * 3 - b * (b - 1) / 2 + (a - b)); }

// block assembly: blocks [0,K) -> diagonal 6x6 (21 upper-tri) + g (6) + per-KF cost share (cost is attributed to cur);
// blocks [K, K+n_pairs) -> off-diagonal 6x6 (row = cur tangent, col = oth tangent)
__global__ void __launch_bounds__(64) k_bin_assemble(const double* __restrict__ S, int K, int n_pairs, const int* __restrict__ kf_inc_start,
                                                     const BinIncidence* __restrict__ inc, const int* __restrict__ pair_kf /*2 per pair*/,
                                                     double* __restrict__ diag, double* __restrict__ off) {
  const int b = blockIdx.x, e = threadIdx.x;
  if (b < K) {
    if (e >= GLIO_NACC) return;
    double v = 0.0;
    // decode entry e: [0,21) upper-tri (p<=q) of the 6x6, [21,27) g, 27 cost
    int p = 0, q = 0;
    if (e < 21) { int k = e; for (p = 0; p < 6; ++p) { if (k < 6 - p) { q = p + k; break; } k -= 6 - p; } }
    for (int i = kf_inc_start[b]; i < kf_inc_start[b + 1]; ++i) {
      const double* s = S + (size_t)inc[i].pair * NB;
      const int role = inc[i].role;
      if (e < 21) {
        if (p < 3 && q < 3) v += s[sym3(p, q)];                                         // uu (both roles)
        else if (p < 3) v += role == 0 ? s[6 + 3 * p + (q - 3)] : -s[21 + 3 * p + (q - 3)];   // uv | -uz
        else v += role == 0 ? s[15 + sym3(p - 3, q - 3)] : s[39 + sym3(p - 3, q - 3)];  // vv | zz
      } else if (e < 27) {
        const int c = e - 21;
        if (c < 3) v += role == 0 ? s[45 + c] : -s[45 + c];                              // r u | -r u
        else v += role == 0 ? s[48 + (c - 3)] : s[51 + (c - 3)];                         // r v | r z
      } else if (role == 0) v += s[54];
    }
    diag[(size_t)b * GLIO_NACC + e] = v;
  } else {
    const int pr = b - K;
    if (pr >= n_pairs || e >= 36) return;
    const double* s = S + (size_t)pr * NB;
    const int p = e / 6, q = e % 6;
    double v;
    if (p < 3 && q < 3) v = -s[sym3(p, q)];                 // -u u^T
    else if (p < 3) v = s[21 + 3 * p + (q - 3)];            //  u z^T
    else if (q < 3) v = -s[6 + 3 * q + (p - 3)];            // -v u^T
    else v = s[30 + 3 * (p - 3) + (q - 3)];                 //  v z^T
    off[(size_t)pr * 36 + e] = v;
  }
}

void eval_binary_run(const BinItem* d_items, int nitems, const int* d_pair_item_start, int n_pairs, int K, const double* d_poses,
                     double score_scale, double huber_delta, bool want_jac, double* d_partials, double* d_pair_sums,
                     const int* d_kf_inc_start, const BinIncidence* d_inc, double* d_diag, double* d_off, double* d_cost,
                     cudaStream_t st, LaunchCounter& lc) {
  (void)d_cost;
  if (nitems > 0) {
    lc.begin(want_jac ? "k_eval_binary" : "k_eval_binary_cost", st);
    if (want_jac) k_eval_binary<true><<<nitems, BIN_T, 0, st>>>(d_items, d_poses, score_scale, huber_delta, d_partials);
    else k_eval_binary<false><<<nitems, BIN_T, 0, st>>>(d_items, d_poses, score_scale, huber_delta, d_partials);
    lc.end(st);
  }
  if (n_pairs > 0) {
    lc.begin("k_bin_pair_sums", st); k_bin_pair_sums<<<n_pairs, 64, 0, st>>>(d_partials, d_pair_item_start, d_pair_sums); lc.end(st);
  }
  lc.begin("k_bin_assemble", st);
  k_bin_assemble<<<K + (want_jac ? n_pairs : 0), 64, 0, st>>>(d_pair_sums, K, n_pairs, d_kf_inc_start, d_inc, nullptr, d_diag, d_off);
  lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}

}  // namespace glio
