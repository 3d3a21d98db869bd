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

template <bool WANT_JAC, int JAC_KIND>
__global__ void __launch_bounds__(EV_T) k_eval_unary(const EvalItem* __restrict__ items, int nitems, int W,
                                                     const double* __restrict__ poses, EvalParams ep,
                                                     double* __restrict__ partials, double* __restrict__ out,
                                                     const int* __restrict__ kf_item_start, unsigned int* __restrict__ ticket) {
  __shared__ KfFrame F;
  __shared__ double red[EV_T / 32][NACC];
  __shared__ bool is_last;
  const EvalItem it = items[blockIdx.x];
  if (threadIdx.x == 0) {
    const double* P = poses + 7 * it.kf;
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

  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;

  for (int i = threadIdx.x; i < it.count; i += EV_T) {
    const float4 c4 = __ldg(&it.cpw[i]);
    const float4 n4 = __ldg(&it.nsd[i]);
    const double s = ep.lidar_const * (double)c4.w;            // Estimator.cpp:3692 score = lidar_const*weight
    const double dx = (double)c4.x - ep.t_lb[0], dy = (double)c4.y - ep.t_lb[1], dz = (double)c4.z - ep.t_lb[2];
    double a[3]; mat_vec3(F.M, dx, dy, dz, a);                 // a = R(q) q_lb^-1 (cp - t_lb)
    const double nx = (double)n4.x, ny = (double)n4.y, nz = (double)n4.z;
    double r = s * (nx * (a[0] + F.t[0]) + ny * (a[1] + F.t[1]) + nz * (a[2] + F.t[2]) + (double)n4.w);
    double scale;
    acc[27] += huber_scale(r, ep.huber_delta, scale);
    if (WANT_JAC) {
      double J[6];
      const double ss = s * scale;
      J[0] = ss * nx; J[1] = ss * ny; J[2] = ss * nz;
      if (JAC_KIND == 0) {
        // tangent (Ceres QuaternionParameterization):  2 s (a x n)
        J[3] = 2.0 * ss * (a[1] * nz - a[2] * ny);
        J[4] = 2.0 * ss * (a[2] * nx - a[0] * nz);
        J[5] = 2.0 * ss * (a[0] * ny - a[1] * nx);
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
      double v = 0.0;
      for (int b = kf_item_start[kf]; b < kf_item_start[kf + 1]; ++b) v += partials[(size_t)b * NACC + k];
      out[o] = v;
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

void eval_unary_run(const EvalItem* d_items, int nitems, int W, const double* d_poses, const EvalParams& ep, int jac_kind,
                    bool want_jac, double* d_partials, double* d_out, const int* d_kf_item_start, unsigned int* d_ticket,
                    cudaStream_t st, LaunchCounter& lc) {
  if (nitems <= 0) {
    GLIO_CUDA_TRY(cudaMemsetAsync(d_out, 0, (size_t)W * NACC * sizeof(double), st));
    return;
  }
  lc.begin(want_jac ? "k_eval_unary" : "k_eval_unary_cost", st);
  if (!want_jac) k_eval_unary<false, 0><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, d_poses, ep, d_partials, d_out, d_kf_item_start, d_ticket);
  else if (jac_kind == 0) k_eval_unary<true, 0><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, d_poses, ep, d_partials, d_out, d_kf_item_start, d_ticket);
  else k_eval_unary<true, 1><<<nitems, EV_T, 0, st>>>(d_items, nitems, W, d_poses, ep, d_partials, d_out, d_kf_item_start, d_ticket);
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
  const double s = ep.lidar_const * (double)c4.w;
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

}  // namespace glio
