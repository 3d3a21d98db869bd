// hostmath.h — host-side pose algebra of the boundary (double, Eigen 3.3 operation order restated).
#pragma once
#include <cmath>

namespace glio {

inline void h_cross3(const double a[3], const double b[3], double o[3]) {
  const double o0 = a[1] * b[2] - a[2] * b[1], o1 = a[2] * b[0] - a[0] * b[2], o2 = a[0] * b[1] - a[1] * b[0];
  o[0] = o0; o[1] = o1; o[2] = o2;
}
// Eigen QuaternionBase::_transformVector
inline void h_qrot(const double q[4], const double v[3], double o[3]) {
  const double u[3] = {q[1], q[2], q[3]};
  double uv[3]; h_cross3(u, v, uv);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3]; h_cross3(u, uv, c);
  const double r0 = v[0] + q[0] * uv[0] + c[0], r1 = v[1] + q[0] * uv[1] + c[1], r2 = v[2] + q[0] * uv[2] + c[2];
  o[0] = r0; o[1] = r1; o[2] = r2;
}
inline void h_qmul(const double a[4], const double b[4], double o[4]) {
  const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  const double y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  const double z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
inline void h_qinv(const double q[4], double o[4]) {
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  o[0] = q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = -q[3] / n2;
}

// Q2 = Q * q_lb^-1 ; T2 = T - Q2 * t_lb     (GLIO/src/Estimator.cpp:2216-2217)
inline void lidar_pose_in_map(const double q_lb[4], const double t_lb[3], const double pose_body[7], double t2[3], double q2[4]) {
  double qi[4]; h_qinv(q_lb, qi);
  double qq[4]; h_qmul(pose_body + 3, qi, qq);
  double r[3]; h_qrot(qq, t_lb, r);
  t2[0] = pose_body[0] - r[0]; t2[1] = pose_body[1] - r[1]; t2[2] = pose_body[2] - r[2];
  q2[0] = qq[0]; q2[1] = qq[1]; q2[2] = qq[2]; q2[3] = qq[3];
}

}  // namespace glio
