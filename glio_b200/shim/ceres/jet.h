// ceres/jet.h (glio_b200 Ceres-API shim) — forward-mode dual numbers with the arithmetic of ceres::Jet
// (ceres.tgz::include/ceres/jet.h).  Written from the published definition f + sum_i v_i e_i, e_i e_j = 0.
#ifndef GLIO_SHIM_CERES_JET_H_
#define GLIO_SHIM_CERES_JET_H_
#include <cmath>
#include <limits>
#include <ostream>

namespace ceres {

template <typename T, int N>
struct Jet {
  enum { DIMENSION = N };
  typedef T Scalar;
  T a;
  T v[N > 0 ? N : 1];
  Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
  Jet(const T& value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); }  // NOLINT (implicit like ceres)
  Jet(const T& value, int k) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1.0); }
  Jet<T, N>& operator+=(const Jet<T, N>& y) { *this = *this + y; return *this; }
  Jet<T, N>& operator-=(const Jet<T, N>& y) { *this = *this - y; return *this; }
  Jet<T, N>& operator*=(const Jet<T, N>& y) { *this = *this * y; return *this; }
  Jet<T, N>& operator/=(const Jet<T, N>& y) { *this = *this / y; return *this; }
  Jet<T, N>& operator+=(const T& s) { a += s; return *this; }
  Jet<T, N>& operator-=(const T& s) { a -= s; return *this; }
  Jet<T, N>& operator*=(const T& s) { *this = *this * s; return *this; }
  Jet<T, N>& operator/=(const T& s) { *this = *this / s; return *this; }
};

#define GLIO_JET_LOOP for (int i = 0; i < N; ++i)
template <typename T, int N> inline Jet<T, N> const& operator+(const Jet<T, N>& f) { return f; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; GLIO_JET_LOOP h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; GLIO_JET_LOOP h.v[i] = f.v[i] + g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
template <typename T, int N> inline Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a = s + f.a; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; GLIO_JET_LOOP h.v[i] = f.v[i] - g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = s - f.a; GLIO_JET_LOOP h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; GLIO_JET_LOOP h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; GLIO_JET_LOOP h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> inline Jet<T, N> operator*(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = f.a * s; GLIO_JET_LOOP h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; const T gi = T(1.0) / g.a; const T fg = f.a * gi; h.a = fg; GLIO_JET_LOOP h.v[i] = (f.v[i] - fg * g.v[i]) * gi; return h; }
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, T s) { Jet<T, N> h; const T si = T(1.0) / s; h.a = f.a * si; GLIO_JET_LOOP h.v[i] = f.v[i] * si; return h; }
template <typename T, int N> inline Jet<T, N> operator/(T s, const Jet<T, N>& g) { Jet<T, N> h; const T m = -s / (g.a * g.a); h.a = s / g.a; GLIO_JET_LOOP h.v[i] = g.v[i] * m; return h; }

#define GLIO_JET_CMP(op) \
  template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
  template <typename T, int N> inline bool operator op(const T& s, const Jet<T, N>& g) { return s op g.a; } \
  template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const T& s) { return f.a op s; }
GLIO_JET_CMP(<) GLIO_JET_CMP(<=) GLIO_JET_CMP(>) GLIO_JET_CMP(>=) GLIO_JET_CMP(==) GLIO_JET_CMP(!=)
#undef GLIO_JET_CMP

// scalar overloads so templated functors can call ceres::sqrt(T) etc. for T = double as well
inline double abs(double x) { return std::fabs(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline double exp(double x) { return std::exp(x); }
inline double log(double x) { return std::log(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double tan(double x) { return std::tan(x); }
inline double asin(double x) { return std::asin(x); }
inline double acos(double x) { return std::acos(x); }
inline double atan(double x) { return std::atan(x); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double pow(double x, double y) { return std::pow(x, y); }
inline bool IsFinite(double x) { return std::isfinite(x); }
inline bool IsNaN(double x) { return std::isnan(x); }
inline bool IsInfinite(double x) { return std::isinf(x); }

#define GLIO_JET_UNARY(name, val, der) \
  template <typename T, int N> inline Jet<T, N> name(const Jet<T, N>& f) { Jet<T, N> h; h.a = (val); const T d = (der); GLIO_JET_LOOP h.v[i] = d * f.v[i]; return h; }
GLIO_JET_UNARY(abs, std::fabs(f.a), (f.a < T(0.0) ? T(-1.0) : T(1.0)))
GLIO_JET_UNARY(fabs, std::fabs(f.a), (f.a < T(0.0) ? T(-1.0) : T(1.0)))
GLIO_JET_UNARY(sqrt, std::sqrt(f.a), T(1.0) / (T(2.0) * std::sqrt(f.a)))
GLIO_JET_UNARY(exp, std::exp(f.a), std::exp(f.a))
GLIO_JET_UNARY(log, std::log(f.a), T(1.0) / f.a)
GLIO_JET_UNARY(sin, std::sin(f.a), std::cos(f.a))
GLIO_JET_UNARY(cos, std::cos(f.a), -std::sin(f.a))
GLIO_JET_UNARY(tan, std::tan(f.a), T(1.0) + std::tan(f.a) * std::tan(f.a))
GLIO_JET_UNARY(asin, std::asin(f.a), T(1.0) / std::sqrt(T(1.0) - f.a * f.a))
GLIO_JET_UNARY(acos, std::acos(f.a), -T(1.0) / std::sqrt(T(1.0) - f.a * f.a))
GLIO_JET_UNARY(atan, std::atan(f.a), T(1.0) / (T(1.0) + f.a * f.a))
GLIO_JET_UNARY(sinh, std::sinh(f.a), std::cosh(f.a))
GLIO_JET_UNARY(cosh, std::cosh(f.a), std::sinh(f.a))
GLIO_JET_UNARY(tanh, std::tanh(f.a), T(1.0) - std::tanh(f.a) * std::tanh(f.a))
#undef GLIO_JET_UNARY
template <typename T, int N> inline Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
  Jet<T, N> h; const T t = T(1.0) / (f.a * f.a + g.a * g.a); h.a = std::atan2(g.a, f.a); GLIO_JET_LOOP h.v[i] = t * (-g.a * f.v[i] + f.a * g.v[i]); return h; }
template <typename T, int N> inline Jet<T, N> pow(const Jet<T, N>& f, double g) {
  Jet<T, N> h; h.a = std::pow(f.a, g); const T t = g * std::pow(f.a, g - T(1.0)); GLIO_JET_LOOP h.v[i] = t * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> pow(double f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = std::pow(f, g.a); const T t = std::log(f) * h.a; GLIO_JET_LOOP h.v[i] = t * g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> pow(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = std::pow(f.a, g.a); const T t1 = g.a * std::pow(f.a, g.a - T(1.0)), t2 = h.a * std::log(f.a);
  GLIO_JET_LOOP h.v[i] = t1 * f.v[i] + t2 * g.v[i]; return h; }
template <typename T, int N> inline bool IsFinite(const Jet<T, N>& f) { if (!std::isfinite(f.a)) return false; GLIO_JET_LOOP if (!std::isfinite(f.v[i])) return false; return true; }
template <typename T, int N> inline bool IsNaN(const Jet<T, N>& f) { if (std::isnan(f.a)) return true; GLIO_JET_LOOP if (std::isnan(f.v[i])) return true; return false; }
template <typename T, int N> inline bool IsInfinite(const Jet<T, N>& f) { return !IsFinite(f) && !IsNaN(f); }
template <typename T, int N> inline std::ostream& operator<<(std::ostream& s, const Jet<T, N>& z) { s << "[" << z.a << " ; "; GLIO_JET_LOOP s << z.v[i] << (i + 1 < N ? ", " : ""); return s << "]"; }
#undef GLIO_JET_LOOP

}  // namespace ceres

// Eigen interoperability (only when the translation unit already includes Eigen, as the reference's factor headers do)
#ifdef EIGEN_WORLD_VERSION
namespace Eigen {
template <typename T, int N>
struct NumTraits<ceres::Jet<T, N>> {
  typedef ceres::Jet<T, N> Real;
  typedef ceres::Jet<T, N> NonInteger;
  typedef ceres::Jet<T, N> Nested;
  typedef ceres::Jet<T, N> Literal;
  static typename ceres::Jet<T, N> dummy_precision() { return ceres::Jet<T, N>(1e-12); }
  static inline Real epsilon() { return Real(std::numeric_limits<T>::epsilon()); }
  static inline int digits10() { return NumTraits<T>::digits10(); }
  enum { IsComplex = 0, IsInteger = 0, IsSigned, ReadCost = 1, AddCost = 1, MulCost = 3, HasFloatingPoint = 1, RequireInitialization = 1 };
  static inline Real highest() { return Real(std::numeric_limits<T>::max()); }
  static inline Real lowest() { return Real(-std::numeric_limits<T>::max()); }
};
#if EIGEN_VERSION_AT_LEAST(3, 3, 0)
template <typename BinaryOp, typename T, int N> struct ScalarBinaryOpTraits<ceres::Jet<T, N>, T, BinaryOp> { typedef ceres::Jet<T, N> ReturnType; };
template <typename BinaryOp, typename T, int N> struct ScalarBinaryOpTraits<T, ceres::Jet<T, N>, BinaryOp> { typedef ceres::Jet<T, N> ReturnType; };
#endif
}  // namespace Eigen
#endif

#endif  // GLIO_SHIM_CERES_JET_H_
