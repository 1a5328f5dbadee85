// oracle/dual.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Forward-mode dual numbers standing in for ForwardDiff.Dual as the reference uses it
// (RayTracer.jl:10 `using ForwardDiff: gradient, derivative, Dual`, seeds at RT.jl:21,
// extraction at RT.jl:24).  ForwardDiff is an un-vendored third-party dependency of the
// reference (no Project.toml ⇒ version unpinned); its semantics on this path are exact
// forward-mode derivatives of +,-,*,/,sqrt,sin,cos,exp,abs,^, which is what is restated here.
#pragma once
#include <cmath>

namespace orc {

template <int N>
struct Dual {
  double v;
  double d[N];
  Dual() : v(0.0) {
    for (int i = 0; i < N; ++i) d[i] = 0.0;
  }
  Dual(double x) : v(x) {  // NOLINT: implicit on purpose (constants promote like in Julia)
    for (int i = 0; i < N; ++i) d[i] = 0.0;
  }
  static Dual seed(double x, int k) {
    Dual r(x);
    r.d[k] = 1.0;
    return r;
  }
};

inline double val(double x) { return x; }
template <int N>
inline double val(const Dual<N>& x) { return x.v; }

template <int N>
inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v + b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N>
inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v - b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N>
inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v * b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N>
inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v / b.v;
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
template <int N>
inline Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r;
  r.v = -a.v;
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}

// mixed with double
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { return b + a; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { return (-b) + a; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) {
  Dual<N> r;
  r.v = a.v * b;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
  return r;
}
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) {
  Dual<N> r;
  r.v = a.v / b;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b;
  return r;
}
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }

template <int N> inline Dual<N> sqrt(const Dual<N>& a) {
  Dual<N> r;
  r.v = std::sqrt(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / (2.0 * r.v);
  return r;
}
template <int N> inline Dual<N> sin(const Dual<N>& a) {
  Dual<N> r;
  r.v = std::sin(a.v);
  double c = std::cos(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c;
  return r;
}
template <int N> inline Dual<N> cos(const Dual<N>& a) {
  Dual<N> r;
  r.v = std::cos(a.v);
  double s = -std::sin(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s;
  return r;
}
template <int N> inline Dual<N> exp(const Dual<N>& a) {
  Dual<N> r;
  r.v = std::exp(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * r.v;
  return r;
}
// ForwardDiff: abs(d) = signbit(value(d)) ? -d : d
template <int N> inline Dual<N> abs(const Dual<N>& a) { return std::signbit(a.v) ? -a : a; }
// x^p for real exponent p
template <int N> inline Dual<N> pow(const Dual<N>& a, double p) {
  Dual<N> r;
  r.v = std::pow(a.v, p);
  double g = p * std::pow(a.v, p - 1.0);
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * g;
  return r;
}

// plain-double overloads so templated physics code reads the same for both
inline double sqrt(double a) { return std::sqrt(a); }
inline double sin(double a) { return std::sin(a); }
inline double cos(double a) { return std::cos(a); }
inline double exp(double a) { return std::exp(a); }
inline double abs(double a) { return std::fabs(a); }
inline double pow(double a, double p) { return std::pow(a, p); }

template <class T> inline T sq(const T& a) { return a * a; }
template <class T> inline T cube(const T& a) { return a * a * a; }

}  // namespace orc
