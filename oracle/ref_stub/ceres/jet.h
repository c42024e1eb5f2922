// ceres/jet.h — STUB (test infrastructure, oracle/_ref).  Forward-mode dual numbers with the arithmetic of Ceres 1.14's
// include/ceres/jet.h restated from its documentation (the real Ceres is absent from this image; docker/Dockerfile:46
// installs libceres-dev of Ubuntu 20.04 = 1.14.0).  The formulas that are visible in results are kept exactly:
//   f / g  = (f.a * (1/g.a),  (f.v - (f.a * (1/g.a)) * g.v) * (1/g.a))
//   s / g  = (s / g.a, g.v * (-s / (g.a * g.a))),      f / s = (f.a * (1/s), f.v * (1/s))
//   sqrt f = (t = sqrt(f.a), f.v * (1 / (2 t))),  asin f = (asin f.a, f.v / sqrt(1 - f.a^2)),
//   atan2(g, f) = (atan2(g.a, f.a), (f.a * g.v - g.a * f.v) / (f.a^2 + g.a^2)),  floor f = (floor f.a, 0).
// Comparisons look at the scalar part only.
#pragma once
#include <cmath>
#include <iostream>
#include <limits>

namespace ceres {

template <typename T, int N>
struct Jet {
    enum { DIMENSION = N };
    T a;
    T v[N];
    Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
    Jet(const T& value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); }   // NOLINT: implicit, as Ceres
    Jet(const T& value, int k) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1.0); }
    Jet& operator+=(const Jet& y) { *this = *this + y; return *this; }
    Jet& operator-=(const Jet& y) { *this = *this - y; return *this; }
    Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
    Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
    Jet& operator+=(const T& s) { *this = *this + s; return *this; }
    Jet& operator-=(const T& s) { *this = *this - s; return *this; }
    Jet& operator*=(const T& s) { *this = *this * s; return *this; }
    Jet& operator/=(const T& s) { *this = *this / s; return *this; }
};

template <typename T, int N> inline Jet<T, N> const& operator+(const Jet<T, N>& f) { return f; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
template <typename T, int N> inline Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a = f.a + s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> inline Jet<T, N> operator*(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
    const T g_a_inverse = T(1.0) / g.a;
    const T f_a_by_g_a = f.a * g_a_inverse;
    Jet<T, N> h;
    h.a = f.a * g_a_inverse;
    for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
    return h;
}
template <typename T, int N> inline Jet<T, N> operator/(T s, const Jet<T, N>& g) {
    const T minus_s_g_a_inverse2 = -s / (g.a * g.a);
    Jet<T, N> h;
    h.a = s / g.a;
    for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2;
    return h;
}
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, T s) {
    const T s_inverse = T(1.0) / s;
    Jet<T, N> h;
    h.a = f.a * s_inverse;
    for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s_inverse;
    return h;
}

#define STUB_JET_CMP(op)                                                                                         \
    template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
    template <typename T, int N> inline bool operator op(const T& s, const Jet<T, N>& g) { return s op g.a; }           \
    template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const T& s) { return f.a op s; }
STUB_JET_CMP(<)
STUB_JET_CMP(<=)
STUB_JET_CMP(>)
STUB_JET_CMP(>=)
STUB_JET_CMP(==)
STUB_JET_CMP(!=)
#undef STUB_JET_CMP

// scalar versions live in namespace ceres too (ceres::floor(double), ceres::sqrt(double), ...)
using std::abs;
using std::acos;
using std::asin;
using std::atan;
using std::atan2;
using std::cos;
using std::cosh;
using std::exp;
using std::floor;
using std::isfinite;
using std::isinf;
using std::isnan;
using std::log;
using std::pow;
using std::sin;
using std::sinh;
using std::sqrt;
using std::tan;
using std::tanh;

template <typename T, int N> inline Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
template <typename T, int N> inline Jet<T, N> log(const Jet<T, N>& f) { const T a_inverse = T(1.0) / f.a; Jet<T, N> h; h.a = log(f.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * a_inverse; return h; }
template <typename T, int N> inline Jet<T, N> exp(const Jet<T, N>& f) { const T tmp = exp(f.a); Jet<T, N> h; h.a = tmp; for (int i = 0; i < N; ++i) h.v[i] = tmp * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& f) {
    const T tmp = sqrt(f.a);
    const T two_a_inverse = T(1.0) / (T(2.0) * tmp);
    Jet<T, N> h;
    h.a = tmp;
    for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
    return h;
}
template <typename T, int N> inline Jet<T, N> cos(const Jet<T, N>& f) { const T s = -sin(f.a); Jet<T, N> h; h.a = cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> sin(const Jet<T, N>& f) { const T c = cos(f.a); Jet<T, N> h; h.a = sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> tan(const Jet<T, N>& f) { const T t = tan(f.a); const T tmp = T(1.0) + t * t; Jet<T, N> h; h.a = t; for (int i = 0; i < N; ++i) h.v[i] = tmp * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> acos(const Jet<T, N>& f) { const T tmp = -T(1.0) / sqrt(T(1.0) - f.a * f.a); Jet<T, N> h; h.a = acos(f.a); for (int i = 0; i < N; ++i) h.v[i] = tmp * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> asin(const Jet<T, N>& f) { const T tmp = T(1.0) / sqrt(T(1.0) - f.a * f.a); Jet<T, N> h; h.a = asin(f.a); for (int i = 0; i < N; ++i) h.v[i] = tmp * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> atan(const Jet<T, N>& f) { const T tmp = T(1.0) / (T(1.0) + f.a * f.a); Jet<T, N> h; h.a = atan(f.a); for (int i = 0; i < N; ++i) h.v[i] = tmp * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
    const T tmp = T(1.0) / (f.a * f.a + g.a * g.a);
    Jet<T, N> h;
    h.a = atan2(g.a, f.a);
    for (int i = 0; i < N; ++i) h.v[i] = tmp * (-g.a * f.v[i] + f.a * g.v[i]);
    return h;
}
template <typename T, int N> inline Jet<T, N> floor(const Jet<T, N>& f) { return Jet<T, N>(floor(f.a)); }
template <typename T, int N> inline Jet<T, N> pow(const Jet<T, N>& f, double g) {
    const T tmp = g * pow(f.a, g - T(1.0));
    Jet<T, N> h;
    h.a = pow(f.a, g);
    for (int i = 0; i < N; ++i) h.v[i] = tmp * f.v[i];
    return h;
}
template <typename T, int N> inline bool isfinite(const Jet<T, N>& f) {
    if (!std::isfinite(f.a)) return false;
    for (int i = 0; i < N; ++i) if (!std::isfinite(f.v[i])) return false;
    return true;
}
template <typename T, int N> inline bool isnan(const Jet<T, N>& f) {
    if (std::isnan(f.a)) return true;
    for (int i = 0; i < N; ++i) if (std::isnan(f.v[i])) return true;
    return false;
}
template <typename T, int N> inline bool isinf(const Jet<T, N>& f) {
    if (std::isinf(f.a)) return true;
    for (int i = 0; i < N; ++i) if (std::isinf(f.v[i])) return true;
    return false;
}
inline bool IsFinite(double x) { return std::isfinite(x); }
inline bool IsNaN(double x) { return std::isnan(x); }
template <typename T, int N> inline bool IsFinite(const Jet<T, N>& f) { return isfinite(f); }
template <typename T, int N> inline bool IsNaN(const Jet<T, N>& f) { return isnan(f); }
template <typename T, int N> inline std::ostream& operator<<(std::ostream& s, const Jet<T, N>& z) { return s << "[" << z.a << " ; ...]"; }

}  // namespace ceres
