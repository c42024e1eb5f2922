// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the product
// path (2dliw-slam_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, and only as the checker / the timed CPU baseline.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path and
// cannot be compiled here (Eigen, Ceres, ROS are absent), so this restatement is pinned only by
// independent cross-checks (tests/test_oracle_*.py: numpy/torch-autograd float64 re-derivations,
// central finite differences, scipy.optimize.least_squares) — see DESIGN.md §Oracle.
//
// jet.hpp — forward-mode dual numbers with N infinitesimals, restating the public semantics of
// ceres::Jet<double,N> (Ceres Solver 1.14, un-vendored dependency of the reference: every
// `ceres::AutoDiffCostFunction` in src/factor/*.h evaluates the functors on this type).
#pragma once
#include <cmath>

namespace oracle {

template <int N>
struct Jet {
    double a;
    double v[N];
    Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
    Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT implicit like ceres::Jet
    Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
    Jet& operator+=(const Jet& o) { a += o.a; for (int i = 0; i < N; ++i) v[i] += o.v[i]; return *this; }
    Jet& operator-=(const Jet& o) { a -= o.a; for (int i = 0; i < N; ++i) v[i] -= o.v[i]; return *this; }
    Jet& operator*=(const Jet& o) { *this = *this * o; return *this; }
    Jet& operator/=(const Jet& o) { *this = *this / o; return *this; }
};

template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> r; r.a = -f.a; for (int i = 0; i < N; ++i) r.v[i] = -f.v[i]; return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a + g.a; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] + g.v[i]; return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> r = f; r.a += s; return r; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> r = f; r.a += s; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a - g.a; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] - g.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> r = f; r.a -= s; return r; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> r; r.a = s - f.a; for (int i = 0; i < N; ++i) r.v[i] = -f.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> r; r.a = f.a * g.a; for (int i = 0; i < N; ++i) r.v[i] = f.a * g.v[i] + f.v[i] * g.a; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> r; r.a = f.a * s; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * s; return r; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
    // ceres::Jet: one reciprocal of the scalar part, then (f.v - f.a/g.a * g.v) / g.a
    const double gi = 1.0 / g.a;
    const double q = f.a * gi;
    Jet<N> r; r.a = q; for (int i = 0; i < N; ++i) r.v[i] = (f.v[i] - q * g.v[i]) * gi; return r;
}
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) { const double si = 1.0 / s; return f * si; }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) {
    const double m = -s / (g.a * g.a);
    Jet<N> r; r.a = s / g.a; for (int i = 0; i < N; ++i) r.v[i] = g.v[i] * m; return r;
}

// comparisons look at the scalar part only (ceres::Jet comparison operators)
#define ORACLE_JET_CMP(op)                                                                      \
    template <int N> inline bool operator op(const Jet<N>& f, const Jet<N>& g) { return f.a op g.a; } \
    template <int N> inline bool operator op(const Jet<N>& f, double s) { return f.a op s; }          \
    template <int N> inline bool operator op(double s, const Jet<N>& g) { return s op g.a; }
ORACLE_JET_CMP(<)
ORACLE_JET_CMP(<=)
ORACLE_JET_CMP(>)
ORACLE_JET_CMP(>=)
ORACLE_JET_CMP(==)
ORACLE_JET_CMP(!=)
#undef ORACLE_JET_CMP

template <int N> inline Jet<N> jet_chain(double fa, double dfa, const Jet<N>& f) { Jet<N> r; r.a = fa; for (int i = 0; i < N; ++i) r.v[i] = dfa * f.v[i]; return r; }

inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double asin(double x) { return std::asin(x); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double floor(double x) { return std::floor(x); }
inline double scalar_part(double x) { return x; }

template <int N> inline Jet<N> sqrt(const Jet<N>& f) { const double t = std::sqrt(f.a); return jet_chain(t, 1.0 / (2.0 * t), f); }
template <int N> inline Jet<N> sin(const Jet<N>& f) { return jet_chain(std::sin(f.a), std::cos(f.a), f); }
template <int N> inline Jet<N> cos(const Jet<N>& f) { return jet_chain(std::cos(f.a), -std::sin(f.a), f); }
template <int N> inline Jet<N> asin(const Jet<N>& f) { return jet_chain(std::asin(f.a), 1.0 / std::sqrt(1.0 - f.a * f.a), f); }
// ceres::floor on a Jet has zero derivative (used by lie::normalize_so3, src/utilies/common.h:130)
template <int N> inline Jet<N> floor(const Jet<N>& f) { return Jet<N>(std::floor(f.a)); }
template <int N> inline Jet<N> atan2(const Jet<N>& g, const Jet<N>& f) {
    // d atan2(g, f) = (f dg - g df) / (f^2 + g^2)
    const double t = 1.0 / (f.a * f.a + g.a * g.a);
    Jet<N> r; r.a = std::atan2(g.a, f.a);
    for (int i = 0; i < N; ++i) r.v[i] = t * (f.a * g.v[i] - g.a * f.v[i]);
    return r;
}
template <int N> inline double scalar_part(const Jet<N>& f) { return f.a; }

}  // namespace oracle
