// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED.
//
// lie.hpp — restatement of the reference's math primitives, templated on the scalar type so the
// same code runs on double and on Jet<N>:
//   lie::{normalize_so3, exp_so3, log_SO3, log_SE3, make_tf}   src/utilies/common.h:121-181
//   e_laser::dis_from_line                                      src/utilies/common.h:86-95
//   convert::cross_matrix                                       src/utilies/common.h:16-29
// plus the public formulas of the two un-vendored dependencies those call:
//   ceres::AngleAxisToQuaternion / QuaternionToAngleAxis (Ceres 1.14 rotation.h)
//   Eigen::Quaternion::toRotationMatrix / Quaternion(Matrix3) / normalize (Eigen 3.3.7)
#pragma once
#include "jet.hpp"

namespace oracle {

template <class T> struct Vec3 {
    T x, y, z;
    Vec3() : x(0.0), y(0.0), z(0.0) {}
    Vec3(const T& x_, const T& y_, const T& z_) : x(x_), y(y_), z(z_) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T> inline Vec3<T> operator+(const Vec3<T>& a, const Vec3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline Vec3<T> operator-(const Vec3<T>& a, const Vec3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline Vec3<T> operator-(const Vec3<T>& a) { return {-a.x, -a.y, -a.z}; }
template <class T> inline Vec3<T> operator*(const Vec3<T>& a, const T& s) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline Vec3<T> operator*(const T& s, const Vec3<T>& a) { return {a.x * s, a.y * s, a.z * s}; }
template <class T> inline Vec3<T> operator/(const Vec3<T>& a, const T& s) { return {a.x / s, a.y / s, a.z / s}; }
template <class T> inline T dot(const Vec3<T>& a, const Vec3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline Vec3<T> cross(const Vec3<T>& a, const Vec3<T>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class T> inline T squared_norm(const Vec3<T>& a) { return dot(a, a); }
template <class T> inline T norm(const Vec3<T>& a) { return sqrt(squared_norm(a)); }
// Eigen::MatrixBase::normalized(): divide by the norm when the squared norm is > 0
template <class T> inline Vec3<T> normalized(const Vec3<T>& a) {
    T z = squared_norm(a);
    if (z > 0.0) return a / sqrt(z);
    return a;
}
template <class T, class S> inline Vec3<T> cast_vec(const Vec3<S>& a) { return {T(a.x), T(a.y), T(a.z)}; }

template <class T> struct Mat3 {
    T m[3][3];
    Mat3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = T(0.0); }
    static Mat3 identity() { Mat3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = T(1.0); return r; }
    T* operator[](int i) { return m[i]; }
    const T* operator[](int i) const { return m[i]; }
};
template <class T> inline Mat3<T> operator*(const Mat3<T>& a, const Mat3<T>& b) {
    Mat3<T> r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
template <class T> inline Vec3<T> operator*(const Mat3<T>& a, const Vec3<T>& v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
template <class T> inline Mat3<T> transpose(const Mat3<T>& a) {
    Mat3<T> r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
template <class T> inline Mat3<T> operator-(const Mat3<T>& a) {
    Mat3<T> r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = -a.m[i][j];
    return r;
}
template <class T, class S> inline Mat3<T> cast_mat(const Mat3<S>& a) {
    Mat3<T> r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = T(a.m[i][j]);
    return r;
}

// Eigen::Transform<T,3,Isometry>: [R | t]
template <class T> struct Iso3 {
    Mat3<T> R;
    Vec3<T> t;
    Iso3() : R(Mat3<T>::identity()), t() {}
    Iso3(const Mat3<T>& R_, const Vec3<T>& t_) : R(R_), t(t_) {}
};
template <class T> inline Iso3<T> operator*(const Iso3<T>& a, const Iso3<T>& b) { return {a.R * b.R, a.R * b.t + a.t}; }
template <class T> inline Vec3<T> operator*(const Iso3<T>& a, const Vec3<T>& v) { return a.R * v + a.t; }
// Transform<.,Isometry>::inverse(): R^T, -R^T t
template <class T> inline Iso3<T> inverse(const Iso3<T>& a) { Mat3<T> Rt = transpose(a.R); return {Rt, -(Rt * a.t)}; }
template <class T, class S> inline Iso3<T> cast_iso(const Iso3<S>& a) { return {cast_mat<T>(a.R), cast_vec<T>(a.t)}; }
inline Iso3<double> iso_from_rowmajor_3x4(const double* m) {
    Iso3<double> r;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) r.R.m[i][j] = m[i * 4 + j]; r.t[i] = m[i * 4 + 3]; }
    return r;
}

namespace convert {
// src/utilies/common.h:16-29
template <class T> inline Mat3<T> cross_matrix(const Vec3<T>& v) {
    Mat3<T> r;
    r.m[0][1] = -v.z; r.m[1][0] = v.z;
    r.m[0][2] = v.y;  r.m[2][0] = -v.y;
    r.m[1][2] = -v.x; r.m[2][1] = v.x;
    return r;
}
}  // namespace convert

namespace rot {
// ceres::AngleAxisToQuaternion (w,x,y,z)
template <class T> inline void angle_axis_to_quaternion(const T* aa, T* q) {
    const T& a0 = aa[0]; const T& a1 = aa[1]; const T& a2 = aa[2];
    const T theta_squared = a0 * a0 + a1 * a1 + a2 * a2;
    if (theta_squared > 0.0) {
        const T theta = sqrt(theta_squared);
        const T half_theta = theta * T(0.5);
        const T k = sin(half_theta) / theta;
        q[0] = cos(half_theta); q[1] = a0 * k; q[2] = a1 * k; q[3] = a2 * k;
    } else {
        const T k(0.5);
        q[0] = T(1.0); q[1] = a0 * k; q[2] = a1 * k; q[3] = a2 * k;
    }
}
// ceres::QuaternionToAngleAxis
template <class T> inline void quaternion_to_angle_axis(const T* q, T* aa) {
    const T& q1 = q[1]; const T& q2 = q[2]; const T& q3 = q[3];
    const T sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3;
    if (sin_squared_theta > 0.0) {
        const T sin_theta = sqrt(sin_squared_theta);
        const T& cos_theta = q[0];
        const T two_theta = T(2.0) * ((cos_theta < 0.0) ? atan2(-sin_theta, -cos_theta) : atan2(sin_theta, cos_theta));
        const T k = two_theta / sin_theta;
        aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
    } else {
        const T k(2.0);
        aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
    }
}
// Eigen::QuaternionBase::toRotationMatrix (no renormalisation)
template <class T> inline Mat3<T> quaternion_to_matrix(const T& w, const T& x, const T& y, const T& z) {
    const T tx = T(2.0) * x, ty = T(2.0) * y, tz = T(2.0) * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w;
    const T txx = tx * x, txy = ty * x, txz = tz * x;
    const T tyy = ty * y, tyz = tz * y, tzz = tz * z;
    Mat3<T> r;
    r.m[0][0] = T(1.0) - (tyy + tzz); r.m[0][1] = txy - twz;            r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;            r.m[1][1] = T(1.0) - (txx + tzz); r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;            r.m[2][1] = tyz + twx;            r.m[2][2] = T(1.0) - (txx + tyy);
    return r;
}
// Eigen::Quaternion(Matrix3): trace branch, else largest-diagonal branch.  q = (w, x, y, z)
template <class T> inline void matrix_to_quaternion(const Mat3<T>& m, T* q) {
    T t = m.m[0][0] + m.m[1][1] + m.m[2][2];
    if (t > 0.0) {
        t = sqrt(t + T(1.0));
        q[0] = T(0.5) * t;
        t = T(0.5) / t;
        q[1] = (m.m[2][1] - m.m[1][2]) * t;
        q[2] = (m.m[0][2] - m.m[2][0]) * t;
        q[3] = (m.m[1][0] - m.m[0][1]) * t;
    } else {
        int i = 0;
        if (m.m[1][1] > m.m[0][0]) i = 1;
        if (m.m[2][2] > m.m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m.m[i][i] - m.m[j][j] - m.m[k][k] + T(1.0));
        q[1 + i] = T(0.5) * t;
        t = T(0.5) / t;
        q[0] = (m.m[k][j] - m.m[j][k]) * t;
        q[1 + j] = (m.m[j][i] + m.m[i][j]) * t;
        q[1 + k] = (m.m[k][i] + m.m[i][k]) * t;
    }
}
}  // namespace rot

namespace lie {
// src/utilies/common.h:121-135
template <class T> inline void normalize_so3(Vec3<T>& so3) {
    T angle = norm(so3);
    T normalize_angle = angle;
    const T two_pi(2.0 * M_PI);
    const T pi(M_PI);
    if (angle > pi)
        normalize_angle -= two_pi * floor((angle + pi) / two_pi);
    else
        return;
    so3 = so3 / angle;
    so3 = so3 * normalize_angle;
}
// src/utilies/common.h:137-146
template <class T> inline Mat3<T> exp_so3(const Vec3<T>& so3) {
    T aa[3] = {so3.x, so3.y, so3.z};
    T q[4];
    rot::angle_axis_to_quaternion(aa, q);
    return rot::quaternion_to_matrix(q[0], q[1], q[2], q[3]);
}
// src/utilies/common.h:148-163
template <class T> inline Vec3<T> log_SO3(const Mat3<T>& R) {
    T q[4];
    rot::matrix_to_quaternion(R, q);
    // Eigen::Quaternion::normalize(): coeffs /= norm
    const T n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] = q[i] / n;
    T aa[3];
    rot::quaternion_to_angle_axis(q, aa);
    Vec3<T> ret(aa[0], aa[1], aa[2]);
    normalize_so3(ret);
    return ret;
}
// src/utilies/common.h:173-181
template <class T> inline Iso3<T> make_tf(const Vec3<T>& p, const Vec3<T>& so3) { return {exp_so3(so3), p}; }
// src/utilies/common.h:165-171
template <class T> inline void log_SE3(const Iso3<T>& tf, Vec3<T>& p, Vec3<T>& so3) { p = tf.t; so3 = log_SO3(tf.R); }
}  // namespace lie

namespace e_laser {
// src/utilies/common.h:86-95 (including the redundant second normalisation of `line`)
template <class T> inline T dis_from_line(const Vec3<T>& p, const Vec3<T>& p1, const Vec3<T>& p2) {
    Vec3<T> line = normalized(p2 - p1);
    Vec3<T> p2p = p - p2;
    return norm(p2p - dot(normalized(line), p2p) * line);
}
}  // namespace e_laser

}  // namespace oracle
