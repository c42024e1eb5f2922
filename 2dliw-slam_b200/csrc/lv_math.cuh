// lv_math.cuh — scalar/dual-number rotation math and the small factors of the front-end solver,
// written once for device code and (for CPU unit tests of the formulas only) host code.
//
// What it restates (reference file:line):
//   lie::exp_so3 / log_SO3 / normalize_so3          src/utilies/common.h:121-163
//   so3_parameterization (Plus = wrap(theta+delta)) src/factor/factor_common.h:40-53
//   imu_factor residual                             src/factor/imu_factor.h:52-86
//   wheel_odom_factor residual                      src/factor/wheel_factor.h:20-71
//   ground_factor_p / ground_factor_q               src/factor/ground_factor.h:27-48, :59-82
// The reference differentiates these with ceres::Jet<double,N>.  Here the derivative is carried by a
// single-direction dual number (`Dual`): one lane of a warp evaluates one column of the Jacobian, and
// only the non-linear pieces (rotations) are pushed through duals — the linear blocks of each
// Jacobian are filled in closed form (see imu_jacobian_column).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LV_HD __host__ __device__ __forceinline__
#else
#define LV_HD inline
#endif

namespace lv {

constexpr double kPi = 3.14159265358979323846;

// ------------------------------------------------------------------ dual number a + d*eps
struct Dual {
    double a, d;
    LV_HD Dual() : a(0.0), d(0.0) {}
    LV_HD Dual(double a_) : a(a_), d(0.0) {}
    LV_HD Dual(double a_, double d_) : a(a_), d(d_) {}
};
LV_HD Dual operator+(Dual x, Dual y) { return Dual(x.a + y.a, x.d + y.d); }
LV_HD Dual operator-(Dual x, Dual y) { return Dual(x.a - y.a, x.d - y.d); }
LV_HD Dual operator-(Dual x) { return Dual(-x.a, -x.d); }
LV_HD Dual operator*(Dual x, Dual y) { return Dual(x.a * y.a, x.a * y.d + x.d * y.a); }
LV_HD Dual operator/(Dual x, Dual y) {
    const double yi = 1.0 / y.a, q = x.a * yi;
    return Dual(q, (x.d - q * y.d) * yi);
}
LV_HD Dual operator+(Dual x, double s) { return Dual(x.a + s, x.d); }
LV_HD Dual operator+(double s, Dual x) { return Dual(x.a + s, x.d); }
LV_HD Dual operator-(Dual x, double s) { return Dual(x.a - s, x.d); }
LV_HD Dual operator-(double s, Dual x) { return Dual(s - x.a, -x.d); }
LV_HD Dual operator*(Dual x, double s) { return Dual(x.a * s, x.d * s); }
LV_HD Dual operator*(double s, Dual x) { return Dual(x.a * s, x.d * s); }
LV_HD Dual operator/(double s, Dual y) { const double yi = 1.0 / y.a, q = s * yi; return Dual(q, -q * y.d * yi); }
LV_HD Dual operator/(Dual x, double s) { const double si = 1.0 / s; return Dual(x.a * si, x.d * si); }

LV_HD double val(double x) { return x; }
LV_HD double val(Dual x) { return x.a; }
LV_HD double lv_sqrt(double x) { return sqrt(x); }
LV_HD Dual lv_sqrt(Dual x) {
    const double t = sqrt(x.a);
#if defined(__CUDA_ARCH__)
    return Dual(t, x.d * (0.5 * rsqrt(x.a)));   // derivative through the cheaper reciprocal square root
#else
    return Dual(t, x.d / (2.0 * t));
#endif
}
LV_HD void lv_sincos(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    sincos(x, s, c);
#else
    *s = sin(x); *c = cos(x);
#endif
}
LV_HD void lv_sincos(Dual x, Dual* s, Dual* c) {
    double sv, cv;
    lv_sincos(x.a, &sv, &cv);
    *s = Dual(sv, cv * x.d);
    *c = Dual(cv, -sv * x.d);
}
LV_HD double lv_asin(double x) { return asin(x); }
LV_HD Dual lv_asin(Dual x) { return Dual(asin(x.a), x.d / sqrt(1.0 - x.a * x.a)); }
LV_HD double lv_atan2(double y, double x) { return atan2(y, x); }
LV_HD Dual lv_atan2(Dual y, Dual x) {
    const double t = 1.0 / (x.a * x.a + y.a * y.a);
    return Dual(atan2(y.a, x.a), t * (x.a * y.d - y.a * x.d));
}

// ------------------------------------------------------------------ 3-vectors / 3x3 (row-major)
template <class T> struct V3 { T x, y, z; };
template <class T> struct M3 { T m[9]; };

template <class T> LV_HD V3<T> v3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <class T> LV_HD V3<T> operator+(const V3<T>& a, const V3<T>& b) { return v3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> LV_HD V3<T> operator-(const V3<T>& a, const V3<T>& b) { return v3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> LV_HD V3<T> neg(const V3<T>& a) { return v3<T>(-a.x, -a.y, -a.z); }
template <class T> LV_HD V3<T> scale(const V3<T>& a, T s) { return v3<T>(a.x * s, a.y * s, a.z * s); }
template <class T> LV_HD T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> LV_HD V3<T> cross(const V3<T>& a, const V3<T>& b) {
    return v3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class T> LV_HD T norm(const V3<T>& a) { return lv_sqrt(dot(a, a)); }
template <class T> LV_HD V3<T> mul(const M3<T>& A, const V3<T>& v) {
    return v3<T>(A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
                 A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z);
}
template <class T> LV_HD V3<T> mul_t(const M3<T>& A, const V3<T>& v) {  // A^T v
    return v3<T>(A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
                 A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z);
}
template <class T> LV_HD M3<T> mul(const M3<T>& A, const M3<T>& B) {
    M3<T> C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C.m[i * 3 + j] = A.m[i * 3] * B.m[j] + A.m[i * 3 + 1] * B.m[3 + j] + A.m[i * 3 + 2] * B.m[6 + j];
    return C;
}
template <class T> LV_HD M3<T> mul_tn(const M3<T>& A, const M3<T>& B) {  // A^T B
    M3<T> C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C.m[i * 3 + j] = A.m[i] * B.m[j] + A.m[3 + i] * B.m[3 + j] + A.m[6 + i] * B.m[6 + j];
    return C;
}
template <class T> LV_HD M3<T> lift(const M3<double>& A) {
    M3<T> C;
#pragma unroll
    for (int i = 0; i < 9; ++i) C.m[i] = T(A.m[i]);
    return C;
}
template <class T> LV_HD V3<T> lift(const V3<double>& a) { return v3<T>(T(a.x), T(a.y), T(a.z)); }
LV_HD V3<double> load3(const double* p) { return v3<double>(p[0], p[1], p[2]); }

// rigid transform [R | t] from a row-major 3x4 array
struct Iso { M3<double> R; V3<double> t; };
LV_HD Iso load_iso(const double* m) {
    Iso T;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) T.R.m[i * 3 + j] = m[i * 4 + j];
    }
    T.t = v3<double>(m[3], m[7], m[11]);
    return T;
}

// ------------------------------------------------------------------ SO(3)
// angle-axis -> unit quaternion -> rotation matrix (common.h:137-146 through ceres::AngleAxisToQuaternion
// and Eigen::Quaternion::toRotationMatrix)
template <class T> LV_HD M3<T> exp_so3(const V3<T>& w) {
    const T th2 = w.x * w.x + w.y * w.y + w.z * w.z;
    T qw, k;
    if (val(th2) > 0.0) {
        const T th = lv_sqrt(th2);
        T s, c;
        lv_sincos(th * 0.5, &s, &c);
        k = s / th;
        qw = c;
    } else {
        k = T(0.5);
        qw = T(1.0);
    }
    const T qx = w.x * k, qy = w.y * k, qz = w.z * k;
    const T tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    const T twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const T txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const T tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    M3<T> R;
    R.m[0] = 1.0 - (tyy + tzz); R.m[1] = txy - twz;         R.m[2] = txz + twy;
    R.m[3] = txy + twz;         R.m[4] = 1.0 - (txx + tzz); R.m[5] = tyz - twx;
    R.m[6] = txz - twy;         R.m[7] = tyz + twx;         R.m[8] = 1.0 - (txx + tyy);
    return R;
}

// common.h:121-135: scale an angle-axis vector whose norm exceeds pi back into (-pi, pi]
template <class T> LV_HD V3<T> normalize_so3(const V3<T>& w) {
    const T angle = norm(w);
    if (!(val(angle) > kPi)) return w;
    const double turns = floor((val(angle) + kPi) / (2.0 * kPi));  // floor: zero derivative
    const T wrapped = angle - (2.0 * kPi) * turns;
    return scale(scale(w, T(1.0) / angle), wrapped);
}

// rotation matrix -> quaternion (Eigen::Quaternion(Matrix3) branches) -> normalised -> angle-axis
// (ceres::QuaternionToAngleAxis) -> wrapped (common.h:148-163)
template <class T> LV_HD V3<T> log_so3(const M3<T>& R) {
    T qw, qx, qy, qz;
    T t = R.m[0] + R.m[4] + R.m[8];
    if (val(t) > 0.0) {
        t = lv_sqrt(t + 1.0);
        qw = 0.5 * t;
        t = 0.5 / t;
        qx = (R.m[7] - R.m[5]) * t;
        qy = (R.m[2] - R.m[6]) * t;
        qz = (R.m[3] - R.m[1]) * t;
    } else {
        // largest diagonal element i, then j = i+1, k = i+2 (mod 3); spelled out so that every index is static
        int i = 0;
        if (val(R.m[4]) > val(R.m[0])) i = 1;
        if (val(R.m[8]) > val(i == 0 ? R.m[0] : R.m[4])) i = 2;
        if (i == 0) {
            t = lv_sqrt(R.m[0] - R.m[4] - R.m[8] + 1.0);
            qx = 0.5 * t; t = 0.5 / t;
            qw = (R.m[7] - R.m[5]) * t; qy = (R.m[3] + R.m[1]) * t; qz = (R.m[6] + R.m[2]) * t;
        } else if (i == 1) {
            t = lv_sqrt(R.m[4] - R.m[8] - R.m[0] + 1.0);
            qy = 0.5 * t; t = 0.5 / t;
            qw = (R.m[2] - R.m[6]) * t; qz = (R.m[7] + R.m[5]) * t; qx = (R.m[1] + R.m[3]) * t;
        } else {
            t = lv_sqrt(R.m[8] - R.m[0] - R.m[4] + 1.0);
            qz = 0.5 * t; t = 0.5 / t;
            qw = (R.m[3] - R.m[1]) * t; qx = (R.m[2] + R.m[6]) * t; qy = (R.m[5] + R.m[7]) * t;
        }
    }
    const T qn = lv_sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    const T qi = T(1.0) / qn;
    const T w = qw * qi, x = qx * qi, y = qy * qi, z = qz * qi;
    const T s2 = x * x + y * y + z * z;
    T k;
    if (val(s2) > 0.0) {
        const T s = lv_sqrt(s2);
        const T two_theta = 2.0 * ((val(w) < 0.0) ? lv_atan2(-s, -w) : lv_atan2(s, w));
        k = two_theta / s;
    } else {
        k = T(2.0);
    }
    return normalize_so3(v3<T>(x * k, y * k, z * k));
}

// so3_parameterization::operator() (factor_common.h:40-53)
LV_HD void so3_plus(const double* theta, const double* delta, double* out) {
    V3<double> w = normalize_so3(v3<double>(theta[0] + delta[0], theta[1] + delta[1], theta[2] + delta[2]));
    out[0] = w.x; out[1] = w.y; out[2] = w.z;
}

// ------------------------------------------------------------------ solver constants (PARAM() values)
struct Consts {
    double T_il[12];  // T_imu_to_laser, row-major 3x4
    double T_io[12];  // T_imu_to_wheel
    double g;
    double laser_sqrt_info;       // 1/line_to_line_sigma      laser_factor.h:21
    double ground_p_sqrt_info;    // 1/manifold_p_sigma        ground_factor.h:20
    double ground_q_sqrt_info;    // 1/manifold_q_sigma        ground_factor.h:21
    double Q[12];                 // diag(imu_noise::Q)        imu_preintegraption.h:30-42
    double wheel_cov[3];          // diag(wheel_noise::cov)    wheel_odom_preintegration.h:19-22
    double huber_delta;           // <= 0: none
};

// ------------------------------------------------------------------ ground factors (ground_factor.h)
// both residuals of one pose; T = double gives values, T = Dual with the pose seeded along one of the
// 6 directions gives one Jacobian column.
template <class T> LV_HD void ground_residuals(const Consts& C, const V3<T>& p, const V3<T>& th, T* res_p, T* res_q) {
    const Iso Tio = load_iso(C.T_io);
    const M3<T> R = exp_so3(th);
    // tf_w_o = make_tf(p, th) * T_io: height of the wheel frame and its z axis in the world
    const T height = R.m[6] * Tio.t.x + R.m[7] * Tio.t.y + R.m[8] * Tio.t.z + p.z;
    const V3<T> zax = v3<T>(R.m[0] * Tio.R.m[2] + R.m[1] * Tio.R.m[5] + R.m[2] * Tio.R.m[8],
                            R.m[3] * Tio.R.m[2] + R.m[4] * Tio.R.m[5] + R.m[5] * Tio.R.m[8],
                            R.m[6] * Tio.R.m[2] + R.m[7] * Tio.R.m[5] + R.m[8] * Tio.R.m[8]);
    const V3<T> ez = v3<T>(T(0.0), T(0.0), T(1.0));
    const T sinn = norm(cross(zax, ez));
    *res_p = C.ground_p_sqrt_info * height;
    *res_q = C.ground_q_sqrt_info * lv_asin(sinn);
}

// ------------------------------------------------------------------ wheel factor (wheel_factor.h:20-71)
// blob = delta_Tij[3x4] | sqrt_info diag[3]
template <class T>
LV_HD void wheel_residuals(const Consts& C, const double* blob, const V3<T>& pi, const V3<T>& thi, const V3<T>& pj,
                           const V3<T>& thj, T* res) {
    const Iso Tio = load_iso(C.T_io);
    const Iso dT = load_iso(blob);
    const M3<T> Ri = exp_so3(thi), Rj = exp_so3(thj);
    const M3<T> Rio = lift<T>(Tio.R);
    const V3<T> tio = lift<T>(Tio.t);
    // tf_i = make_tf(pi, thi) * T_io ; tf_j likewise ; w_tf_ij = tf_i^-1 * tf_j
    const M3<T> Roi = mul(Ri, Rio), Roj = mul(Rj, Rio);
    const V3<T> toi = mul(Ri, tio) + pi, toj = mul(Rj, tio) + pj;
    const V3<T> p = mul_t(Roi, toj - toi);
    const V3<T> q = log_so3(mul_tn(Roi, Roj));
    const V3<double> op = dT.t;
    const V3<double> oq = log_so3(dT.R);
    const double o_len = sqrt(op.x * op.x + op.y * op.y);
    const T len = lv_sqrt(p.x * p.x + p.y * p.y);
    T angle;
    if (o_len > 0.0001 && val(len) > 0.0001) {
        // |o_dir x dir| of the two unit xy-directions
        const T cz = (op.x / o_len) * (p.y / len) - (op.y / o_len) * (p.x / len);
        const T sinn = lv_sqrt(cz * cz);
        angle = lv_asin(sinn);
    } else {
        angle = len;
    }
    if (val(len) < 0.0001 || o_len < 0.0001)
        res[0] = blob[12] * len;
    else
        res[0] = blob[12] * (o_len - len);
    res[1] = blob[13] * angle;
    const T qn = norm(q);
    const double oqn = norm(oq);
    if (val(qn) < 0.001 || oqn < 0.001)
        res[2] = blob[14] * qn;
    else
        res[2] = blob[14] * (oqn - qn);
}

// ------------------------------------------------------------------ imu factor (imu_factor.h:52-86)
// blob = X[15] | J[15x15] | sqrt_inverse_P[15x15] | Dt.  States are [p q v ba bw].
// Un-whitened residual (before res_all = sqrt_info * res_all, imu_factor.h:85-86).
template <class T>
LV_HD void imu_rotation_residual(const double* blob, const V3<T>& thi, const V3<T>& bwi, const V3<T>& thj, V3<T>* r_gamma) {
    const double* X = blob;
    const double* J = blob + 15;
    const V3<T> dbw = v3<T>(bwi.x - X[12], bwi.y - X[13], bwi.z - X[14]);
    // gamma_hat = gamma + J_gamma_bw (bw_i - bw_lin)
    const V3<T> gam = v3<T>(X[6] + (J[6 * 15 + 12] * dbw.x + J[6 * 15 + 13] * dbw.y + J[6 * 15 + 14] * dbw.z),
                            X[7] + (J[7 * 15 + 12] * dbw.x + J[7 * 15 + 13] * dbw.y + J[7 * 15 + 14] * dbw.z),
                            X[8] + (J[8 * 15 + 12] * dbw.x + J[8 * 15 + 13] * dbw.y + J[8 * 15 + 14] * dbw.z));
    const M3<T> E = exp_so3(neg(gam));
    const M3<T> RiT = exp_so3(neg(thi));
    const M3<T> Rj = exp_so3(thj);
    *r_gamma = log_so3(mul(E, mul(RiT, Rj)));
}

LV_HD void imu_raw_residual(const Consts& C, const double* blob, const double* si, const double* sj, double* r /*[15]*/) {
    const double* X = blob;
    const double* J = blob + 15;
    const double Dt = blob[465];
    const V3<double> pi = load3(si), thi = load3(si + 3), vi = load3(si + 6), bai = load3(si + 9), bwi = load3(si + 12);
    const V3<double> pj = load3(sj), thj = load3(sj + 3), vj = load3(sj + 6), baj = load3(sj + 9), bwj = load3(sj + 12);
    const V3<double> dba = v3<double>(bai.x - X[9], bai.y - X[10], bai.z - X[11]);
    const V3<double> dbw = v3<double>(bwi.x - X[12], bwi.y - X[13], bwi.z - X[14]);
    double ab[6];
#pragma unroll
    for (int k = 0; k < 6; ++k)
        ab[k] = X[k] + (J[k * 15 + 9] * dba.x + J[k * 15 + 10] * dba.y + J[k * 15 + 11] * dba.z) +
                (J[k * 15 + 12] * dbw.x + J[k * 15 + 13] * dbw.y + J[k * 15 + 14] * dbw.z);
    const M3<double> RiT = exp_so3(neg(thi));
    const double gz = C.g;
    const V3<double> y1 = v3<double>(pj.x - pi.x - vi.x * Dt, pj.y - pi.y - vi.y * Dt, pj.z - pi.z + 0.5 * gz * Dt * Dt - vi.z * Dt);
    const V3<double> y2 = v3<double>(vj.x - vi.x, vj.y - vi.y, vj.z + gz * Dt - vi.z);
    const V3<double> a = mul(RiT, y1), b = mul(RiT, y2);
    r[0] = ab[0] - a.x; r[1] = ab[1] - a.y; r[2] = ab[2] - a.z;
    r[3] = ab[3] - b.x; r[4] = ab[4] - b.y; r[5] = ab[5] - b.z;
    V3<double> rg;
    imu_rotation_residual<double>(blob, thi, bwi, thj, &rg);
    r[6] = rg.x; r[7] = rg.y; r[8] = rg.z;
    r[9] = baj.x - bai.x; r[10] = baj.y - bai.y; r[11] = baj.z - bai.z;
    r[12] = bwj.x - bwi.x; r[13] = bwj.y - bwi.y; r[14] = bwj.z - bwi.z;
}

// Column `c` (0..29) of the un-whitened 15x30 Jacobian; columns follow the functor's parameter order
// (p_i q_i v_i bs_i p_j q_j v_j bs_j).  Linear blocks in closed form, rotation blocks through Dual.
LV_HD void imu_jacobian_column(const Consts& C, const double* blob, const double* si, const double* sj, int c, double* col /*[15]*/) {
    const double* J = blob + 15;
    const double Dt = blob[465];
#pragma unroll
    for (int k = 0; k < 15; ++k) col[k] = 0.0;
    const V3<double> thi = load3(si + 3), bwi = load3(si + 12), thj = load3(sj + 3);
    const int blk = c / 3, k = c % 3;  // 0 p_i 1 q_i 2 v_i 3 ba_i 4 bw_i 5 p_j 6 q_j 7 v_j 8 ba_j 9 bw_j
    if (blk == 1) {
        // theta_i: rows 0-5 = -(d R_i^T / d theta_k) y, rows 6-8 through the rotation residual
        V3<Dual> th = lift<Dual>(thi);
        (k == 0 ? th.x : (k == 1 ? th.y : th.z)).d = 1.0;
        const M3<Dual> RiT = exp_so3(neg(th));
        const V3<double> pi = load3(si), vi = load3(si + 6), pj = load3(sj), vj = load3(sj + 6);
        const double gz = C.g;
        const V3<double> y1 = v3<double>(pj.x - pi.x - vi.x * Dt, pj.y - pi.y - vi.y * Dt, pj.z - pi.z + 0.5 * gz * Dt * Dt - vi.z * Dt);
        const V3<double> y2 = v3<double>(vj.x - vi.x, vj.y - vi.y, vj.z + gz * Dt - vi.z);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            col[r] = -(RiT.m[r * 3].d * y1.x + RiT.m[r * 3 + 1].d * y1.y + RiT.m[r * 3 + 2].d * y1.z);
            col[3 + r] = -(RiT.m[r * 3].d * y2.x + RiT.m[r * 3 + 1].d * y2.y + RiT.m[r * 3 + 2].d * y2.z);
        }
        V3<Dual> rg;
        imu_rotation_residual<Dual>(blob, th, lift<Dual>(bwi), lift<Dual>(thj), &rg);
        col[6] = rg.x.d; col[7] = rg.y.d; col[8] = rg.z.d;
        return;
    }
    if (blk == 4) {
        // bw_i: first-order bias correction columns + rotation residual + (-I) on r_bw
#pragma unroll
        for (int r = 0; r < 6; ++r) col[r] = J[r * 15 + 12 + k];
        V3<Dual> bw = lift<Dual>(bwi);
        (k == 0 ? bw.x : (k == 1 ? bw.y : bw.z)).d = 1.0;
        V3<Dual> rg;
        imu_rotation_residual<Dual>(blob, lift<Dual>(thi), bw, lift<Dual>(thj), &rg);
        col[6] = rg.x.d; col[7] = rg.y.d; col[8] = rg.z.d;
        col[12 + k] = -1.0;
        return;
    }
    if (blk == 6) {
        V3<Dual> th = lift<Dual>(thj);
        (k == 0 ? th.x : (k == 1 ? th.y : th.z)).d = 1.0;
        V3<Dual> rg;
        imu_rotation_residual<Dual>(blob, lift<Dual>(thi), lift<Dual>(bwi), th, &rg);
        col[6] = rg.x.d; col[7] = rg.y.d; col[8] = rg.z.d;
        return;
    }
    if (blk == 3) {  // ba_i
#pragma unroll
        for (int r = 0; r < 6; ++r) col[r] = J[r * 15 + 9 + k];
        col[9 + k] = -1.0;
        return;
    }
    if (blk == 8) { col[9 + k] = 1.0; return; }    // ba_j
    if (blk == 9) { col[12 + k] = 1.0; return; }   // bw_j
    // p_i, v_i, p_j, v_j: +-R_i^T columns
    const M3<double> RiT = exp_so3(neg(thi));
    const double c0 = RiT.m[k], c1 = RiT.m[3 + k], c2 = RiT.m[6 + k];
    if (blk == 0) { col[0] = c0; col[1] = c1; col[2] = c2; }                       // d r_alpha / d p_i = +R_i^T
    if (blk == 5) { col[0] = -c0; col[1] = -c1; col[2] = -c2; }                    // d r_alpha / d p_j = -R_i^T
    if (blk == 2) { col[0] = c0 * Dt; col[1] = c1 * Dt; col[2] = c2 * Dt; col[3] = c0; col[4] = c1; col[5] = c2; }  // v_i
    if (blk == 7) { col[3] = -c0; col[4] = -c1; col[5] = -c2; }                    // v_j
}

// ------------------------------------------------------------------ all factors of one (frame i-1, frame i) item
// Generic in T so that one warp lane can carry one derivative direction (Dual seeded on one of the 30 state
// columns) while another lane carries the plain values: every lane runs the same instruction stream.  The two
// rotations exp(theta_i), exp(theta_j) are shared by the IMU, wheel and ground residuals
// (exp_so3(-theta) == exp_so3(theta)^T bit for bit: the quaternion only changes the sign of its vector part).
template <class T> struct FrameState { V3<T> p, th, v, ba, bw; };
template <class T> LV_HD M3<T> transpose(const M3<T>& A) {
    M3<T> B;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B.m[i * 3 + j] = A.m[j * 3 + i];
    return B;
}
// r_imu: the 15 residuals BEFORE whitening by sqrt_inverse_P.  X = the preintegrated state (15); J(k, 9 + c) =
// J[k * JS + JO + c], k = 0..8, c = 0..5 are the only entries of the preintegration Jacobian the factor reads (the bias
// columns of the alpha / beta / gamma rows, imu_factor.h:57-77): JS = 15, JO = 9 on the ABI blob, JS = 6, JO = 0 on the
// compact copy factor_pair_kernel stages in shared memory.
template <class T, int JS, int JO>
LV_HD void item_imu_t(const Consts& C, const double* X, const double* J, const double Dt, const FrameState<T>& a, const FrameState<T>& b,
                      const M3<T>& Ri, const M3<T>& Rj, T* r_imu) {
    const V3<T> dba = v3<T>(a.ba.x - X[9], a.ba.y - X[10], a.ba.z - X[11]);
    const V3<T> dbw = v3<T>(a.bw.x - X[12], a.bw.y - X[13], a.bw.z - X[14]);
    T ab[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        ab[k] = X[k] + (J[k * JS + JO + 3] * dbw.x + J[k * JS + JO + 4] * dbw.y + J[k * JS + JO + 5] * dbw.z);
        if (k < 6) ab[k] = ab[k] + (J[k * JS + JO] * dba.x + J[k * JS + JO + 1] * dba.y + J[k * JS + JO + 2] * dba.z);
    }
    const double gz = C.g;
    const V3<T> y1 = v3<T>(b.p.x - a.p.x - a.v.x * Dt, b.p.y - a.p.y - a.v.y * Dt, b.p.z - a.p.z + 0.5 * gz * Dt * Dt - a.v.z * Dt);
    const V3<T> y2 = v3<T>(b.v.x - a.v.x, b.v.y - a.v.y, b.v.z + gz * Dt - a.v.z);
    const V3<T> ra = mul_t(Ri, y1), rb = mul_t(Ri, y2);   // R_i^T y
    r_imu[0] = ab[0] - ra.x; r_imu[1] = ab[1] - ra.y; r_imu[2] = ab[2] - ra.z;
    r_imu[3] = ab[3] - rb.x; r_imu[4] = ab[4] - rb.y; r_imu[5] = ab[5] - rb.z;
    const M3<T> E = exp_so3(v3<T>(-ab[6], -ab[7], -ab[8]));
    const V3<T> rg = log_so3(mul(E, mul_tn(Ri, Rj)));
    r_imu[6] = rg.x; r_imu[7] = rg.y; r_imu[8] = rg.z;
    r_imu[9] = b.ba.x - a.ba.x; r_imu[10] = b.ba.y - a.ba.y; r_imu[11] = b.ba.z - a.ba.z;
    r_imu[12] = b.bw.x - a.bw.x; r_imu[13] = b.bw.y - a.bw.y; r_imu[14] = b.bw.z - a.bw.z;
}
template <class T>
LV_HD void item_imu(const Consts& C, const double* imu_blob, const FrameState<T>& a, const FrameState<T>& b, const M3<T>& Ri,
                    const M3<T>& Rj, T* r_imu) {
    item_imu_t<T, 15, 9>(C, imu_blob, imu_blob + 15, imu_blob[465], a, b, Ri, Rj, r_imu);
}
template <class T>
LV_HD void item_wheel(const Consts& C, const double* wheel_blob, const V3<T>& pa, const V3<T>& pb, const M3<T>& Ri, const M3<T>& Rj,
                      T* r_wheel) {
    const Iso Tio = load_iso(C.T_io);
    const Iso dT = load_iso(wheel_blob);
    const M3<T> Rio = lift<T>(Tio.R);
    const V3<T> tio = lift<T>(Tio.t);
    const M3<T> Roi = mul(Ri, Rio), Roj = mul(Rj, Rio);
    const V3<T> toi = mul(Ri, tio) + pa, toj = mul(Rj, tio) + pb;
    const V3<T> p = mul_t(Roi, toj - toi);
    const V3<T> q = log_so3(mul_tn(Roi, Roj));
    const V3<double> op = dT.t;
    const V3<double> oq = log_so3(dT.R);
    const double o_len = sqrt(op.x * op.x + op.y * op.y);
    const T len = lv_sqrt(p.x * p.x + p.y * p.y);
    T angle;
    if (o_len > 0.0001 && val(len) > 0.0001) {
        const T cz = (op.x / o_len) * (p.y / len) - (op.y / o_len) * (p.x / len);
        angle = lv_asin(lv_sqrt(cz * cz));
    } else {
        angle = len;
    }
    if (val(len) < 0.0001 || o_len < 0.0001) r_wheel[0] = wheel_blob[12] * len;
    else r_wheel[0] = wheel_blob[12] * (o_len - len);
    r_wheel[1] = wheel_blob[13] * angle;
    const T qn = norm(q);
    const double oqn = norm(oq);
    if (val(qn) < 0.001 || oqn < 0.001) r_wheel[2] = wheel_blob[14] * qn;
    else r_wheel[2] = wheel_blob[14] * (oqn - qn);
}
template <class T> LV_HD void item_ground(const Consts& C, const V3<T>& pb, const M3<T>& Rj, T* r_ground) {
    const Iso Tio = load_iso(C.T_io);
    const T height = Rj.m[6] * Tio.t.x + Rj.m[7] * Tio.t.y + Rj.m[8] * Tio.t.z + pb.z;
    const V3<T> zax = v3<T>(Rj.m[0] * Tio.R.m[2] + Rj.m[1] * Tio.R.m[5] + Rj.m[2] * Tio.R.m[8],
                            Rj.m[3] * Tio.R.m[2] + Rj.m[4] * Tio.R.m[5] + Rj.m[5] * Tio.R.m[8],
                            Rj.m[6] * Tio.R.m[2] + Rj.m[7] * Tio.R.m[5] + Rj.m[8] * Tio.R.m[8]);
    const T sinn = norm(cross(zax, v3<T>(T(0.0), T(0.0), T(1.0))));
    r_ground[0] = C.ground_p_sqrt_info * height;
    r_ground[1] = C.ground_q_sqrt_info * lv_asin(sinn);
}
// all three together (host tests)
template <class T>
LV_HD void item_residuals(const Consts& C, const double* imu_blob, const double* wheel_blob, bool ground, const FrameState<T>& a,
                          const FrameState<T>& b, T* r_imu, T* r_wheel, T* r_ground) {
    const M3<T> Ri = exp_so3(a.th), Rj = exp_so3(b.th);
    if (imu_blob) item_imu(C, imu_blob, a, b, Ri, Rj, r_imu);
    if (wheel_blob) item_wheel(C, wheel_blob, a.p, b.p, Ri, Rj, r_wheel);
    if (ground) item_ground(C, b.p, Rj, r_ground);
}
LV_HD FrameState<Dual> seed_frame_state(const double* s, int seed /* column 0..14 or -1 */) {
    // branch-free seeding keeps the struct in registers (no address is taken)
    FrameState<Dual> f;
    f.p = v3<Dual>(Dual(s[0], seed == 0 ? 1.0 : 0.0), Dual(s[1], seed == 1 ? 1.0 : 0.0), Dual(s[2], seed == 2 ? 1.0 : 0.0));
    f.th = v3<Dual>(Dual(s[3], seed == 3 ? 1.0 : 0.0), Dual(s[4], seed == 4 ? 1.0 : 0.0), Dual(s[5], seed == 5 ? 1.0 : 0.0));
    f.v = v3<Dual>(Dual(s[6], seed == 6 ? 1.0 : 0.0), Dual(s[7], seed == 7 ? 1.0 : 0.0), Dual(s[8], seed == 8 ? 1.0 : 0.0));
    f.ba = v3<Dual>(Dual(s[9], seed == 9 ? 1.0 : 0.0), Dual(s[10], seed == 10 ? 1.0 : 0.0), Dual(s[11], seed == 11 ? 1.0 : 0.0));
    f.bw = v3<Dual>(Dual(s[12], seed == 12 ? 1.0 : 0.0), Dual(s[13], seed == 13 ? 1.0 : 0.0), Dual(s[14], seed == 14 ? 1.0 : 0.0));
    return f;
}

// ------------------------------------------------------------------ closed-form item Jacobians
// The same derivatives as the dual-number path above, written with the SO(3) right Jacobian
//   R(theta + d) = R(theta) Exp(J_r(theta) d),   Log(M Exp(e)) = Log(M) + J_r^-1(Log M) e
// so that the expensive pieces (3 exponentials, 2 logarithms, 3 J_r, 2 J_r^-1) are computed ONCE per item and every
// warp lane only does a few 3x3 matrix-vector products for its own state column.  The angle-axis parameterisation is
// additive (so3_parameterization, factor_common.h:40-53), hence the plain d/d(theta) columns.
// (Ceres differentiates the quaternion formulas with Jets; they are the same functions, see tests.)
LV_HD M3<double> so3_right_jacobian(const V3<double>& v) {
    const double t2 = v.x * v.x + v.y * v.y + v.z * v.z, t = sqrt(t2);
    double a, b;  // J_r = I - a [v]x + b [v]x^2
    if (t < 0.05) {
        a = 0.5 - t2 * (1.0 / 24.0) + t2 * t2 * (1.0 / 720.0) - t2 * t2 * t2 * (1.0 / 40320.0);
        b = 1.0 / 6.0 - t2 * (1.0 / 120.0) + t2 * t2 * (1.0 / 5040.0) - t2 * t2 * t2 * (1.0 / 362880.0);
    } else {
        double sh, ch;
        lv_sincos(0.5 * t, &sh, &ch);
        a = 2.0 * sh * sh / t2;                 // (1 - cos t) / t^2
        b = (t - 2.0 * sh * ch) / (t2 * t);     // (t - sin t) / t^3
    }
    const double xx = v.x * v.x, yy = v.y * v.y, zz = v.z * v.z, xy = v.x * v.y, xz = v.x * v.z, yz = v.y * v.z;
    M3<double> J;
    J.m[0] = 1.0 - b * (yy + zz); J.m[1] = a * v.z + b * xy;    J.m[2] = -a * v.y + b * xz;
    J.m[3] = -a * v.z + b * xy;   J.m[4] = 1.0 - b * (xx + zz); J.m[5] = a * v.x + b * yz;
    J.m[6] = a * v.y + b * xz;    J.m[7] = -a * v.x + b * yz;   J.m[8] = 1.0 - b * (xx + yy);
    return J;
}
LV_HD M3<double> so3_right_jacobian_inverse(const V3<double>& v) {
    const double t2 = v.x * v.x + v.y * v.y + v.z * v.z, t = sqrt(t2);
    double c;  // J_r^-1 = I + 1/2 [v]x + c [v]x^2
    if (t < 0.1) {
        c = 1.0 / 12.0 + t2 * (1.0 / 720.0) + t2 * t2 * (1.0 / 30240.0) + t2 * t2 * t2 * (1.0 / 1209600.0);
    } else {
        double sh, ch;
        lv_sincos(0.5 * t, &sh, &ch);
        c = 1.0 / t2 - ch / (2.0 * t * sh);     // 1/t^2 - (1 + cos t) / (2 t sin t)
    }
    const double xx = v.x * v.x, yy = v.y * v.y, zz = v.z * v.z, xy = v.x * v.y, xz = v.x * v.z, yz = v.y * v.z;
    M3<double> J;
    J.m[0] = 1.0 - c * (yy + zz);   J.m[1] = -0.5 * v.z + c * xy;   J.m[2] = 0.5 * v.y + c * xz;
    J.m[3] = 0.5 * v.z + c * xy;    J.m[4] = 1.0 - c * (xx + zz);   J.m[5] = -0.5 * v.x + c * yz;
    J.m[6] = -0.5 * v.y + c * xz;   J.m[7] = 0.5 * v.x + c * yz;    J.m[8] = 1.0 - c * (xx + yy);
    return J;
}

// shared, per item (plain doubles; lives in shared memory on the device)
struct ItemShared {
    double R[3][9];      // R(theta_i), R(theta_j), R(-gamma_hat)
    double Jr[3][9];     // their right Jacobians
    double N[9], rg[3], JinvG[9], a1[3], a2[3];   // N = Ri^T Rj, rg = Log(E N), a1 = Ri^T y1, a2 = Ri^T y2
    double pa[3], pb[3];  // positions of the two frames
};
LV_HD void load_m3(const double* s, M3<double>* M) {
#pragma unroll
    for (int i = 0; i < 9; ++i) M->m[i] = s[i];
}
LV_HD void store_m3(const M3<double>& M, double* d) {
#pragma unroll
    for (int i = 0; i < 9; ++i) d[i] = M.m[i];
}
LV_HD V3<double> imu_gamma_hat(const double* blob, const double* sa) {
    const double* X = blob;
    const double* J = blob + 15;
    const double dx = sa[12] - X[12], dy = sa[13] - X[13], dz = sa[14] - X[14];
    return v3<double>(X[6] + (J[6 * 15 + 12] * dx + J[6 * 15 + 13] * dy + J[6 * 15 + 14] * dz),
                      X[7] + (J[7 * 15 + 12] * dx + J[7 * 15 + 13] * dy + J[7 * 15 + 14] * dz),
                      X[8] + (J[8 * 15 + 12] * dx + J[8 * 15 + 13] * dy + J[8 * 15 + 14] * dz));
}
// phase 1, slot 0/1/2: rotation + right Jacobian of theta_i / theta_j / -gamma_hat
LV_HD void item_phase1(const double* imu_blob, const double* sa, const double* sb, int slot, ItemShared* S) {
    V3<double> v;
    if (slot == 0) v = load3(sa + 3);
    else if (slot == 1) v = load3(sb + 3);
    else v = imu_blob ? neg(imu_gamma_hat(imu_blob, sa)) : v3<double>(0.0, 0.0, 0.0);
    store_m3(exp_so3(v), S->R[slot]);
    store_m3(so3_right_jacobian(v), S->Jr[slot]);
}
// phase 2 (one lane): IMU rotation residual pieces
LV_HD void item_phase2(const Consts& C, const double* imu_blob, const double* sa, const double* sb, int slot, ItemShared* S) {
    M3<double> Ri, Rj;
    load_m3(S->R[0], &Ri);
    load_m3(S->R[1], &Rj);
    if (slot == 0) {
        M3<double> E;
        load_m3(S->R[2], &E);
        const M3<double> N = mul_tn(Ri, Rj);
        const V3<double> rg = log_so3(mul(E, N));
        store_m3(N, S->N);
        S->rg[0] = rg.x; S->rg[1] = rg.y; S->rg[2] = rg.z;
        store_m3(so3_right_jacobian_inverse(rg), S->JinvG);
        const double Dt = imu_blob ? imu_blob[465] : 0.0;
        const double gz = C.g;
        const V3<double> y1 = v3<double>(sb[0] - sa[0] - sa[6] * Dt, sb[1] - sa[1] - sa[7] * Dt, sb[2] - sa[2] + 0.5 * gz * Dt * Dt - sa[8] * Dt);
        const V3<double> y2 = v3<double>(sb[6] - sa[6], sb[7] - sa[7], sb[8] + gz * Dt - sa[8]);
        const V3<double> a1 = mul_t(Ri, y1), a2 = mul_t(Ri, y2);
        S->a1[0] = a1.x; S->a1[1] = a1.y; S->a1[2] = a1.z;
        S->a2[0] = a2.x; S->a2[1] = a2.y; S->a2[2] = a2.z;
        S->pa[0] = sa[0]; S->pa[1] = sa[1]; S->pa[2] = sa[2];
        S->pb[0] = sb[0]; S->pb[1] = sb[1]; S->pb[2] = sb[2];
    }
}
// R as Dual with derivative R [w]x (w = J_r e_k of that rotation), or zero derivative
LV_HD M3<Dual> seeded_rotation(const double* R, const double* Jr, int k /* 0..2, or -1: no seed */) {
    M3<Dual> D;
    if (k < 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) D.m[i] = Dual(R[i]);
        return D;
    }
    const double wx = Jr[k], wy = Jr[3 + k], wz = Jr[6 + k];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double r0 = R[r * 3], r1 = R[r * 3 + 1], r2 = R[r * 3 + 2];
        D.m[r * 3] = Dual(r0, r1 * wz - r2 * wy);
        D.m[r * 3 + 1] = Dual(r1, r2 * wx - r0 * wz);
        D.m[r * 3 + 2] = Dual(r2, r0 * wy - r1 * wx);
    }
    return D;
}
// wheel + ground residuals of the item as duals along state column c (c = 30: values only)
LV_HD void item_wheel_ground_dual(const Consts& C, const double* wheel_blob, bool ground, const ItemShared& S, int c, Dual* rw, Dual* rg);
// column c (0..29) of the un-whitened IMU Jacobian (15), the wheel Jacobian (3) and the ground Jacobian (2, frame b)
LV_HD void item_column(const Consts& C, const double* imu_blob, const double* wheel_blob, bool ground, const ItemShared& S, int c,
                       double* col_imu, double* col_wheel, double* col_ground) {
    const int blk = c / 3, k = c % 3;   // 0 p_i 1 q_i 2 v_i 3 ba_i 4 bw_i 5 p_j 6 q_j 7 v_j 8 ba_j 9 bw_j
#pragma unroll
    for (int r = 0; r < 15; ++r) col_imu[r] = 0.0;
    col_wheel[0] = col_wheel[1] = col_wheel[2] = 0.0;
    col_ground[0] = col_ground[1] = 0.0;
    M3<double> Ri, Rj;
    load_m3(S.R[0], &Ri);
    load_m3(S.R[1], &Rj);
    // w = column k of the right Jacobian of this lane's rotation (if any)
    const double* Jr = blk == 1 ? S.Jr[0] : S.Jr[1];
    const V3<double> w = v3<double>(Jr[k], Jr[3 + k], Jr[6 + k]);
    const V3<double> riTk = v3<double>(Ri.m[k * 3], Ri.m[k * 3 + 1], Ri.m[k * 3 + 2]);   // column k of Ri^T = row k of Ri
    if (imu_blob) {
        const double* J = imu_blob + 15;
        const double Dt = imu_blob[465];
        M3<double> JinvG, N;
        load_m3(S.JinvG, &JinvG);
        load_m3(S.N, &N);
        if (blk == 0) { col_imu[0] = riTk.x; col_imu[1] = riTk.y; col_imu[2] = riTk.z; }
        else if (blk == 5) { col_imu[0] = -riTk.x; col_imu[1] = -riTk.y; col_imu[2] = -riTk.z; }
        else if (blk == 2) { col_imu[0] = riTk.x * Dt; col_imu[1] = riTk.y * Dt; col_imu[2] = riTk.z * Dt; col_imu[3] = riTk.x; col_imu[4] = riTk.y; col_imu[5] = riTk.z; }
        else if (blk == 7) { col_imu[3] = -riTk.x; col_imu[4] = -riTk.y; col_imu[5] = -riTk.z; }
        else if (blk == 3) {
#pragma unroll
            for (int r = 0; r < 6; ++r) col_imu[r] = J[r * 15 + 9 + k];
#pragma unroll
            for (int r = 0; r < 3; ++r) col_imu[9 + r] = (r == k) ? -1.0 : 0.0;   // static indices keep the column in registers
        } else if (blk == 8) {
#pragma unroll
            for (int r = 0; r < 3; ++r) col_imu[9 + r] = (r == k) ? 1.0 : 0.0;
        } else if (blk == 9) {
#pragma unroll
            for (int r = 0; r < 3; ++r) col_imu[12 + r] = (r == k) ? 1.0 : 0.0;
        }
        else if (blk == 1) {
            // r_alpha = .. - Ri^T y1: d(Ri^T y)/d theta = [Ri^T y]x J_r  ->  column = w x (Ri^T y)
            const V3<double> a1 = load3(S.a1), a2 = load3(S.a2);
            const V3<double> c1 = cross(w, a1), c2 = cross(w, a2);
            col_imu[0] = c1.x; col_imu[1] = c1.y; col_imu[2] = c1.z;
            col_imu[3] = c2.x; col_imu[4] = c2.y; col_imu[5] = c2.z;
            const V3<double> g = mul(JinvG, mul_t(N, w));
            col_imu[6] = -g.x; col_imu[7] = -g.y; col_imu[8] = -g.z;
        } else if (blk == 6) {
            const V3<double> g = mul(JinvG, w);
            col_imu[6] = g.x; col_imu[7] = g.y; col_imu[8] = g.z;
        } else {  // blk == 4, bw_i
#pragma unroll
            for (int r = 0; r < 6; ++r) col_imu[r] = J[r * 15 + 12 + k];
            M3<double> JrG;
            load_m3(S.Jr[2], &JrG);
            const V3<double> dg = v3<double>(J[6 * 15 + 12 + k], J[7 * 15 + 12 + k], J[8 * 15 + 12 + k]);
            const V3<double> g = mul(JinvG, mul_t(N, mul(JrG, dg)));
            col_imu[6] = -g.x; col_imu[7] = -g.y; col_imu[8] = -g.z;
#pragma unroll
            for (int r = 0; r < 3; ++r) col_imu[12 + r] = (r == k) ? -1.0 : 0.0;
        }
    }
    // wheel and ground involve the extrinsic T_imu_to_wheel, whose rotation block is only approximately orthonormal
    // (it goes through a non-normalised quaternion at load time, params.cpp:52): differentiate the actual formulas
    // with duals, seeded with the exact dR/dtheta_k = R [J_r e_k]x of the two (orthonormal) frame rotations.
    if (wheel_blob || ground) {
        Dual rw[3], rgd[2];
        item_wheel_ground_dual(C, wheel_blob, ground, S, c, rw, rgd);
#pragma unroll
        for (int r = 0; r < 3; ++r) col_wheel[r] = wheel_blob ? rw[r].d : 0.0;
        col_ground[0] = ground ? rgd[0].d : 0.0;
        col_ground[1] = ground ? rgd[1].d : 0.0;
    }
}
// the raw residual values of the item
LV_HD void item_values(const Consts& C, const double* imu_blob, const double* wheel_blob, bool ground, const ItemShared& S,
                       const double* sa, const double* sb, double* r_imu, double* r_wheel, double* r_ground) {
    if (imu_blob) {
        const double* X = imu_blob;
        const double* J = imu_blob + 15;
        const double dax = sa[9] - X[9], day = sa[10] - X[10], daz = sa[11] - X[11];
        const double dwx = sa[12] - X[12], dwy = sa[13] - X[13], dwz = sa[14] - X[14];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double ab = X[k] + (J[k * 15 + 12] * dwx + J[k * 15 + 13] * dwy + J[k * 15 + 14] * dwz) +
                              (J[k * 15 + 9] * dax + J[k * 15 + 10] * day + J[k * 15 + 11] * daz);
            r_imu[k] = ab - (k < 3 ? S.a1[k] : S.a2[k - 3]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { r_imu[6 + k] = S.rg[k]; r_imu[9 + k] = sb[9 + k] - sa[9 + k]; r_imu[12 + k] = sb[12 + k] - sa[12 + k]; }
    }
    if (wheel_blob || ground) {
        Dual rw[3], rgd[2];
        item_wheel_ground_dual(C, wheel_blob, ground, S, 30, rw, rgd);
        if (wheel_blob) { r_wheel[0] = rw[0].a; r_wheel[1] = rw[1].a; r_wheel[2] = rw[2].a; }
        if (ground) { r_ground[0] = rgd[0].a; r_ground[1] = rgd[1].a; }
    }
}
LV_HD void item_wheel_ground_dual(const Consts& C, const double* wheel_blob, bool ground, const ItemShared& S, int c, Dual* rw, Dual* rg) {
    const int blk = c / 3, k = c % 3;
    const M3<Dual> Ri = seeded_rotation(S.R[0], S.Jr[0], blk == 1 ? k : -1);
    const M3<Dual> Rj = seeded_rotation(S.R[1], S.Jr[1], blk == 6 ? k : -1);
    const V3<Dual> pa = v3<Dual>(Dual(S.pa[0], c == 0 ? 1.0 : 0.0), Dual(S.pa[1], c == 1 ? 1.0 : 0.0), Dual(S.pa[2], c == 2 ? 1.0 : 0.0));
    const V3<Dual> pb = v3<Dual>(Dual(S.pb[0], c == 15 ? 1.0 : 0.0), Dual(S.pb[1], c == 16 ? 1.0 : 0.0), Dual(S.pb[2], c == 17 ? 1.0 : 0.0));
    if (wheel_blob) item_wheel<Dual>(C, wheel_blob, pa, pb, Ri, Rj, rw);
    if (ground) item_ground<Dual>(C, pb, Rj, rg);
}

// ------------------------------------------------------------------ laser frame table
// C = P (R(theta_j) (R_il c + t_il) + p_j) restricted to xy is the affine map  C = M c + t  on the 2-D
// scan point; d C / d theta_k = B_k c + b_k (SURVEY.md verification note).  Layout (24 doubles):
//   [M00 M01 M10 M11 t0 t1 | B0(4) b0(2) | B1(4) b1(2) | B2(4) b2(2)]
constexpr int kFrameTab = 24;
LV_HD void laser_frame_table(const Consts& C, const double* pose /*p q*/, double* tab) {
    const Iso Til = load_iso(C.T_il);
    const V3<double> th = load3(pose + 3);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        V3<Dual> thd = lift<Dual>(th);
        (k == 0 ? thd.x : (k == 1 ? thd.y : thd.z)).d = 1.0;
        const M3<Dual> R = exp_so3(thd);
        // rows 0,1 of R * R_il (columns 0,1) and of R * t_il
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            Dual m0 = R.m[r * 3] * Til.R.m[0] + R.m[r * 3 + 1] * Til.R.m[3] + R.m[r * 3 + 2] * Til.R.m[6];
            Dual m1 = R.m[r * 3] * Til.R.m[1] + R.m[r * 3 + 1] * Til.R.m[4] + R.m[r * 3 + 2] * Til.R.m[7];
            Dual tt = R.m[r * 3] * Til.t.x + R.m[r * 3 + 1] * Til.t.y + R.m[r * 3 + 2] * Til.t.z;
            tab[6 + 6 * k + 2 * r] = m0.d;
            tab[6 + 6 * k + 2 * r + 1] = m1.d;
            tab[6 + 6 * k + 4 + r] = tt.d;
            if (k == 0) {
                tab[2 * r] = m0.a;
                tab[2 * r + 1] = m1.a;
                tab[4 + r] = tt.a + pose[r];
            }
        }
    }
}

}  // namespace lv
