// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED.
//
// factors.hpp — the reference's residual functors restated on the oracle's own Vec3/Mat3/Jet types,
// keeping the Ceres functor convention `bool operator()(const T* const..., T* res) const`:
//   laser_factor            src/factor/laser_factor.h:26-89
//   imu_factor              src/factor/imu_factor.h:13-89
//   wheel_odom_factor       src/factor/wheel_factor.h:12-73
//   ground_factor_p / _q    src/factor/ground_factor.h:25-82
//   marginalization_factor  src/factor/marginalization_factor.h:22-53   (camera off: no world points)
//   so3_parameterization    src/factor/factor_common.h:37-53
// and auto_diff::compute_res_and_jacobi (src/utilies/common.h:201-217) on Jet<N>.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <initializer_list>
#include <vector>

#include "../include/lvio2d.h"
#include "lie.hpp"

namespace oracle {

// the PARAM() values + noise singletons the functors read (laser_noise laser_factor.h:19-24,
// ground_noise ground_factor.h:18-22)
struct Params {
    Iso3<double> T_imu_to_laser, T_imu_to_wheel;
    double g;
    double laser_sqrt_info;       // 1 / line_to_line_sigma
    double manifold_p_sqrt_info;  // 1 / manifold_p_sigma
    double manifold_q_sqrt_info;  // 1 / manifold_q_sigma
    double Q[12];                 // diagonal of imu_noise::Q: na, nw, nba, nbw (imu_preintegraption.h:30-42)
    double wheel_cov[3];          // diagonal of wheel_noise::wheel_cov
    double huber_delta;           // <= 0: no loss (the reference)
    int assoc_mode;               // 0 fixed (the reference), 1 nearest line (BASELINE config 3, an extension)
    double assoc_gate, assoc_max_dist;
    bool analytic_laser = false;  // cpu_baseline flavour: closed-form Jacobian of the point-to-line factor instead of Jets
    explicit Params(const lvio2d_params& p) {
        T_imu_to_laser = iso_from_rowmajor_3x4(p.T_imu_to_laser);
        T_imu_to_wheel = iso_from_rowmajor_3x4(p.T_imu_to_wheel);
        g = p.g;
        laser_sqrt_info = 1.0 / p.line_to_line_sigma;
        manifold_p_sqrt_info = 1.0 / p.manifold_p_sigma;
        manifold_q_sqrt_info = 1.0 / p.manifold_q_sigma;
        huber_delta = (p.huber_delta > 0 && std::isfinite(p.huber_delta)) ? p.huber_delta : 0.0;
        assoc_mode = p.assoc_mode;
        assoc_gate = p.assoc_gate > 0 ? p.assoc_gate : 0.1;
        assoc_max_dist = p.assoc_max_dist > 0 ? p.assoc_max_dist : 0.5;
        for (int i = 0; i < 3; ++i) {
            Q[0 + i] = p.imu_noise_acc_sigma[i] * p.imu_noise_acc_sigma[i];
            Q[3 + i] = p.imu_noise_gyro_sigma[i] * p.imu_noise_gyro_sigma[i];
            Q[6 + i] = p.imu_bias_acc_sigma[i] * p.imu_bias_acc_sigma[i];
            Q[9 + i] = p.imu_bias_gyro_sigma[i] * p.imu_bias_gyro_sigma[i];
            wheel_cov[i] = p.wheel_sigma[i] * p.wheel_sigma[i];
        }
    }
};

template <class T> inline Vec3<T> map3(const T* p) { return {p[0], p[1], p[2]}; }

// ---------------------------------------------------------------- laser (laser_factor.h:26-89)
struct laser_factor {
    const Params* P;
    Vec3<double> l1_p1, l1_p2, l2_p1, l2_p2;
    double len1, len2, sum;
    laser_factor(const Params* P_, const Vec3<double>& a1, const Vec3<double>& a2, const Vec3<double>& c1,
                 const Vec3<double>& c2)
        : P(P_), l1_p1(a1), l1_p2(a2), l2_p1(c1), l2_p2(c2) {
        len1 = norm(l1_p1 - l1_p2);
        len2 = norm(l2_p1 - l2_p2);
        double tmp = std::min(len1, len2);
        sum = tmp / 2.0 / 0.02;
        sum = std::sqrt(sum);
    }
    template <class T>
    bool operator()(const T* const p_w_i, const T* const theta_w_i, const T* const p_w_j, const T* const theta_w_j,
                    T* res) const {
        Iso3<T> T_i_l = cast_iso<T>(P->T_imu_to_laser);
        Iso3<T> T_w_i = lie::make_tf<T>(map3(p_w_i), map3(theta_w_i)) * T_i_l;
        Iso3<T> T_w_j = lie::make_tf<T>(map3(p_w_j), map3(theta_w_j)) * T_i_l;
        Vec3<T> l2_point1 = T_w_j * cast_vec<T>(l2_p1);
        Vec3<T> l2_point2 = T_w_j * cast_vec<T>(l2_p2);
        Vec3<T> l1_point1 = T_w_i * cast_vec<T>(l1_p1);
        Vec3<T> l1_point2 = T_w_i * cast_vec<T>(l1_p2);
        l2_point1.z = T(0.0); l2_point2.z = T(0.0);
        l1_point1.z = T(0.0); l1_point2.z = T(0.0);
        T dis1 = e_laser::dis_from_line<T>(l2_point1, l1_point1, l1_point2);
        T dis2 = e_laser::dis_from_line<T>(l2_point2, l1_point1, l1_point2);
        T e1 = T(P->laser_sqrt_info) * dis1;
        T e2 = T(P->laser_sqrt_info) * dis2;
        res[0] = T(sum) * e1;
        res[1] = T(sum) * e2;
        return true;
    }
};

// One point of a scan against one line: the same expression as ONE residual of laser_factor with an
// explicit weight in place of laser_factor::sum ("beam mode").
struct laser_point_factor {
    const Params* P;
    Vec3<double> a1, a2, c;
    double weight;
    laser_point_factor(const Params* P_, const Vec3<double>& a1_, const Vec3<double>& a2_, const Vec3<double>& c_, double w)
        : P(P_), a1(a1_), a2(a2_), c(c_), weight(w) {}
    template <class T>
    bool operator()(const T* const p_w_i, const T* const theta_w_i, const T* const p_w_j, const T* const theta_w_j,
                    T* res) const {
        Iso3<T> T_i_l = cast_iso<T>(P->T_imu_to_laser);
        Iso3<T> T_w_i = lie::make_tf<T>(map3(p_w_i), map3(theta_w_i)) * T_i_l;
        Iso3<T> T_w_j = lie::make_tf<T>(map3(p_w_j), map3(theta_w_j)) * T_i_l;
        Vec3<T> C = T_w_j * cast_vec<T>(c);
        Vec3<T> A1 = T_w_i * cast_vec<T>(a1);
        Vec3<T> A2 = T_w_i * cast_vec<T>(a2);
        C.z = T(0.0); A1.z = T(0.0); A2.z = T(0.0);
        T dis = e_laser::dis_from_line<T>(C, A1, A2);
        T e = T(P->laser_sqrt_info) * dis;
        res[0] = T(weight) * e;
        return true;
    }
};

// The same residual with a closed-form 1x12 Jacobian (no Jets) — the "analytic" CPU-baseline flavour SURVEY.md §8d asks
// to be timed next to the faithful Jet flavour (the reference itself only has the Jet path).  With
// R(theta + d) = R(theta) Exp(J_r(theta) d):  d(R v)/d theta = -R [v]x J_r(theta).
inline Mat3<double> so3_right_jacobian(const Vec3<double>& v) {
    const double t2 = v.x * v.x + v.y * v.y + v.z * v.z, t = std::sqrt(t2);
    double a, b;  // J_r = I - a [v]x + b [v]x^2
    if (t < 1e-6) { a = 0.5 - t2 / 24.0; b = 1.0 / 6.0 - t2 / 120.0; }
    else { a = (1.0 - std::cos(t)) / t2; b = (t - std::sin(t)) / (t2 * t); }
    Mat3<double> K;
    K.m[0][1] = -v.z; K.m[0][2] = v.y; K.m[1][0] = v.z; K.m[1][2] = -v.x; K.m[2][0] = -v.y; K.m[2][1] = v.x;
    const Mat3<double> K2 = K * K;
    Mat3<double> J = Mat3<double>::identity();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J.m[i][j] += -a * K.m[i][j] + b * K2.m[i][j];
    return J;
}
// d(R(theta) v)/d theta as a 3x3 matrix
inline Mat3<double> d_rotated_d_theta(const Mat3<double>& R, const Vec3<double>& theta, const Vec3<double>& v) {
    Mat3<double> Vx;
    Vx.m[0][1] = -v.z; Vx.m[0][2] = v.y; Vx.m[1][0] = v.z; Vx.m[1][2] = -v.x; Vx.m[2][0] = -v.y; Vx.m[2][1] = v.x;
    Mat3<double> M = R * Vx * so3_right_jacobian(theta);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M.m[i][j] = -M.m[i][j];
    return M;
}
// res[0] and jac[12] over (p_i, theta_i, p_j, theta_j) of laser_point_factor
inline void laser_point_analytic(const Params& P, const Vec3<double>& a1, const Vec3<double>& a2, const Vec3<double>& c, double weight,
                                 const double* xi, const double* xj, double* res, double* jac) {
    const Vec3<double> thi(xi[3], xi[4], xi[5]), thj(xj[3], xj[4], xj[5]);
    const Mat3<double> Ri = lie::exp_so3<double>(thi), Rj = lie::exp_so3<double>(thj);
    const Mat3<double>& Ril = P.T_imu_to_laser.R;
    const Vec3<double>& til = P.T_imu_to_laser.t;
    const Vec3<double> cj = Ril * c + til, b1 = Ril * a1 + til, b2 = Ril * a2 + til;   // in the IMU frames
    Vec3<double> C = Rj * cj + Vec3<double>(xj[0], xj[1], xj[2]);
    Vec3<double> A1 = Ri * b1 + Vec3<double>(xi[0], xi[1], xi[2]), A2 = Ri * b2 + Vec3<double>(xi[0], xi[1], xi[2]);
    // flattened to z = 0: only the x, y rows of every derivative matter
    const double ex = A2.x - A1.x, ey = A2.y - A1.y, L = std::sqrt(ex * ex + ey * ey);
    const double ux = ex / L, uy = ey / L, nx = -uy, ny = ux;
    const double wx = C.x - A2.x, wy = C.y - A2.y;
    const double d = nx * wx + ny * wy;
    const double sgn = d < 0 ? -1.0 : 1.0, k = weight * P.laser_sqrt_info;
    res[0] = k * std::fabs(d);
    const Mat3<double> dC = d_rotated_d_theta(Rj, thj, cj), dA1 = d_rotated_d_theta(Ri, thi, b1), dA2 = d_rotated_d_theta(Ri, thi, b2);
    const double t = ux * wx + uy * wy;   // along-line coordinate of C relative to A2
    for (int q = 0; q < 3; ++q) {
        // pose j: only C moves
        jac[6 + q] = k * sgn * (q == 0 ? nx : (q == 1 ? ny : 0.0));
        jac[9 + q] = k * sgn * (nx * dC.m[0][q] + ny * dC.m[1][q]);
        // pose i: A1 and A2 move.  d = n.(C - A2), n = perp(u), u = (A2 - A1)/L:
        //   dd = -n.dA2 + dn.(C - A2),  dn.(C - A2) = -(u.(C - A2)) n.d(A2 - A1)/L
        const double da1x = (q == 0), da1y = (q == 1);      // d A /d p_i = I
        jac[q] = k * sgn * (-(nx * da1x + ny * da1y));       // d(A2 - A1)/d p_i = 0
        const double dex = dA2.m[0][q] - dA1.m[0][q], dey = dA2.m[1][q] - dA1.m[1][q];
        jac[3 + q] = k * sgn * (-(nx * dA2.m[0][q] + ny * dA2.m[1][q]) - t * (nx * dex + ny * dey) / L);
    }
}

// ---------------------------------------------------------------- imu (imu_factor.h:13-89)
struct imu_factor {
    const Params* P;
    const double* blob;  // X[15] | J[225] | sqrt_inverse_P[225] | Dt  (imu_preint_result)
    imu_factor(const Params* P_, const double* blob_) : P(P_), blob(blob_) {}
    template <class T>
    bool operator()(const T* const p_w_i, const T* const theta_w_i, const T* const v_w_i, const T* const bs_w_i,
                    const T* const p_w_j, const T* const theta_w_j, const T* const v_w_j, const T* const bs_w_j,
                    T* res) const {
        const double* X = blob;
        const double* J = blob + 15;
        const double* S = blob + 15 + 225;
        Vec3<T> pi = map3(p_w_i), vi = map3(v_w_i), thetai = map3(theta_w_i), bai = map3(bs_w_i), bwi = map3(bs_w_i + 3);
        Vec3<T> pj = map3(p_w_j), vj = map3(v_w_j), thetaj = map3(theta_w_j), baj = map3(bs_w_j), bwj = map3(bs_w_j + 3);
        T g_norm = T(P->g);
        Vec3<T> g(T(0.0), T(0.0), T(1.0));
        Vec3<T> alpha{T(X[0]), T(X[1]), T(X[2])};
        Vec3<T> beta{T(X[3]), T(X[4]), T(X[5])};
        Vec3<T> gamma{T(X[6]), T(X[7]), T(X[8])};
        Vec3<T> ba{T(X[9]), T(X[10]), T(X[11])};
        Vec3<T> bw{T(X[12]), T(X[13]), T(X[14])};
        T Dt = T(blob[465]);
        Mat3<T> bk_R_w = lie::exp_so3<T>(-thetai);
        auto Jblock = [&](int r0, int c0) {
            Mat3<T> m;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m.m[r][c] = T(J[(r0 + r) * 15 + c0 + c]);
            return m;
        };
        Mat3<T> alpha_J_ba = Jblock(0, 9), alpha_J_bw = Jblock(0, 12);
        Mat3<T> beta_J_ba = Jblock(3, 9), beta_J_bw = Jblock(3, 12);
        Mat3<T> gamma_J_bw = Jblock(6, 12);
        alpha = alpha + alpha_J_ba * (bai - ba) + alpha_J_bw * (bwi - bw);
        beta = beta + beta_J_ba * (bai - ba) + beta_J_bw * (bwi - bw);
        gamma = gamma + gamma_J_bw * (bwi - bw);
        T r[15];
        // 0.5 * g * g_norm * Dt * Dt evaluates left to right
        Vec3<T> half_g = ((g * T(0.5)) * g_norm * Dt) * Dt;
        Vec3<T> res_alpha = alpha - bk_R_w * (pj - pi + half_g - vi * Dt);
        Vec3<T> res_beta = beta - bk_R_w * (vj + (g * g_norm) * Dt - vi);
        Vec3<T> res_gamma =
            lie::log_SO3<T>(lie::exp_so3<T>(-gamma) * (lie::exp_so3<T>(-thetai) * lie::exp_so3<T>(thetaj)));
        Vec3<T> res_ba = baj - bai;
        Vec3<T> res_bw = bwj - bwi;
        for (int k = 0; k < 3; ++k) {
            r[0 + k] = res_alpha[k]; r[3 + k] = res_beta[k]; r[6 + k] = res_gamma[k];
            r[9 + k] = res_ba[k];    r[12 + k] = res_bw[k];
        }
        for (int i = 0; i < 15; ++i) {
            T acc(0.0);
            for (int k = 0; k < 15; ++k) acc = acc + T(S[i * 15 + k]) * r[k];
            res[i] = acc;
        }
        return true;
    }
};

// ---------------------------------------------------------------- wheel (wheel_factor.h:12-73)
struct wheel_odom_factor {
    const Params* P;
    Iso3<double> delta_Tij;
    double s00, s11, s22;  // diag of sqrt_inverse_P
    wheel_odom_factor(const Params* P_, const double* blob) : P(P_) {
        delta_Tij = iso_from_rowmajor_3x4(blob);
        s00 = blob[12]; s11 = blob[13]; s22 = blob[14];
    }
    template <class T>
    bool operator()(const T* const p_w_i, const T* const theta_w_i, const T* const p_w_j, const T* const theta_w_j,
                    T* res) const {
        Iso3<T> T_i_w = cast_iso<T>(P->T_imu_to_wheel);
        Iso3<T> tf_i = lie::make_tf<T>(map3(p_w_i), map3(theta_w_i)) * T_i_w;
        Iso3<T> tf_j = lie::make_tf<T>(map3(p_w_j), map3(theta_w_j)) * T_i_w;
        Iso3<T> w_tf_ij = inverse(tf_i) * tf_j;
        Vec3<T> p, q, op, oq;
        lie::log_SE3(w_tf_ij, p, q);
        lie::log_SE3(cast_iso<T>(delta_Tij), op, oq);
        T o_len = sqrt(op.x * op.x + op.y * op.y);
        T len = sqrt(p.x * p.x + p.y * p.y);
        Vec3<T> o_dir(op.x, op.y, T(0.0));
        Vec3<T> dir(p.x, p.y, T(0.0));
        T angle = T(0.0);
        if (norm(o_dir) > T(0.0001) && norm(dir) > T(0.0001)) {
            o_dir = normalized(o_dir);
            dir = normalized(dir);
            T sinn = norm(cross(o_dir, dir));
            angle = asin(sinn);
        } else {
            angle = norm(dir);
        }
        if (len < T(0.0001) || o_len < T(0.0001))
            res[0] = T(s00) * len;
        else
            res[0] = T(s00) * (o_len - len);
        res[1] = T(s11) * angle;
        if (norm(q) < T(0.001) || norm(oq) < T(0.001))
            res[2] = T(s22) * norm(q);
        else
            res[2] = T(s22) * (norm(oq) - norm(q));
        return true;
    }
};

// ---------------------------------------------------------------- ground (ground_factor.h:25-82)
struct ground_factor_p {
    const Params* P;
    explicit ground_factor_p(const Params* P_) : P(P_) {}
    template <class T> bool operator()(const T* const p_w_i, const T* const theta_w_i, T* res) const {
        Iso3<T> tf_w_i = lie::make_tf<T>(map3(p_w_i), map3(theta_w_i));
        Iso3<T> tf_w_o = tf_w_i * cast_iso<T>(P->T_imu_to_wheel);
        T dis_from_plane = tf_w_o.t.z;
        res[0] = T(P->manifold_p_sqrt_info) * dis_from_plane;
        return true;
    }
};
struct ground_factor_q {
    const Params* P;
    explicit ground_factor_q(const Params* P_) : P(P_) {}
    template <class T> bool operator()(const T* const p_w_i, const T* const theta_w_i, T* res) const {
        Iso3<T> tf_w_i = lie::make_tf<T>(map3(p_w_i), map3(theta_w_i));
        Iso3<T> tf_w_o = tf_w_i * cast_iso<T>(P->T_imu_to_wheel);
        Vec3<T> ABC(T(0.0), T(0.0), T(1.0));
        Vec3<T> z_axis(tf_w_o.R.m[0][2], tf_w_o.R.m[1][2], tf_w_o.R.m[2][2]);
        T sinn = norm(cross(z_axis, ABC));
        T angle = asin(sinn);
        res[0] = T(P->manifold_q_sqrt_info) * angle;
        return true;
    }
};

// ---------------------------------------------------------------- prior (marginalization_factor.h:22-53)
struct marginalization_factor {
    const double* X0;  // [15]
    const double* J;   // [15][15] row-major
    marginalization_factor(const double* X0_, const double* J_) : X0(X0_), J(J_) {}
    template <class T>
    bool operator()(const T* const p, const T* const q, const T* const v, const T* const bs, T* res) const {
        T X[15];
        for (int i = 0; i < 3; ++i) { X[i] = p[i]; X[3 + i] = q[i]; X[6 + i] = v[i]; }
        for (int i = 0; i < 6; ++i) X[9 + i] = bs[i];
        for (int r = 0; r < 15; ++r) {
            T acc(0.0);
            // linearized_R is NOT added (marginalization_factor.h:50)
            for (int c = 0; c < 15; ++c) acc = acc + T(J[r * 15 + c]) * (X[c] - T(X0[c]));
            res[r] = acc;
        }
        return true;
    }
};

// so3_parameterization::operator() on doubles (factor_common.h:40-53): Plus(theta, d) = wrap(theta + d)
inline void so3_plus(const double* theta, const double* delta, double* out) {
    Vec3<double> tmp(theta[0] + delta[0], theta[1] + delta[1], theta[2] + delta[2]);
    lie::normalize_so3(tmp);
    out[0] = tmp.x; out[1] = tmp.y; out[2] = tmp.z;
}

// ------------------------------------------------ auto_diff::compute_res_and_jacobi (common.h:201-217)
// params: one pointer per parameter block, sizes: block sizes.  jac is row-major [NR][NP] with the
// blocks concatenated column-wise (NP = sum of sizes).
template <int NR, int NP, class Functor, class Call>
inline void autodiff(const Functor& f, std::initializer_list<const double*> params, std::initializer_list<int> sizes, double* res,
                     double* jac, Call call) {
    Jet<NP> x[NP];
    const Jet<NP>* ptrs[8];
    int k = 0, b = 0;
    auto sz = sizes.begin();
    for (const double* p : params) {
        ptrs[b++] = &x[k];
        for (int i = 0; i < *sz; ++i, ++k) x[k] = Jet<NP>(p[i], k);
        ++sz;
    }
    Jet<NP> r[NR];
    call(f, ptrs, r);
    for (int i = 0; i < NR; ++i) {
        res[i] = r[i].a;
        if (jac) for (int c = 0; c < NP; ++c) jac[i * NP + c] = r[i].v[c];
    }
}

}  // namespace oracle
