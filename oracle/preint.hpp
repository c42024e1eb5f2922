// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED.
//
// preint.hpp — restatement of the two preintegrators that produce the IMU / wheel factor constants:
//   imu_preintegraption::{reset_imu_measure, update, get_preintegraption_result}
//       src/factor/imu_preintegraption.h:113-124, :170-208, :147-152
//   wheel_odom_preintegration::{reset_wheel_odom_measure, update_by_v, get_preintegraption_result}
//       src/factor/wheel_odom_preintegration.h:52-61, :141-152, :111-125
// Dense 15x15 helpers stand in for Eigen (LLT, PartialPivLU inverse).
#pragma once
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "factors.hpp"

namespace oracle {

// ---- tiny dense helpers on row-major n x n arrays
inline void matmul(const double* A, const double* B, double* C, int n, int m, int k) {  // C[n][k] = A[n][m] B[m][k]
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            for (int l = 0; l < m; ++l) s += A[i * m + l] * B[l * k + j];
            C[i * k + j] = s;
        }
}
// inverse by LU with partial pivoting (what Eigen's MatrixBase::inverse() does for n > 4)
inline bool inverse_pplu(const double* A, double* Ainv, int n) {
    std::vector<double> lu(A, A + n * n);
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = std::fabs(lu[k * n + k]);
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(lu[i * n + k]) > best) { best = std::fabs(lu[i * n + k]); piv = i; }
        if (best == 0.0) return false;
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(lu[k * n + j], lu[piv * n + j]);
            std::swap(perm[k], perm[piv]);
        }
        for (int i = k + 1; i < n; ++i) {
            lu[i * n + k] /= lu[k * n + k];
            const double f = lu[i * n + k];
            for (int j = k + 1; j < n; ++j) lu[i * n + j] -= f * lu[k * n + j];
        }
    }
    // solve LU X = P I column by column
    std::vector<double> y(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) {
            double s = (perm[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s -= lu[i * n + j] * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < n; ++j) s -= lu[i * n + j] * Ainv[j * n + c];
            Ainv[i * n + c] = s / lu[i * n + i];
        }
    }
    return true;
}
// lower Cholesky A = L L^T (Eigen::LLT); returns false when a pivot is not positive
inline bool cholesky_lower(const double* A, double* L, int n) {
    std::memset(L, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0.0)) return false;
        const double ljj = std::sqrt(d);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / ljj;
        }
    }
    return true;
}

// ---------------------------------------------------------------- IMU
struct imu_preintegraption {
    double J[225], Pm[225], X[15], Dt;
    const Params* P;
    explicit imu_preintegraption(const Params* P_) : P(P_) { const double z[3] = {0, 0, 0}; reset(z, z); }
    // reset_imu_measure (imu_preintegraption.h:113-124)
    void reset(const double* acc_bias, const double* gyr_bias) {
        std::memset(J, 0, sizeof(J)); std::memset(Pm, 0, sizeof(Pm)); std::memset(X, 0, sizeof(X));
        for (int i = 0; i < 15; ++i) { J[i * 15 + i] = 1.0; Pm[i * 15 + i] = 0.00001; }
        for (int i = 0; i < 3; ++i) { X[9 + i] = acc_bias[i]; X[12 + i] = gyr_bias[i]; }
        Dt = 0;
    }
    // update(dt) with hat_acc/hat_gyro = last_info (imu_preintegraption.h:170-208)
    void update(double dt, const double* acc, const double* gyro) {
        Vec3<double> last_alpha(X[0], X[1], X[2]), last_beta(X[3], X[4], X[5]), last_gamma(X[6], X[7], X[8]);
        Vec3<double> last_ba(X[9], X[10], X[11]), last_bw(X[12], X[13], X[14]);
        Mat3<double> last_Rz = lie::exp_so3<double>(last_gamma);
        Vec3<double> hat_acc(acc[0], acc[1], acc[2]), hat_gyro(gyro[0], gyro[1], gyro[2]);
        Vec3<double> a_unb = hat_acc - last_ba;
        Vec3<double> Ra = last_Rz * a_unb;
        Vec3<double> alpha = last_alpha + last_beta * dt + ((Ra * 0.5) * dt) * dt;
        Vec3<double> beta = last_beta + Ra * dt;
        Vec3<double> gamma = lie::log_SO3<double>(lie::exp_so3(last_gamma) * lie::exp_so3<double>((hat_gyro - last_bw) * dt));
        for (int i = 0; i < 3; ++i) { X[i] = alpha[i]; X[3 + i] = beta[i]; X[6 + i] = gamma[i]; }
        double F[225] = {0};
        auto setblk = [&](double* M, int r0, int c0, const Mat3<double>& B) {
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[(r0 + r) * 15 + c0 + c] = B.m[r][c];
        };
        Mat3<double> I3 = Mat3<double>::identity();
        setblk(F, 0, 3, I3);
        setblk(F, 3, 6, -(last_Rz * convert::cross_matrix<double>(a_unb)));
        setblk(F, 3, 9, -last_Rz);
        // imu_preintegraption.h:192 subtracts last_ba (not last_bw) from the gyro — reproduced
        setblk(F, 6, 6, -convert::cross_matrix<double>(hat_gyro - last_ba));
        setblk(F, 6, 12, -I3);
        double G[15 * 12] = {0};
        auto setg = [&](int r0, int c0, const Mat3<double>& B) {
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) G[(r0 + r) * 12 + c0 + c] = B.m[r][c];
        };
        setg(3, 0, -last_Rz);
        setg(6, 3, -I3);
        setg(9, 6, I3);
        setg(12, 9, I3);
        for (int i = 0; i < 225; ++i) F[i] *= dt;
        for (int i = 0; i < 15; ++i) F[i * 15 + i] += 1.0;
        double tmp[225], tmp2[225];
        matmul(F, J, tmp, 15, 15, 15);
        std::memcpy(J, tmp, sizeof(J));
        // P = F P F^T + (G dt) Q (G dt)^T
        matmul(F, Pm, tmp, 15, 15, 15);
        for (int i = 0; i < 15; ++i)
            for (int j = 0; j < 15; ++j) {
                double s = 0.0;
                for (int l = 0; l < 15; ++l) s += tmp[i * 15 + l] * F[j * 15 + l];
                tmp2[i * 15 + j] = s;
            }
        for (int i = 0; i < 15; ++i)
            for (int j = 0; j < 15; ++j) {
                double s = 0.0;
                for (int l = 0; l < 12; ++l) s += ((G[i * 12 + l] * dt) * P->Q[l]) * (G[j * 12 + l] * dt);
                Pm[i * 15 + j] = tmp2[i * 15 + j] + s;
            }
        Dt += dt;
    }
    // get_preintegraption_result (imu_preintegraption.h:147-152): sqrt_inverse_P = LLT(P^-1).L^T
    bool result(double* blob) const {
        std::memcpy(blob, X, sizeof(X));
        std::memcpy(blob + 15, J, sizeof(J));
        double Pinv[225], L[225];
        if (!inverse_pplu(Pm, Pinv, 15)) return false;
        if (!cholesky_lower(Pinv, L, 15)) return false;
        for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) blob[240 + i * 15 + j] = L[j * 15 + i];
        blob[465] = Dt;
        return true;
    }
};

// ---------------------------------------------------------------- wheel
struct wheel_odom_preintegration {
    Iso3<double> delta_Tij;
    double Dt;
    const Params* P;
    explicit wheel_odom_preintegration(const Params* P_) : P(P_) { reset(); }
    void reset() { delta_Tij = Iso3<double>(); Dt = 0; }
    // update_by_v (wheel_odom_preintegration.h:141-152)
    void update_by_v(double dt, const double* v, const double* omega) {
        if (dt <= 0 || dt >= 10) return;
        Dt = Dt + dt;
        Iso3<double> delta_T = lie::make_tf<double>(Vec3<double>(v[0] * dt, v[1] * dt, v[2] * dt),
                                                    Vec3<double>(omega[0] * dt, omega[1] * dt, omega[2] * dt));
        delta_Tij = delta_Tij * delta_T;
    }
    // get_preintegraption_result (wheel_odom_preintegration.h:111-125)
    void result(double* blob) const {
        Vec3<double> dp, dq;
        lie::log_SE3(delta_Tij, dp, dq);
        const double len_norm = std::max(squared_norm(dp), 0.005 * 0.005);
        const double delta_yaw_norm = std::max(squared_norm(dq), 0.005 * 0.005);
        const double k[3] = {len_norm, len_norm, delta_yaw_norm};
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) blob[i * 4 + j] = delta_Tij.R.m[i][j]; blob[i * 4 + 3] = delta_Tij.t[i]; }
        // cov is diagonal: LLT(cov^-1).L^T is diag(1/sqrt(cov_ii))
        for (int i = 0; i < 3; ++i) blob[12 + i] = std::sqrt(1.0 / (P->wheel_cov[i] * k[i]));
    }
};

}  // namespace oracle
