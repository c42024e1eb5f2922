// aux_kernels.cuh — the smaller device kernels around the solver:
//   * IMU / wheel preintegration, one warp (IMU) or one thread (wheel) per inter-frame interval
//       imu_preintegraption::update + get_preintegraption_result   (reference src/factor/imu_preintegraption.h:170-208, :147-152)
//       wheel_odom_preintegration::update_by_v + get_preintegraption_result (src/factor/wheel_odom_preintegration.h:141-152, :111-125)
//   * the tail of solver::marginalization: eigen-decomposition of the Schur complement and the
//     sqrt-information prior (src/factor/solver.cpp:390-428)
//   * single-factor evaluation hooks = auto_diff::compute_res_and_jacobi (src/utilies/common.h:201-217)
#pragma once
#include "window.cuh"

namespace lv {

// in-place lower Cholesky of a 15x15 block in shared memory (lane r owns row r); false when a pivot is not positive
__device__ __forceinline__ bool chol15(double* A, int lane) {
    for (int k = 0; k < 15; ++k) {
        const double d = A[k * 15 + k];
        if (!(d > 0.0) || !isfinite(d)) return false;
        const double l = sqrt(d), inv = 1.0 / l;
        __syncwarp();
        if (lane == k) A[k * 15 + k] = l;
        else if (lane > k && lane < 15) A[lane * 15 + k] *= inv;
        __syncwarp();
        if (lane > k && lane < 15) {
            const double lr = A[lane * 15 + k];
            for (int c = k + 1; c <= lane; ++c) A[lane * 15 + c] -= lr * A[c * 15 + k];
        }
        __syncwarp();
    }
    return true;
}

// ------------------------------------------------------------------ IMU preintegration: one warp per interval
// shared memory per warp: J[225] P[225] F[225] T[225] G[15x12 = 180] X[15] -> 1100 doubles
constexpr int kImuPreSmem = 1104;
__global__ void __launch_bounds__(128) imu_preintegrate_kernel(Consts C, int n_intervals, const int64_t* sample_offset,
                                                               const double* samples, const double* bias0, double* out) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int it = blockIdx.x * (blockDim.x >> 5) + warp;
    if (it >= n_intervals) return;
    double* J = smem + (size_t)warp * kImuPreSmem;
    double* P = J + 225;
    double* F = P + 225;
    double* T = F + 225;
    double* G = T + 225;
    double* X = G + 180;
    // reset_imu_measure (imu_preintegraption.h:113-124)
    for (int e = lane; e < 225; e += 32) { const bool dg = (e / 15) == (e % 15); J[e] = dg ? 1.0 : 0.0; P[e] = dg ? 0.00001 : 0.0; }
    if (lane < 15) X[lane] = lane < 9 ? 0.0 : bias0[(size_t)it * 6 + lane - 9];
    double Dt = 0.0;
    __syncwarp();
    for (int64_t s = sample_offset[it]; s < sample_offset[it + 1]; ++s) {
        const double dt = samples[7 * s];
        const V3<double> acc = load3(samples + 7 * s + 1), gyr = load3(samples + 7 * s + 4);
        const V3<double> al = load3(X), be = load3(X + 3), ga = load3(X + 6), ba = load3(X + 9), bw = load3(X + 12);
        const M3<double> Rz = exp_so3(ga);
        const V3<double> au = acc - ba;
        const V3<double> Ra = mul(Rz, au);
        __syncwarp();
        // F = I + dt * F_c, G (imu_preintegraption.h:188-202); every lane builds its own entries
        for (int e = lane; e < 225; e += 32) F[e] = (e / 15) == (e % 15) ? 1.0 : 0.0;
        for (int e = lane; e < 180; e += 32) G[e] = 0.0;
        __syncwarp();
        if (lane < 9) {
            const int r = lane / 3, c = lane % 3;
            // -Rz [a]x  and  -Rz
            const double ax[9] = {0.0, -au.z, au.y, au.z, 0.0, -au.x, -au.y, au.x, 0.0};
            double v = 0.0;
            for (int k = 0; k < 3; ++k) v += Rz.m[r * 3 + k] * ax[k * 3 + c];
            F[(3 + r) * 15 + 6 + c] = -v * dt;
            F[(3 + r) * 15 + 9 + c] = -Rz.m[r * 3 + c] * dt;
            // imu_preintegraption.h:192 subtracts last_ba from the gyro (not last_bw): reproduced
            const V3<double> wu = gyr - ba;
            const double wx[9] = {0.0, -wu.z, wu.y, wu.z, 0.0, -wu.x, -wu.y, wu.x, 0.0};
            F[(6 + r) * 15 + 6 + c] += -wx[r * 3 + c] * dt;
            if (r == c) {
                F[r * 15 + 3 + c] = dt;
                F[(6 + r) * 15 + 12 + c] = -dt;
                G[(6 + r) * 12 + 3 + c] = -1.0;
                G[(9 + r) * 12 + 6 + c] = 1.0;
                G[(12 + r) * 12 + 9 + c] = 1.0;
            }
            G[(3 + r) * 12 + c] = -Rz.m[r * 3 + c];
        }
        __syncwarp();
        // X update (imu_preintegraption.h:183-185)
        if (lane == 0) {
            const V3<double> gn = log_so3(mul(Rz, exp_so3(scale(gyr - bw, dt))));
            X[0] = al.x + be.x * dt + 0.5 * Ra.x * dt * dt; X[1] = al.y + be.y * dt + 0.5 * Ra.y * dt * dt; X[2] = al.z + be.z * dt + 0.5 * Ra.z * dt * dt;
            X[3] = be.x + Ra.x * dt; X[4] = be.y + Ra.y * dt; X[5] = be.z + Ra.z * dt;
            X[6] = gn.x; X[7] = gn.y; X[8] = gn.z;
        }
        // J = F J
        if (lane < 15) {
            double f[15];
#pragma unroll
            for (int k = 0; k < 15; ++k) f[k] = F[lane * 15 + k];
            for (int c = 0; c < 15; ++c) { double v = 0.0; for (int k = 0; k < 15; ++k) v += f[k] * J[k * 15 + c]; T[lane * 15 + c] = v; }
        }
        __syncwarp();
        for (int e = lane; e < 225; e += 32) J[e] = T[e];
        __syncwarp();   // T is rewritten below (racecheck)
        // T = F P
        if (lane < 15) {
            double f[15];
#pragma unroll
            for (int k = 0; k < 15; ++k) f[k] = F[lane * 15 + k];
            for (int c = 0; c < 15; ++c) { double v = 0.0; for (int k = 0; k < 15; ++k) v += f[k] * P[k * 15 + c]; T[lane * 15 + c] = v; }
        }
        __syncwarp();
        // P = T F^T + (G dt) Q (G dt)^T
        if (lane < 15) {
            double t[15], gr[12];
#pragma unroll
            for (int k = 0; k < 15; ++k) t[k] = T[lane * 15 + k];
#pragma unroll
            for (int k = 0; k < 12; ++k) gr[k] = G[lane * 12 + k] * dt;
            for (int c = 0; c < 15; ++c) {
                double v = 0.0;
                for (int k = 0; k < 15; ++k) v += t[k] * F[c * 15 + k];
                double q = 0.0;
                for (int k = 0; k < 12; ++k) q += (gr[k] * C.Q[k]) * (G[c * 12 + k] * dt);
                P[lane * 15 + c] = v + q;
            }
        }
        Dt += dt;
        __syncwarp();
    }
    // result: X, J, sqrt_inverse_P = chol(P^-1)^T, Dt
    double* o = out + (size_t)it * 466;
    if (lane < 15) o[lane] = X[lane];
    for (int e = lane; e < 225; e += 32) o[15 + e] = J[e];
    // P^-1 through the Cholesky factor of P: P = L L^T, P^-1 = L^-T L^-1
    for (int e = lane; e < 225; e += 32) F[e] = P[e];
    __syncwarp();
    const bool ok = chol15(F, lane);
    // T = L^-1 (lower): column c solved by lane c
    if (lane < 15) {
        for (int r = 0; r < 15; ++r) {
            double s = (r == lane) ? 1.0 : 0.0;
            for (int k = lane; k < r; ++k) s -= F[r * 15 + k] * T[k * 15 + lane];
            T[r * 15 + lane] = (r < lane) ? 0.0 : s / F[r * 15 + r];
        }
    }
    __syncwarp();
    // P^-1 = T^T T
    if (lane < 15) {
        double col[15];
        for (int k = 0; k < 15; ++k) col[k] = T[k * 15 + lane];
        for (int c = 0; c < 15; ++c) { double v = 0.0; for (int k = 0; k < 15; ++k) v += col[k] * T[k * 15 + c]; P[lane * 15 + c] = v; }
    }
    __syncwarp();
    const bool ok2 = chol15(P, lane);
    for (int e = lane; e < 225; e += 32) {
        const int r = e / 15, c = e % 15;
        o[240 + e] = (ok && ok2 && c >= r) ? P[c * 15 + r] : (ok && ok2 ? 0.0 : nan(""));
    }
    if (lane == 0) o[465] = Dt;
}

// ------------------------------------------------------------------ wheel preintegration: one thread per interval
__global__ void wheel_preintegrate_kernel(Consts C, int n_intervals, const int64_t* step_offset, const double* steps, double* out) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_intervals) return;
    M3<double> R;
#pragma unroll
    for (int i = 0; i < 9; ++i) R.m[i] = (i % 4 == 0) ? 1.0 : 0.0;
    V3<double> t = v3<double>(0.0, 0.0, 0.0);
    for (int64_t s = step_offset[it]; s < step_offset[it + 1]; ++s) {
        const double dt = steps[7 * s];
        if (dt <= 0 || dt >= 10) continue;  // wheel_odom_preintegration.h:143-147
        const V3<double> v = scale(load3(steps + 7 * s + 1), dt), w = scale(load3(steps + 7 * s + 4), dt);
        const M3<double> dR = exp_so3(w);
        t = mul(R, v) + t;   // delta_Tij = delta_Tij * make_tf(v dt, omega dt)
        R = mul(R, dR);
    }
    const V3<double> dq = log_so3(R);
    const double len_norm = fmax(dot(t, t), 0.005 * 0.005);
    const double yaw_norm = fmax(dot(dq, dq), 0.005 * 0.005);
    double* o = out + (size_t)it * 15;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) o[i * 4 + j] = R.m[i * 3 + j];
    }
    o[3] = t.x; o[7] = t.y; o[11] = t.z;
    o[12] = sqrt(1.0 / (C.wheel_cov[0] * len_norm));
    o[13] = sqrt(1.0 / (C.wheel_cov[1] * len_norm));
    o[14] = sqrt(1.0 / (C.wheel_cov[2] * yaw_norm));
}

// ------------------------------------------------------------------ marginalisation tail (solver.cpp:390-428)
// cyclic Jacobi eigen-decomposition of the 15x15 Schur complement (one thread per window; not on the per-iteration
// path).  Eigenvalues ascending like Eigen::SelfAdjointEigenSolver; eigenvector sign: largest component positive.
__global__ void marginal_prior_kernel(int n_windows, int n_frames, const double* marg_H, const double* marg_g, const double* x,
                                      double* X0, double* J_lin, double* r_lin) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    double A[225], V[225];
    const double* H = marg_H + (size_t)w * 225;
    for (int a = 0; a < 15; ++a)
        for (int b = 0; b < 15; ++b) {
            A[a * 15 + b] = H[(a >= b ? a : b) * 15 + (a >= b ? b : a)];  // lower triangle, like SelfAdjointEigenSolver
            V[a * 15 + b] = a == b ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int i = 0; i < 15; ++i) { dg += A[i * 15 + i] * A[i * 15 + i]; for (int j = i + 1; j < 15; ++j) off += A[i * 15 + j] * A[i * 15 + j]; }
        if (off <= 1e-60 + 1e-32 * dg) break;
        for (int p = 0; p < 15; ++p)
            for (int q = p + 1; q < 15; ++q) {
                const double apq = A[p * 15 + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * 15 + q] - A[p * 15 + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 15; ++k) { const double akp = A[k * 15 + p], akq = A[k * 15 + q]; A[k * 15 + p] = c * akp - s * akq; A[k * 15 + q] = s * akp + c * akq; }
                for (int k = 0; k < 15; ++k) { const double apk = A[p * 15 + k], aqk = A[q * 15 + k]; A[p * 15 + k] = c * apk - s * aqk; A[q * 15 + k] = s * apk + c * aqk; }
                for (int k = 0; k < 15; ++k) { const double vkp = V[k * 15 + p], vkq = V[k * 15 + q]; V[k * 15 + p] = c * vkp - s * vkq; V[k * 15 + q] = s * vkp + c * vkq; }
            }
    }
    int order[15];
    for (int i = 0; i < 15; ++i) order[i] = i;
    for (int i = 0; i < 15; ++i)
        for (int j = i + 1; j < 15; ++j)
            if (A[order[j] * 15 + order[j]] < A[order[i] * 15 + order[i]]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
    const double eps = 1e-8;
    const double* dgv = marg_g + (size_t)w * 15;
    for (int r = 0; r < 15; ++r) {
        const int o = order[r];
        const double ev = A[o * 15 + o];
        int big = 0;
        for (int k = 1; k < 15; ++k) if (fabs(V[k * 15 + o]) > fabs(V[big * 15 + o])) big = k;
        const double sgn = V[big * 15 + o] < 0 ? -1.0 : 1.0;
        const double ss = ev > eps ? sqrt(ev) : 0.0, sis = ev > eps ? sqrt(1.0 / ev) : 0.0;
        double acc = 0.0;
        for (int c = 0; c < 15; ++c) { J_lin[(size_t)w * 225 + r * 15 + c] = ss * sgn * V[c * 15 + o]; acc += sgn * V[c * 15 + o] * dgv[c]; }
        r_lin[(size_t)w * 15 + r] = -(sis * acc);
    }
    for (int c = 0; c < 15; ++c) X0[(size_t)w * 15 + c] = x[((size_t)w * n_frames + n_frames - 1) * 15 + c];
}

// ------------------------------------------------------------------ single-factor hooks
__global__ void eval_imu_factor_kernel(Consts C, const double* blob, const double* si, const double* sj, double* res, double* jac) {
    const int lane = threadIdx.x;
    __shared__ double raw[15][32];
    double col[15];
    if (lane < 30) imu_jacobian_column(C, blob, si, sj, lane, col);
    else if (lane == 30) imu_raw_residual(C, blob, si, sj, col);
    if (lane < 31) for (int r = 0; r < 15; ++r) raw[r][lane] = col[r];
    __syncwarp();
    const double* Sq = blob + 240;
    if (lane < 31)
        for (int r = 0; r < 15; ++r) {
            double s = 0.0;
            for (int k = r; k < 15; ++k) s += Sq[r * 15 + k] * raw[k][lane];
            if (lane == 30) res[r] = s; else jac[r * 30 + lane] = s;
        }
}
__global__ void eval_wheel_factor_kernel(Consts C, const double* blob, const double* pi, const double* pj, double* res, double* jac) {
    const int lane = threadIdx.x;
    if (lane > 12) return;
    V3<Dual> q[4] = {lift<Dual>(load3(pi)), lift<Dual>(load3(pi + 3)), lift<Dual>(load3(pj)), lift<Dual>(load3(pj + 3))};
    if (lane < 12) { V3<Dual>& t = q[lane / 3]; const int k = lane % 3; (k == 0 ? t.x : (k == 1 ? t.y : t.z)).d = 1.0; }
    Dual r[3];
    wheel_residuals<Dual>(C, blob, q[0], q[1], q[2], q[3], r);
    for (int k = 0; k < 3; ++k) { if (lane == 12) res[k] = r[k].a; else jac[k * 12 + lane] = r[k].d; }
}
__global__ void eval_ground_factors_kernel(Consts C, const double* pose, double* res, double* jac) {
    const int lane = threadIdx.x;
    if (lane > 6) return;
    V3<Dual> p = lift<Dual>(load3(pose)), th = lift<Dual>(load3(pose + 3));
    if (lane < 6) { V3<Dual>& t = lane < 3 ? p : th; const int k = lane % 3; (k == 0 ? t.x : (k == 1 ? t.y : t.z)).d = 1.0; }
    Dual dp, dq;
    ground_residuals<Dual>(C, p, th, &dp, &dq);
    if (lane == 6) { res[0] = dp.a; res[1] = dq.a; } else { jac[lane] = dp.d; jac[6 + lane] = dq.d; }
}
// laser_factor on one matched pair (laser_factor.h:31-89) through the scan-match algebra: res[2], jac[2][12]
__global__ void eval_laser_factor_kernel(Consts C, const double* l1_p1, const double* l1_p2, const double* l2_p1, const double* l2_p2,
                                         const double* pose_i, const double* pose_j, double* res, double* jac) {
    if (threadIdx.x != 0) return;
    double Ti[kFrameTab], Tj[kFrameTab];
    laser_frame_table(C, pose_i, Ti);
    laser_frame_table(C, pose_j, Tj);
    const double len1 = sqrt((l1_p1[0] - l1_p2[0]) * (l1_p1[0] - l1_p2[0]) + (l1_p1[1] - l1_p2[1]) * (l1_p1[1] - l1_p2[1]) + (l1_p1[2] - l1_p2[2]) * (l1_p1[2] - l1_p2[2]));
    const double len2 = sqrt((l2_p1[0] - l2_p2[0]) * (l2_p1[0] - l2_p2[0]) + (l2_p1[1] - l2_p2[1]) * (l2_p1[1] - l2_p2[1]) + (l2_p1[2] - l2_p2[2]) * (l2_p1[2] - l2_p2[2]));
    const double wgt = sqrt(fmin(len1, len2) / 2.0 / 0.02) * C.laser_sqrt_info;
    const double A1x = Ti[0] * l1_p1[0] + Ti[1] * l1_p1[1] + Ti[4], A1y = Ti[2] * l1_p1[0] + Ti[3] * l1_p1[1] + Ti[5];
    const double A2x = Ti[0] * l1_p2[0] + Ti[1] * l1_p2[1] + Ti[4], A2y = Ti[2] * l1_p2[0] + Ti[3] * l1_p2[1] + Ti[5];
    const double dx = A2x - A1x, dy = A2y - A1y, len = sqrt(dx * dx + dy * dy), inv = 1.0 / len;
    const double ux = dx * inv, uy = dy * inv, nx = -uy, ny = ux;
    const double ex = l1_p2[0] - l1_p1[0], ey = l1_p2[1] - l1_p1[1];
    for (int m = 0; m < 2; ++m) {
        const double* c = m == 0 ? l2_p1 : l2_p2;
        const double Cx = Tj[0] * c[0] + Tj[1] * c[1] + Tj[4], Cy = Tj[2] * c[0] + Tj[3] * c[1] + Tj[5];
        const double d = nx * (Cx - A2x) + ny * (Cy - A2y);
        const double tt = ux * (Cx - A2x) + uy * (Cy - A2y);
        const double s = d < 0 ? -wgt : wgt;
        res[m] = s * d;
        double* Jr = jac + m * 12;
        Jr[0] = -s * nx; Jr[1] = -s * ny; Jr[2] = 0.0;
        Jr[6] = s * nx; Jr[7] = s * ny; Jr[8] = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double* Bj = Tj + 6 + 6 * k;
            const double* Bi = Ti + 6 + 6 * k;
            Jr[9 + k] = s * (nx * (Bj[0] * c[0] + Bj[1] * c[1] + Bj[4]) + ny * (Bj[2] * c[0] + Bj[3] * c[1] + Bj[5]));
            const double alpha = (nx * (Bi[0] * ex + Bi[1] * ey) + ny * (Bi[2] * ex + Bi[3] * ey)) * inv;
            const double beta = nx * (Bi[0] * l1_p2[0] + Bi[1] * l1_p2[1] + Bi[4]) + ny * (Bi[2] * l1_p2[0] + Bi[3] * l1_p2[1] + Bi[5]);
            Jr[3 + k] = s * (-tt * alpha - beta);
        }
    }
}

// sum the scan-match tiles of every frame into one block per frame (the buffer that is all-reduced over ranks)
__global__ void reduce_tiles_kernel(const double* partial, const uint8_t* frame_active, double* out, int n_frames_total, int tiles, int pad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_frames_total * pad) return;
    const int f = idx / pad, k = idx - f * pad;
    double s = 0.0;
    if (frame_active[f])
        for (int t = 0; t < tiles; ++t) s += partial[((size_t)f * tiles + t) * pad + k];
    out[idx] = s;
}

// lvio2d_set_windows_wire: beam k of frame f -> point slot f * n_beams + k, exactly as convert::laser_to_point_times
// builds it (reference src/utilies/common.cpp:6-24: float32 angle_min + k * angle_increment with two roundings, the
// cosine / sine in double, times the float32 range) minus its 1 cm thinning, which has no meaning when every beam
// carries its own line index: a beam the reference would reject (NaN, inf, <= 0.1 m) or that has no line gets -1.
template <class IDX>   // uint16_t (0xFFFF = none) or uint8_t (0xFF = none)
__global__ void expand_wire_kernel(const float* __restrict__ ranges, const float* __restrict__ angle, const IDX* __restrict__ beam_line,
                                   int n_beams, int64_t total, double2* __restrict__ points, int32_t* __restrict__ point_line) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t f = i / n_beams;
    const int k = (int)(i - f * n_beams);
    const float r = ranges[i];
    const float ang = __fadd_rn(angle[2 * f], __fmul_rn((float)k, angle[2 * f + 1]));
    double sn, cs;
    sincos((double)ang, &sn, &cs);
    const bool valid = !isnan(r) && !isinf(r) && (double)r > 0.1;
    points[i] = valid ? make_double2(cs * (double)r, sn * (double)r) : make_double2(0.0, 0.0);
    const IDX l = beam_line[i];
    point_line[i] = (valid && l != (IDX)~(IDX)0) ? (int32_t)l : -1;
}

// lvio2d_scan_wire::shared_lines: one line list per window -> the per-frame layout the kernels read
__global__ void expand_shared_lines_kernel(const double4* __restrict__ shared, const int64_t* __restrict__ window_offset,
                                           const int64_t* __restrict__ line_offset, int n_frames, int n_total_frames, double4* __restrict__ out) {
    const int f = blockIdx.x;
    if (f >= n_total_frames) return;
    const double4* src = shared + window_offset[f / n_frames];
    double4* dst = out + line_offset[f];
    const int nl = (int)(line_offset[f + 1] - line_offset[f]);
    for (int l = threadIdx.x; l < nl; l += blockDim.x) dst[l] = src[l];
}

}  // namespace lv
