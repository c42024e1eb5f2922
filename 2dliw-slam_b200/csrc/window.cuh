// window.cuh — one warp per sliding window: everything of an LM iteration that is NOT the scan-point pass.
//
// Replaces ceres::Solve as configured by solver::solve / solver::do_init_solve (reference
// src/factor/solver.cpp:795-802, :161-168) — Ceres 1.14 TrustRegionMinimizer + LevenbergMarquardtStrategy
// with Jacobi scaling and an exact Schur/Cholesky step — together with the residual blocks Ceres would
// evaluate with Jets: imu_factor (imu_factor.h:13-89), wheel_odom_factor (wheel_factor.h:12-73),
// ground_factor_p/q (ground_factor.h:27-82; added `multiplicity` times, solver.cpp:727-743),
// marginalization_factor (marginalization_factor.h:22-53) and the constness rules of solver.cpp:787-794.
//
// One call of window_step_kernel = one trip through the minimiser loop for every window of the batch:
//   (1) sum the scan-match partials of the candidate point (scan_match.cuh), add the small factors' cost;
//   (2) Ceres' parameter-/function-tolerance tests and the step-quality test rho > 1e-3; accept or reject,
//       update the trust-region radius;
//   (3) linearise IMU / wheel / ground / prior at the accepted point, assemble the block-tridiagonal
//       (+ frame-0 arrow in the initialisation topology) normal equations with the laser blocks;
//   (4) Jacobi scaling, LM damping, block Cholesky in reverse frame order (no fill-in for either topology),
//       step, model cost change; invalid steps shrink the radius and retry in place;
//   (5) candidate = Plus(x, step) (so3_parameterization: wrap(theta + delta), factor_common.h:40-53) and
//       its laser frame tables for the next scan-match pass.
// The candidate's linearisation is evaluated in the same scan pass as its cost, so an accepted step costs
// one pass over the points, not two.
//
// Layout of every 15x15 block: packed row-major, 225 doubles (conflict-free for lane-per-row access).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lv_math.cuh"
#include "scan_match.cuh"

namespace lv {

struct LMState {
    double radius, decrease_factor, cost, model_cost_change, x_norm, initial_cost;
    int32_t iteration, num_invalid, status, termination, n_success, n_unsuccess, last_success, started;
};

struct LMOptions {
    int32_t max_iters;
    double function_tolerance, gradient_tolerance, parameter_tolerance, initial_radius;
    double max_radius, min_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
    int32_t max_consecutive_invalid;
};

struct WindowArgs {
    Consts C;
    LMOptions opt;
    int32_t n_windows, n_frames, tiles, arrow, mode;  // mode 0: solver program, 1: marginalisation program
    int32_t ground_multiplicity, prior_frame, has_imu, has_wheel;
    const uint8_t* const_mask;      // [B*n]
    const uint8_t* frame_active;    // [B*n] laser block active
    const int32_t* ref_frame;       // [B*n] or nullptr
    const double* imu;              // [B*(n-1)][466]
    const double* wheel;            // [B*(n-1)][15]
    const double* prior_X0;         // [B][15]
    const double* prior_J;          // [B][225]
    const double* partial;          // [B*n][tiles][pad]   scan-match output at the candidate
    double* x;                      // [B*n][15] accepted point
    double* xc;                     // [B*n][15] candidate
    double* scale;                  // [B*n][15] Jacobi scaling
    double* laser_blocks;           // [B*n][pad] blocks at the accepted point
    double* frame_tab;              // [B*n][24] tables of the candidate
    double* pair;                   // [B*(n-1)][3][225]  H_aa | H_ab | H_bb of pair (i-1, i)
    double* fac;                    // [B*n][3][225]      L_i | E_i | F_i
    LMState* state;                 // [B]
    int32_t* win_status;            // [B] mirror of state.status for the scan-match kernel
    // optional dense outputs (lvio2d_linearize): H [B][15n][15n], g [B][15n], cost [B]
    double* dense_H;
    double* dense_g;
    double* dense_cost;
    // optional marginalisation outputs
    double* marg_H;                 // [B][225] Schur complement on the last frame
    double* marg_g;                 // [B][15]
};

constexpr int kBlk = 225;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// ---- warp-cooperative 15x15 kernels on shared memory (lane r owns row r; lanes >= 15 idle)
// in-place lower Cholesky; returns false (uniformly) when a pivot is not positive
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
// X <- X L^-T : row r of X solved against lower-triangular L (both in shared memory)
__device__ __forceinline__ void trsm_rlt15(double* X, const double* L, int lane) {
    if (lane < 15) {
        double row[15];
#pragma unroll
        for (int k = 0; k < 15; ++k) row[k] = X[lane * 15 + k];
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            double s = row[k];
#pragma unroll
            for (int m = 0; m < k; ++m) s -= row[m] * L[k * 15 + m];
            row[k] = s / L[k * 15 + k];
        }
#pragma unroll
        for (int k = 0; k < 15; ++k) X[lane * 15 + k] = row[k];
    }
    __syncwarp();
}
// C -= A B^T   (all 15x15, row r of C by lane r)
__device__ __forceinline__ void gemm_sub_abt15(double* Cm, const double* A, const double* B, int lane) {
    if (lane < 15) {
        double a[15];
#pragma unroll
        for (int k = 0; k < 15; ++k) a[k] = A[lane * 15 + k];
        for (int c = 0; c < 15; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 15; ++k) s += a[k] * B[c * 15 + k];
            Cm[lane * 15 + c] -= s;
        }
    }
    __syncwarp();
}
// y = L^-1 b (forward), in place on a 15-vector in shared memory; executed redundantly by lane 0
__device__ __forceinline__ void fwd15(const double* L, double* b, int lane) {
    if (lane == 0) {
        for (int k = 0; k < 15; ++k) {
            double s = b[k];
            for (int m = 0; m < k; ++m) s -= L[k * 15 + m] * b[m];
            b[k] = s / L[k * 15 + k];
        }
    }
    __syncwarp();
}
// y = L^-T b (backward)
__device__ __forceinline__ void bwd15(const double* L, double* b, int lane) {
    if (lane == 0) {
        for (int k = 14; k >= 0; --k) {
            double s = b[k];
            for (int m = k + 1; m < 15; ++m) s -= L[m * 15 + k] * b[m];
            b[k] = s / L[k * 15 + k];
        }
    }
    __syncwarp();
}
// out[r] -= sum_k M[r][k] v[k]      (TRANS: M[k][r])
template <bool TRANS>
__device__ __forceinline__ void gemv_sub15(double* out, const double* M, const double* v, int lane) {
    if (lane < 15) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 15; ++k) s += (TRANS ? M[k * 15 + lane] : M[lane * 15 + k]) * v[k];
        out[lane] -= s;
    }
    __syncwarp();
}
__device__ __forceinline__ void copy_blk(double* dst, const double* src, int lane) {
    for (int i = lane; i < kBlk; i += 32) dst[i] = src[i];
}

// per-warp shared memory carve-up (doubles)
struct WarpSmem {
    double* cb;      // [n][pad]   laser blocks of the candidate / accepted point
    double* g;       // [n][15]    gradient J^T r
    double* b;       // [n][15]    rhs / solution
    double* hdiag;   // [n][15]    diag(J^T J)
    double* sc;      // [n][15]    Jacobi scaling
    double* own;     // [n][21]    6x6 pose block (upper) from ground (+ laser added at assembly)
    double* blk;     // 6 x 225 work blocks; aliased by the pair linearisation scratch
    __host__ __device__ static size_t doubles(int n, int pad) { return (size_t)n * (pad + 15 * 4 + 21) + 7 * kBlk + 16; }
};

__device__ __forceinline__ int pose_index(int k) { return k < 2 ? k : k + 1; }  // 5-vector (px py th0 th1 th2) -> 6-dim pose

// ---------------------------------------------------------------------------------------------------
template <bool ARROW>
__device__ void window_step(const WindowArgs& a, int w, int lane, double* ws) {
    constexpr int NPAD = ARROW ? kPadFree : kPadTrack;
    constexpr int ICOST = ARROW ? 44 : 20;
    const int n = a.n_frames;
    LMState st = a.state[w];
    if (st.status != 0) return;
    const LMOptions& opt = a.opt;

    WarpSmem S;
    S.cb = ws;
    S.g = S.cb + (size_t)n * NPAD;
    S.b = S.g + n * 15;
    S.hdiag = S.b + n * 15;
    S.sc = S.hdiag + n * 15;
    S.own = S.sc + n * 15;
    S.blk = S.own + n * 21;
    if ((reinterpret_cast<uintptr_t>(S.blk) & 15) != 0) S.blk += 1;

    double* x = a.x + (size_t)w * n * 15;
    double* xc = a.xc + (size_t)w * n * 15;
    const uint8_t* cm = a.const_mask + (size_t)w * n;
    const uint8_t* fa = a.frame_active + (size_t)w * n;
    const int mode = a.mode;
    auto is_const = [&](int f, int c) -> bool {  // c: column 0..14 of frame f
        if (mode == 1) return false;
        const int blk = c < 3 ? 0 : (c < 6 ? 1 : (c < 9 ? 2 : 3));
        return (cm[f] >> blk) & 1;
    };
    const double lsq = a.C.laser_sqrt_info * a.C.laser_sqrt_info;

    // ---- (1) candidate laser blocks: sum the tiles in a fixed order
    for (int idx = lane; idx < n * NPAD; idx += 32) {
        const int f = idx / NPAD, k = idx - f * NPAD;
        double s = 0.0;
        if (fa[f]) {
            const double* p = a.partial + ((size_t)(w * n + f) * a.tiles) * NPAD + k;
            for (int t = 0; t < a.tiles; ++t) s += p[(size_t)t * NPAD];
        }
        S.cb[idx] = s;
    }
    __syncwarp();

    // ---- small-factor cost at a point (value only): lanes over pairs / frames
    auto small_cost = [&](const double* X) -> double {
        double c = 0.0;
        for (int i = 1 + lane; i < n; i += 32) {
            const uint8_t ma = mode == 1 ? 0 : cm[i - 1], mb = mode == 1 ? 0 : cm[i];
            if (a.has_imu && ((ma & 15) != 15 || (mb & 15) != 15)) {
                const double* blob = a.imu + ((size_t)w * (n - 1) + (i - 1)) * 466;
                double r[15];
                imu_raw_residual(a.C, blob, X + 15 * (i - 1), X + 15 * i, r);
                const double* Sq = blob + 240;
                for (int row = 0; row < 15; ++row) {
                    double s = 0.0;
                    for (int k = row; k < 15; ++k) s += Sq[row * 15 + k] * r[k];  // sqrt_inverse_P = L^T is upper triangular
                    c += s * s;
                }
            }
            if (a.has_wheel && ((ma & 3) != 3 || (mb & 3) != 3)) {
                const double* blob = a.wheel + ((size_t)w * (n - 1) + (i - 1)) * 15;
                double r[3];
                wheel_residuals<double>(a.C, blob, load3(X + 15 * (i - 1)), load3(X + 15 * (i - 1) + 3), load3(X + 15 * i),
                                        load3(X + 15 * i + 3), r);
                c += r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            }
        }
        if (a.ground_multiplicity > 0)
            for (int f = lane; f < n; f += 32) {
                if (mode != 1 && (cm[f] & 3) == 3) continue;
                double rp, rq;
                ground_residuals<double>(a.C, load3(X + 15 * f), load3(X + 15 * f + 3), &rp, &rq);
                c += a.ground_multiplicity * (rp * rp + rq * rq);
            }
        if (a.prior_frame >= 0 && lane < 15 && (mode == 1 || (cm[a.prior_frame] & 15) != 15)) {
            const double* J = a.prior_J + (size_t)w * kBlk;
            const double* X0 = a.prior_X0 + (size_t)w * 15;
            const double* xp = X + 15 * a.prior_frame;
            double s = 0.0;
            for (int k = 0; k < 15; ++k) s += J[lane * 15 + k] * (xp[k] - X0[k]);
            c += s * s;
        }
        return warp_sum(c);
    };
    auto laser_cost = [&]() -> double {
        double c = 0.0;
        for (int f = lane; f < n; f += 32) c += S.cb[f * NPAD + ICOST];
        return lsq * warp_sum(c);
    };

    // ---- (2) decide on the candidate
    bool accepted = false;
    if (!st.started) {
        // iteration 0: the "candidate" is the initial point
        st.started = 1;
        st.cost = 0.5 * (laser_cost() + small_cost(xc));
        st.initial_cost = st.cost;
        accepted = true;
        double s = 0.0;
        for (int i = lane; i < n * 15; i += 32) {
            const double v = xc[i];
            x[i] = v;
            if (!is_const(i / 15, i % 15)) s += v * v;
        }
        st.x_norm = sqrt(warp_sum(s));
        st.last_success = 1;
    } else {
        double cand_cost = 0.5 * (laser_cost() + small_cost(xc));
        if (!isfinite(cand_cost)) cand_cost = 1.7976931348623157e308;
        double s = 0.0;
        for (int i = lane; i < n * 15; i += 32)
            if (!is_const(i / 15, i % 15)) { const double d = x[i] - xc[i]; s += d * d; }
        const double step_norm = sqrt(warp_sum(s));
        // ParameterToleranceReached / FunctionToleranceReached come before the step-quality test
        if (step_norm <= opt.parameter_tolerance * (st.x_norm + opt.parameter_tolerance)) {
            st.status = 1; st.termination = 2;
        } else {
            const double cost_change = st.cost - cand_cost;
            if (fabs(cost_change) <= opt.function_tolerance * st.cost) {
                st.status = 1; st.termination = 1;
            } else {
                const double rho = cost_change / st.model_cost_change;
                accepted = rho > opt.min_relative_decrease;
                if (accepted) {
                    double s2 = 0.0;
                    for (int i = lane; i < n * 15; i += 32) {
                        const double v = xc[i];
                        x[i] = v;
                        if (!is_const(i / 15, i % 15)) s2 += v * v;
                    }
                    st.x_norm = sqrt(warp_sum(s2));
                    st.cost = cand_cost;
                    const double t = 2.0 * rho - 1.0;
                    st.radius = st.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
                    st.radius = fmin(opt.max_radius, st.radius);
                    st.decrease_factor = 2.0;
                    st.last_success = 1;
                    ++st.n_success;
                } else {
                    st.radius = st.radius / st.decrease_factor;
                    st.decrease_factor *= 2.0;
                    st.last_success = 0;
                    ++st.n_unsuccess;
                }
            }
        }
        if (st.status != 0) {
            if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
            return;
        }
    }
    __syncwarp();
    double* lbw = a.laser_blocks + (size_t)w * n * NPAD;
    if (accepted) {
        for (int idx = lane; idx < n * NPAD; idx += 32) lbw[idx] = S.cb[idx];
    } else {
        for (int idx = lane; idx < n * NPAD; idx += 32) S.cb[idx] = lbw[idx];
    }
    __syncwarp();
    if (mode == 0 && st.iteration >= opt.max_iters) {
        st.status = 1; st.termination = 0;
        if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
        return;
    }

    // ---- (3) linearise the small factors at x
    for (int i = lane; i < n * 15; i += 32) { S.g[i] = 0.0; S.hdiag[i] = 0.0; }
    for (int i = lane; i < n * 21; i += 32) S.own[i] = 0.0;
    __syncwarp();
    double* pairw = a.pair + (size_t)w * (n - 1) * 3 * kBlk;
    {
        // scratch aliases the work blocks: blob 466 -> 480 | Jw [15][32] | col-major H rows handled in registers
        double* sblob = S.blk;             // 480
        double* sJ = S.blk + 480;          // 15 x 32 : whitened Jacobian, column 30 = whitened residual
        double* sW = sJ + 480;             // wheel: 3 x 16 (cols 0..11 jac, col 12 residual)
        for (int i = 1; i < n; ++i) {
            const uint8_t ma = mode == 1 ? 0 : cm[i - 1], mb = mode == 1 ? 0 : cm[i];
            const double* xa = x + 15 * (i - 1);
            const double* xb = x + 15 * i;
            const bool imu_on = a.has_imu && ((ma & 15) != 15 || (mb & 15) != 15);
            const bool wheel_on = a.has_wheel && ((ma & 3) != 3 || (mb & 3) != 3);
            double col[15];
            if (imu_on) {
                const double* blob = a.imu + ((size_t)w * (n - 1) + (i - 1)) * 466;
                for (int k = lane; k < 466; k += 32) sblob[k] = blob[k];
                __syncwarp();
                if (lane < 30) imu_jacobian_column(a.C, sblob, xa, xb, lane, col);
                else if (lane == 30) imu_raw_residual(a.C, sblob, xa, xb, col);
                if (lane < 31) {
                    // whiten with the upper-triangular sqrt_inverse_P; constant parameter blocks get zero columns
                    const bool dead = lane < 30 && is_const(lane < 15 ? i - 1 : i, lane % 15);
                    const double* Sq = sblob + 240;
#pragma unroll
                    for (int r = 0; r < 15; ++r) {
                        double s = 0.0;
#pragma unroll
                        for (int k = r; k < 15; ++k) s += Sq[r * 15 + k] * col[k];
                        sJ[r * 32 + lane] = dead ? 0.0 : s;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 15; ++r) sJ[r * 32 + 31] = 0.0;
                }
            } else {
                for (int k = lane; k < 480; k += 32) sJ[k] = 0.0;
            }
            if (wheel_on) {
                const double* blob = a.wheel + ((size_t)w * (n - 1) + (i - 1)) * 15;
                if (lane < 13) {
                    V3<Dual> q[4] = {lift<Dual>(load3(xa)), lift<Dual>(load3(xa + 3)), lift<Dual>(load3(xb)), lift<Dual>(load3(xb + 3))};
                    if (lane < 12) {
                        V3<Dual>& t = q[lane / 3];
                        const int k = lane % 3;
                        (k == 0 ? t.x : (k == 1 ? t.y : t.z)).d = 1.0;
                    }
                    Dual r[3];
                    wheel_residuals<Dual>(a.C, blob, q[0], q[1], q[2], q[3], r);
                    const bool dead = lane < 12 && is_const(lane < 6 ? i - 1 : i, lane % 6);
#pragma unroll
                    for (int k = 0; k < 3; ++k) sW[k * 16 + lane] = lane == 12 ? r[k].a : (dead ? 0.0 : r[k].d);
                }
            } else {
                for (int k = lane; k < 48; k += 32) sW[k] = 0.0;
            }
            __syncwarp();
            // H_pair row `lane` (30 entries) + gradient entry, straight from shared memory
            if (lane < 30) {
                double hrow[30];
                double mine[15];
#pragma unroll
                for (int r = 0; r < 15; ++r) mine[r] = sJ[r * 32 + lane];
                double gsum = 0.0;
#pragma unroll
                for (int r = 0; r < 15; ++r) gsum += mine[r] * sJ[r * 32 + 30];
#pragma unroll
                for (int c = 0; c < 30; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < 15; ++r) s += mine[r] * sJ[r * 32 + c];
                    hrow[c] = s;
                }
                // wheel: pose columns only.  lane -> wheel column: frame a (0..5) | frame b (15..20)
                const int fl = lane % 15;
                if (fl < 6) {
                    const int wc = (lane < 15 ? 0 : 6) + fl;
                    const double w0 = sW[wc], w1 = sW[16 + wc], w2 = sW[32 + wc];
                    gsum += w0 * sW[12] + w1 * sW[16 + 12] + w2 * sW[32 + 12];
#pragma unroll
                    for (int c = 0; c < 12; ++c) {
                        const int hc = (c < 6 ? 0 : 15) + (c % 6);
                        hrow[hc] += w0 * sW[c] + w1 * sW[16 + c] + w2 * sW[32 + c];
                    }
                }
                const int fr = lane < 15 ? i - 1 : i;
                S.g[fr * 15 + fl] += gsum;
                S.hdiag[fr * 15 + fl] += hrow[lane];
                double* P = pairw + (size_t)(i - 1) * 3 * kBlk;
                if (lane < 15) {
#pragma unroll
                    for (int c = 0; c < 15; ++c) { P[fl * 15 + c] = hrow[c]; P[kBlk + fl * 15 + c] = hrow[15 + c]; }
                } else {
#pragma unroll
                    for (int c = 0; c < 15; ++c) P[2 * kBlk + fl * 15 + c] = hrow[15 + c];
                }
            }
            __syncwarp();
        }
    }
    // ground: lanes over frames, 6 dual directions each -> 6x6 upper block + gradient
    if (a.ground_multiplicity > 0)
        for (int f = lane; f < n; f += 32) {
            if (mode != 1 && (cm[f] & 3) == 3) continue;
            double rp, rq, jp[6], jq[6];
            ground_residuals<double>(a.C, load3(x + 15 * f), load3(x + 15 * f + 3), &rp, &rq);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                V3<Dual> p = lift<Dual>(load3(x + 15 * f)), th = lift<Dual>(load3(x + 15 * f + 3));
                V3<Dual>& t = c < 3 ? p : th;
                (c % 3 == 0 ? t.x : (c % 3 == 1 ? t.y : t.z)).d = 1.0;
                Dual dp, dq;
                ground_residuals<Dual>(a.C, p, th, &dp, &dq);
                const bool dead = is_const(f, c);
                jp[c] = dead ? 0.0 : dp.d;
                jq[c] = dead ? 0.0 : dq.d;
            }
            const double m = (double)a.ground_multiplicity;
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                S.g[f * 15 + r] += m * (jp[r] * rp + jq[r] * rq);
#pragma unroll
                for (int c = r; c < 6; ++c) S.own[f * 21 + k++] += m * (jp[r] * jp[c] + jq[r] * jq[c]);
            }
        }
    __syncwarp();
    // laser blocks -> gradient and diag; arrow parts are added at assembly time
    for (int f = lane; f < n; f += 32) {
        if (!fa[f]) continue;
        const double* cbf = S.cb + f * NPAD;
        const int gj = ARROW ? 36 : 15;
        const bool own_free = !(is_const(f, 0) && is_const(f, 3));
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int pc = pose_index(k);
            if (own_free && !is_const(f, pc)) S.g[f * 15 + pc] += lsq * cbf[gj + k];
        }
        int k2 = 0;
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int c = r; c < 5; ++c) {
                // 15 upper entries of the 5x5 block in the order aa(3) | a x bj(6) | bj bj(6)
                const int src = (r < 2 && c < 2) ? (r + c) : (r < 2 ? 3 + r * 3 + (c - 2) : 9 + (r == 2 ? c - 2 : (r == 3 ? 2 + c - 2 : 5)));
                const int pr = pose_index(r), pc = pose_index(c);
                // position of (pr, pc) in the 21-entry upper triangle of the 6x6 block
                const int dst = pr * 6 - pr * (pr - 1) / 2 + (pc - pr);
                if (!is_const(f, pr) && !is_const(f, pc)) S.own[f * 21 + dst] += lsq * cbf[src];
                ++k2;
            }
    }
    __syncwarp();
    if (ARROW && mode == 0) {
        // reference-frame side of the laser blocks: gradient of frame rf and its 6x6 block (summed over all j)
        for (int f = 0; f < n; ++f) {
            if (!fa[f]) continue;
            const int rf = a.ref_frame ? a.ref_frame[(size_t)w * n + f] : -1;
            if (rf < 0) continue;
            const double* cbf = S.cb + f * NPAD;
            if (lane < 5) {
                const int pc = pose_index(lane);
                if (!is_const(rf, pc)) {
                    const double gi = lane < 2 ? -cbf[36 + lane] : cbf[41 + lane - 2];
                    S.g[rf * 15 + pc] += lsq * gi;
                }
            }
            if (lane < 15) {
                // upper 5x5 of H_ii: aa | -(a x bi) | bi bi
                int r = 0, c = lane;
                while (c >= 5 - r) { c -= 5 - r; ++r; }
                c += r;
                double v;
                if (r < 2 && c < 2) v = cbf[r + c];
                else if (r < 2) v = -cbf[15 + r * 3 + (c - 2)];
                else v = cbf[21 + (r == 2 ? c - 2 : (r == 3 ? 2 + c - 2 : 5))];
                const int pr = pose_index(r), pc = pose_index(c);
                const int dst = pr * 6 - pr * (pr - 1) / 2 + (pc - pr);
                if (!is_const(rf, pr) && !is_const(rf, pc)) S.own[rf * 21 + dst] += lsq * v;
            }
            __syncwarp();
        }
    }
    // own-block diagonals -> hdiag
    for (int f = lane; f < n; f += 32) {
#pragma unroll
        for (int r = 0; r < 6; ++r) S.hdiag[f * 15 + r] += S.own[f * 21 + r * 6 - r * (r - 1) / 2];
    }
    __syncwarp();
    // prior: r = J (x - X0); H += J^T J; g += J^T r
    const bool prior_on = a.prior_frame >= 0 && (mode == 1 || (cm[a.prior_frame] & 15) != 15);
    double* sPrior = S.blk + 5 * kBlk;  // J^T J of the prior (kept through the solve)
    if (prior_on) {
        const double* J = a.prior_J + (size_t)w * kBlk;
        const double* X0 = a.prior_X0 + (size_t)w * 15;
        const double* xp = x + 15 * a.prior_frame;
        double* tmp = S.blk;  // r
        if (lane < 15) {
            double s = 0.0;
            for (int k = 0; k < 15; ++k) s += J[lane * 15 + k] * (xp[k] - X0[k]);
            tmp[lane] = s;
        }
        __syncwarp();
        if (lane < 15) {
            const bool dead = is_const(a.prior_frame, lane);
            double gs = 0.0;
            for (int r = 0; r < 15; ++r) gs += J[r * 15 + lane] * tmp[r];
            for (int c = 0; c < 15; ++c) {
                double s = 0.0;
                for (int r = 0; r < 15; ++r) s += J[r * 15 + lane] * J[r * 15 + c];
                sPrior[lane * 15 + c] = (dead || is_const(a.prior_frame, c)) ? 0.0 : s;
            }
            if (!dead) {
                S.g[a.prior_frame * 15 + lane] += gs;
                S.hdiag[a.prior_frame * 15 + lane] += sPrior[lane * 15 + lane];
            }
        }
        __syncwarp();
    }

    // assembled (unscaled) diagonal block of frame i into dst
    auto assemble_D = [&](int i, double* dst) {
        for (int e = lane; e < kBlk; e += 32) {
            double v = 0.0;
            if (i >= 1) v += pairw[(size_t)(i - 1) * 3 * kBlk + 2 * kBlk + e];
            if (i + 1 < n) v += pairw[(size_t)i * 3 * kBlk + e];
            const int r = e / 15, c = e - r * 15;
            if (r < 6 && c < 6) {
                const int lo = r < c ? r : c, hi = r < c ? c : r;
                v += S.own[i * 21 + lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
            }
            if (prior_on && i == a.prior_frame) v += sPrior[e];
            dst[e] = v;
        }
        __syncwarp();
    };
    // laser cross block between frame j and its reference frame rf, as H(rf, j) (rows rf, cols j), 6x6 pose part
    auto cross_entry = [&](int j, int r, int c) -> double {  // r: pose row of rf, c: pose col of j (0..5)
        if (r == 2 || c == 2) return 0.0;
        const double* cbf = S.cb + j * NPAD;
        const int ri = r < 2 ? r : r - 1, ci = c < 2 ? c : c - 1;  // 5-vector indices
        double v;
        if (ri < 2 && ci < 2) v = -cbf[ri + ci];                      // -(a a^T)
        else if (ri < 2) v = -cbf[3 + ri * 3 + (ci - 2)];             // rows p_i (-a), cols theta_j (bj)
        else if (ci < 2) v = cbf[15 + ci * 3 + (ri - 2)];             // rows theta_i (bi), cols p_j (a)
        else v = cbf[27 + (ci - 2) * 3 + (ri - 2)];                   // bj x bi stored [bj][bi]
        return lsq * v;
    };

    // dense outputs for lvio2d_linearize
    if (a.dense_H) {
        const int dim = 15 * n;
        double* H = a.dense_H + (size_t)w * dim * dim;
        for (size_t e = lane; e < (size_t)dim * dim; e += 32) H[e] = 0.0;
        __syncwarp();
        double* Dm = S.blk;
        for (int i = 0; i < n; ++i) {
            assemble_D(i, Dm);
            for (int e = lane; e < kBlk; e += 32) H[(size_t)(15 * i + e / 15) * dim + 15 * i + e % 15] = Dm[e];
            if (i >= 1)
                for (int e = lane; e < kBlk; e += 32) {
                    const double v = pairw[(size_t)(i - 1) * 3 * kBlk + kBlk + e];  // H(i-1, i)
                    H[(size_t)(15 * (i - 1) + e / 15) * dim + 15 * i + e % 15] += v;
                    H[(size_t)(15 * i + e % 15) * dim + 15 * (i - 1) + e / 15] += v;
                }
            __syncwarp();
            if (ARROW && mode == 0 && fa[i]) {
                const int rf = a.ref_frame ? a.ref_frame[(size_t)w * n + i] : -1;
                if (rf >= 0)
                    for (int e = lane; e < 36; e += 32) {
                        const int r = e / 6, c = e % 6;
                        if (is_const(rf, r) || is_const(i, c)) continue;
                        const double v = cross_entry(i, r, c);
                        H[(size_t)(15 * rf + r) * dim + 15 * i + c] += v;
                        H[(size_t)(15 * i + c) * dim + 15 * rf + r] += v;
                    }
            }
            __syncwarp();
        }
        for (int i = lane; i < dim; i += 32) a.dense_g[(size_t)w * dim + i] = S.g[i];
        if (lane == 0) a.dense_cost[w] = st.cost;
        st.status = 1;
        if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
        return;
    }

    // ---- marginalisation program: forward elimination of frames 0..n-2 (solver.cpp:4-40)
    if (mode == 1) {
        double* Dm = S.blk;            // current diagonal block
        double* Um = S.blk + kBlk;     // coupling H(i, i+1)
        double* Nx = S.blk + 2 * kBlk; // next diagonal block
        double* rhs = S.b;             // g = -J^T r
        for (int i = lane; i < n * 15; i += 32) rhs[i] = -S.g[i];
        __syncwarp();
        assemble_D(0, Dm);
        for (int i = 0; i + 1 < n; ++i) {
            assemble_D(i + 1, Nx);
            copy_blk(Um, pairw + (size_t)i * 3 * kBlk + kBlk, lane);  // H(i, i+1): rows i, cols i+1
            __syncwarp();
            // E = U^T L^-T (rows i+1) ; Nx -= E E^T ; rhs_{i+1} -= E (L^-1 rhs_i)
            if (!chol15(Dm, lane)) { st.termination = 5; break; }
            double* Et = S.blk + 3 * kBlk;
            for (int e = lane; e < kBlk; e += 32) Et[e] = Um[(e % 15) * 15 + e / 15];
            __syncwarp();
            trsm_rlt15(Et, Dm, lane);
            gemm_sub_abt15(Nx, Et, Et, lane);
            fwd15(Dm, rhs + 15 * i, lane);
            gemv_sub15<false>(rhs + 15 * (i + 1), Et, rhs + 15 * i, lane);
            copy_blk(Dm, Nx, lane);
            __syncwarp();
        }
        for (int e = lane; e < kBlk; e += 32) a.marg_H[(size_t)w * kBlk + e] = Dm[e];
        if (lane < 15) a.marg_g[(size_t)w * 15 + lane] = rhs[15 * (n - 1) + lane];
        st.status = 1;
        if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
        return;
    }

    // ---- gradient tolerance: |x - Plus(x, -g)|_inf (only after a successful step)
    if (st.last_success) {
        double mx = 0.0;
        for (int f = lane; f < n; f += 32) {
#pragma unroll
            for (int c = 0; c < 15; ++c) {
                if (is_const(f, c) || (c >= 3 && c < 6)) continue;
                mx = fmax(mx, fabs(S.g[f * 15 + c]));
            }
            if (!is_const(f, 3)) {
                double neg[3] = {-S.g[f * 15 + 3], -S.g[f * 15 + 4], -S.g[f * 15 + 5]}, out[3];
                so3_plus(x + 15 * f + 3, neg, out);
#pragma unroll
                for (int c = 0; c < 3; ++c) mx = fmax(mx, fabs(x[15 * f + 3 + c] - out[c]));
            }
        }
        mx = warp_max(mx);
        if (mx <= opt.gradient_tolerance) {
            st.status = 1; st.termination = 3;
            if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
            return;
        }
    }

    // ---- Jacobi scaling (fixed at iteration 0)
    double* scw = a.scale + (size_t)w * n * 15;
    if (st.iteration == 0) {
        for (int i = lane; i < n * 15; i += 32) {
            const double s = is_const(i / 15, i % 15) ? 1.0 : 1.0 / (1.0 + sqrt(S.hdiag[i]));
            S.sc[i] = s;
            scw[i] = s;
        }
    } else {
        for (int i = lane; i < n * 15; i += 32) S.sc[i] = scw[i];
    }
    __syncwarp();

    // ---- (4) trust-region step; invalid steps shrink the radius and retry without a new evaluation
    double* facw = a.fac + (size_t)w * n * 3 * kBlk;
    double* Dm = S.blk;               // diagonal block being factored
    double* Cy = S.blk + kBlk;        // carry: updated diagonal block of frame i-1
    double* Em = S.blk + 2 * kBlk;    // E_i
    double* Fm = S.blk + 3 * kBlk;    // F_i (arrow)
    double* Wc = S.blk + 4 * kBlk;    // carry: updated arrow block H(0, i-1)
    // S.blk + 5*kBlk holds the prior block
    bool have_step = false;
    double step_dot_g = 0.0, lm_quad = 0.0;
    for (;;) {
        if (st.radius < opt.min_radius) { st.status = 1; st.termination = 4; break; }
        ++st.iteration;
        bool ok = true;
        // rhs = scaled gradient
        for (int i = lane; i < n * 15; i += 32) S.b[i] = is_const(i / 15, i % 15) ? 0.0 : S.g[i] * S.sc[i];
        __syncwarp();
        auto scale_damp = [&](int i, double* blkp) {  // A = S H S + diag(clamp(diag(S H S)) / radius); const entries -> identity
            for (int e = lane; e < kBlk; e += 32) {
                const int r = e / 15, c = e - r * 15;
                double v = blkp[e] * S.sc[i * 15 + r] * S.sc[i * 15 + c];
                if (r == c) {
                    if (is_const(i, r)) v = 1.0;
                    else v += fmin(fmax(v, opt.min_lm_diagonal), opt.max_lm_diagonal) / st.radius;
                }
                blkp[e] = v;
            }
            __syncwarp();
        };
        double* D0acc = nullptr;
        // reverse elimination n-1 .. 1
        assemble_D(n - 1, Dm);
        scale_damp(n - 1, Dm);
        bool have_wc = false;
        for (int i = n - 1; i >= 1 && ok; --i) {
            // coupling U_i = H(i-1, i) scaled, plus the laser cross block when frame i's reference is frame i-1 == 0
            for (int e = lane; e < kBlk; e += 32) {
                const int r = e / 15, c = e - r * 15;
                Em[e] = pairw[(size_t)(i - 1) * 3 * kBlk + kBlk + e] * S.sc[(i - 1) * 15 + r] * S.sc[i * 15 + c];
            }
            __syncwarp();
            const int rf = (ARROW && fa[i] && a.ref_frame) ? a.ref_frame[(size_t)w * n + i] : -1;
            bool arrow_i = false;
            if (ARROW) {
                // arrow block H(0, i): laser cross term (if rf == 0) + fill-in carried from frame i+1
                if (i >= 2) {
                    for (int e = lane; e < kBlk; e += 32) {
                        const int r = e / 15, c = e - r * 15;
                        double v = have_wc ? Wc[e] : 0.0;
                        if (rf == 0 && r < 6 && c < 6 && !is_const(0, r) && !is_const(i, c))
                            v += cross_entry(i, r, c) * S.sc[r] * S.sc[i * 15 + c];
                        Fm[e] = v;
                    }
                    arrow_i = have_wc || rf == 0;
                } else {
                    // i == 1: the arrow block IS the tridiagonal coupling H(0,1)
                    for (int e = lane; e < kBlk; e += 32) {
                        const int r = e / 15, c = e - r * 15;
                        double v = have_wc ? Wc[e] : 0.0;
                        if (rf == 0 && r < 6 && c < 6 && !is_const(0, r) && !is_const(1, c))
                            v += cross_entry(1, r, c) * S.sc[r] * S.sc[15 + c];
                        Em[e] += v;
                    }
                }
                __syncwarp();
            }
            ok = chol15(Dm, lane);
            if (!ok) break;
            trsm_rlt15(Em, Dm, lane);
            if (ARROW && arrow_i) trsm_rlt15(Fm, Dm, lane);
            // store the factor blocks for the back substitution
            for (int e = lane; e < kBlk; e += 32) {
                facw[(size_t)i * 3 * kBlk + e] = Dm[e];
                facw[(size_t)i * 3 * kBlk + kBlk + e] = Em[e];
                if (ARROW) facw[(size_t)i * 3 * kBlk + 2 * kBlk + e] = arrow_i ? Fm[e] : 0.0;
            }
            // rhs: z_i = L_i^-1 b_i ; b_{i-1} -= E_i z_i ; b_0 -= F_i z_i
            fwd15(Dm, S.b + 15 * i, lane);
            gemv_sub15<false>(S.b + 15 * (i - 1), Em, S.b + 15 * i, lane);
            if (ARROW && arrow_i) gemv_sub15<false>(S.b, Fm, S.b + 15 * i, lane);
            // next diagonal block (frame i-1) minus E E^T
            assemble_D(i - 1, Cy);
            scale_damp(i - 1, Cy);
            gemm_sub_abt15(Cy, Em, Em, lane);
            if (ARROW && arrow_i) {
                // D_0 -= F F^T (kept in the dedicated accumulator until frame 0 is reached) ; H(0, i-1) -= F E^T
                if (!D0acc) {
                    D0acc = S.blk + 5 * kBlk + kBlk;  // after the prior block
                    for (int e = lane; e < kBlk; e += 32) D0acc[e] = 0.0;
                    __syncwarp();
                }
                gemm_sub_abt15(D0acc, Fm, Fm, lane);
                for (int e = lane; e < kBlk; e += 32) Wc[e] = 0.0;
                __syncwarp();
                gemm_sub_abt15(Wc, Fm, Em, lane);
                have_wc = true;
            } else if (ARROW) {
                have_wc = false;
            }
            if (ARROW && i - 1 == 0 && D0acc) {
                for (int e = lane; e < kBlk; e += 32) Cy[e] += D0acc[e];
                __syncwarp();
            }
            copy_blk(Dm, Cy, lane);
            __syncwarp();
        }
        if (ok) ok = chol15(Dm, lane);
        if (ok) {
            for (int e = lane; e < kBlk; e += 32) facw[e] = Dm[e];
            fwd15(Dm, S.b, lane);
            bwd15(Dm, S.b, lane);  // y_0
            for (int i = 1; i < n; ++i) {
                copy_blk(Dm, facw + (size_t)i * 3 * kBlk, lane);
                copy_blk(Em, facw + (size_t)i * 3 * kBlk + kBlk, lane);
                if (ARROW) copy_blk(Fm, facw + (size_t)i * 3 * kBlk + 2 * kBlk, lane);
                __syncwarp();
                gemv_sub15<true>(S.b + 15 * i, Em, S.b + 15 * (i - 1), lane);
                if (ARROW && i >= 2) gemv_sub15<true>(S.b + 15 * i, Fm, S.b, lane);
                bwd15(Dm, S.b + 15 * i, lane);
            }
            // step = -y ; model_cost_change = -1/2 step.gs + 1/2 sum lm_diag step^2
            double sg = 0.0, lq = 0.0;
            bool finite = true;
            for (int i = lane; i < n * 15; i += 32) {
                if (is_const(i / 15, i % 15)) { S.b[i] = 0.0; continue; }
                const double stp = -S.b[i];
                S.b[i] = stp;
                if (!isfinite(stp)) finite = false;
                const double gs = S.g[i] * S.sc[i];
                const double hs = S.hdiag[i] * S.sc[i] * S.sc[i];
                sg += stp * gs;
                lq += fmin(fmax(hs, opt.min_lm_diagonal), opt.max_lm_diagonal) / st.radius * stp * stp;
            }
            sg = warp_sum(sg);
            lq = warp_sum(lq);
            finite = __all_sync(0xffffffffu, finite);
            st.model_cost_change = -0.5 * sg + 0.5 * lq;
            ok = finite && st.model_cost_change > 0.0;
            step_dot_g = sg; lm_quad = lq;
        }
        __syncwarp();
        if (ok) { have_step = true; st.num_invalid = 0; break; }
        // HandleInvalidStep
        ++st.num_invalid;
        st.last_success = 0;
        ++st.n_unsuccess;
        if (st.num_invalid >= opt.max_consecutive_invalid) { st.status = 1; st.termination = 5; break; }
        st.radius = st.radius / st.decrease_factor;
        st.decrease_factor *= 2.0;
        if (st.iteration >= opt.max_iters) { st.status = 1; st.termination = 0; break; }
    }
    (void)step_dot_g; (void)lm_quad;
    if (!have_step) {
        if (lane == 0) { a.state[w] = st; a.win_status[w] = st.status; }
        return;
    }

    // ---- (5) candidate = Plus(x, step * scale) and its laser frame tables
    for (int f = lane; f < n; f += 32) {
        double d[15];
#pragma unroll
        for (int c = 0; c < 15; ++c) d[c] = S.b[f * 15 + c] * S.sc[f * 15 + c];
#pragma unroll
        for (int c = 0; c < 15; ++c) xc[15 * f + c] = is_const(f, c) ? x[15 * f + c] : x[15 * f + c] + d[c];
        if (!is_const(f, 3)) so3_plus(x + 15 * f + 3, d + 3, xc + 15 * f + 3);
        laser_frame_table(a.C, xc + 15 * f, a.frame_tab + ((size_t)w * n + f) * kFrameTab);
    }
    if (lane == 0) { a.state[w] = st; a.win_status[w] = 0; }
}

template <bool ARROW>
__global__ void __launch_bounds__(128) window_step_kernel(WindowArgs a, int per_warp_doubles) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + warp;
    if (w >= a.n_windows) return;
    window_step<ARROW>(a, w, lane, smem + (size_t)warp * per_warp_doubles);
}

// frame tables of arbitrary poses ([F][6] -> [F][24]); used for the initial point and the external reference poses
__global__ void frame_table_kernel(Consts C, const double* poses, int stride, double* tabs, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    laser_frame_table(C, poses + (size_t)f * stride, tabs + (size_t)f * kFrameTab);
}

}  // namespace lv
