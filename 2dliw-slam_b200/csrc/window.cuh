// window.cuh — everything of an LM iteration that is NOT the scan-point pass, as two kernels:
//
//   factor_pair_kernel  the residual blocks hanging on frame i of a window — IMU and wheel factor of the pair (i-1, i),
//                  `multiplicity` ground factors and the marginalisation prior of frame i — evaluated AT THE CANDIDATE
//                  point: whitened Jacobians, J^T J blocks, J^T r, r^2.  Two items per warp (12 dual columns + 15
//                  closed-form columns each); factor_kernel is the one-item-per-warp, all-dual variant kept for A/B.
//   window_kernel  Ceres' minimiser logic + the linear solve; one warp per window for large batches, four or eight warps
//                  per window (template parameter NT) for small ones.
//   solve_small_kernel  the whole loop (scan-match items, factor pairs, window step) in one launch for tiny windows.
//
// Together they replace ceres::Solve as configured by solver::solve / solver::do_init_solve (reference
// src/factor/solver.cpp:795-802, :161-168) — Ceres 1.14 TrustRegionMinimizer + LevenbergMarquardtStrategy with Jacobi
// scaling and an exact Schur/Cholesky step — and the residual blocks Ceres would evaluate with Jets: imu_factor
// (imu_factor.h:13-89), wheel_odom_factor (wheel_factor.h:12-73), ground_factor_p/q (ground_factor.h:27-82; added
// `multiplicity` times, solver.cpp:727-743), marginalization_factor (marginalization_factor.h:22-53) under the
// constness rules of solver.cpp:787-794.
//
// One trip of the minimiser loop = scan_match_kernel + factor_pair_kernel (both at the candidate, independent of each
// other) followed by window_kernel:
//   (1) candidate cost = laser tiles + factor items; Ceres' parameter-/function-tolerance tests, then the
//       step-quality test rho > 1e-3; accept (flip the per-window buffer parity) or reject; trust-region update;
//   (2) assemble the block-tridiagonal (+ frame-0 arrow in the initialisation topology) normal equations of the
//       accepted point, Jacobi scaling, LM damping;
//   (3) block elimination in reverse frame order (no fill-in for either topology) with Gauss-Jordan inverses of
//       the 15x15 pivots (rows in registers, no serial triangular solves), step, model cost change; invalid steps
//       shrink the radius and retry in place;
//   (4) candidate = Plus(x, step) (so3_parameterization: wrap(theta + delta), factor_common.h:40-53) and its laser
//       frame tables for the next scan-match pass.
// Because the candidate's linearisation is produced together with its cost, an accepted step costs one pass over
// the data, not two (Ceres: cost-only evaluation, then a Jacobian evaluation at the same point).
//
// Layout of every 15x15 block: packed row-major, 225 doubles (conflict-free for lane-per-row access).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lv_math.cuh"
#include "scan_match.cuh"

#ifndef LV_FACTOR_MINB
#define LV_FACTOR_MINB 4   // resident CTAs (of 4 warps) per SM the register budget of factor_kernel is sized for
#endif
#ifndef LV_PAIR_MINB
#define LV_PAIR_MINB 3     // resident CTAs per SM the register budget of factor_pair_kernel is sized for
#endif
#ifndef LV_WINDOW_WPC
#define LV_WINDOW_WPC 2    // windows (warps) per CTA of the one-warp-per-window shape
#endif
#ifndef LV_WINDOW_WARPS
#define LV_WINDOW_WARPS 16 // resident warps per SM the register budget of the one-warp-per-window shape is sized for
#endif
#ifndef LV_PAIR_WPC
#define LV_PAIR_WPC 4      // warps (item pairs) per CTA of factor_pair_kernel
#endif
#ifndef LV_FACTOR_ROLL
#define LV_FACTOR_ROLL 1   // 1: keep the column loop of the J^T J product rolled (smaller instruction footprint)
#endif

namespace lv {

struct LMState {
    double radius, decrease_factor, cost, model_cost_change, x_norm, initial_cost;
    int32_t iteration, num_invalid, status, termination, n_success, n_unsuccess, last_success, started;
    int32_t cur, pad0;  // parity of the buffers that hold the linearisation of the accepted point
};

struct LMOptions {
    int32_t max_iters;
    double function_tolerance, gradient_tolerance, parameter_tolerance, initial_radius;
    double max_radius, min_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
    int32_t max_consecutive_invalid;
};

constexpr int kBlk = 225;
// Factor item record = the symmetric 32 x 32 product W^T W of the item's whitened Jacobian | residual matrix W with the
// columns ordered [frame i-1 (0..14) | zero (15) | frame i (16..30) | whitened residual (31)], i.e. J^T J, J^T r (column
// 31) and r^T r (entry 31,31) in one matrix, stored as its 10 upper 8 x 8 tiles (64 doubles each, row-major) grouped by
// the 16 x 16 block the window kernel consumes:
//   haa = H(i-1, i-1): tiles (0,0) (0,1) (1,1)          at   0,  64, 128
//   hab = H(i-1, i)  : tiles (0,2) (0,3) (1,2) (1,3)    at 192, 256, 320, 384   (its column 15 is g_a)
//   hbb = H(i, i)    : tiles (2,2) (2,3) (3,3)          at 448, 512, 576        (its column 15 is g_b, entry (15,15) the cost)
// — exactly what the DMMA accumulators of the factor kernels hold, written with one 16-byte store per lane and tile, and
// read by window_kernel as contiguous 16-byte chunks straight into its tile-form blocks.  640 doubles per item (the
// dense 30 x 30 + g record of round 1 was 936).
constexpr int kItem = 640;
__host__ __device__ __forceinline__ int item_tile_slot(int ti, int tj) {   // upper tiles only (ti <= tj)
    return ti == 0 ? (tj == 0 ? 0 : (tj == 1 ? 1 : tj + 1)) : (ti == 1 ? (tj == 1 ? 2 : tj + 3) : (ti == 2 ? tj + 5 : 9));
}
__host__ __device__ __forceinline__ int item_idx(int r, int c) {   // entry (r, c) = (c, r) of the 32 x 32 matrix
    if (r > c) { const int t = r; r = c; c = t; }
    return 64 * item_tile_slot(r >> 3, c >> 3) + (r & 7) * 8 + (c & 7);
}
__device__ __forceinline__ int item_haa(int r, int c) { return item_idx(r, c); }
__device__ __forceinline__ int item_hab(int r, int c) { return item_idx(r, 16 + c); }
__device__ __forceinline__ int item_hbb(int r, int c) { return item_idx(16 + r, 16 + c); }
__device__ __forceinline__ int item_ga(int c) { return item_idx(c, 31); }
__device__ __forceinline__ int item_gb(int c) { return item_idx(16 + c, 31); }
constexpr int kItemCost = 64 * 9 + 63;   // item_idx(31, 31)
constexpr int kItemHab = 192, kItemHbb = 448;

struct WindowArgs {
    Consts C;
    LMOptions opt;
    int32_t n_windows, n_frames, tiles, arrow, mode;  // mode 0: solver program, 1: marginalisation program
    int32_t ground_multiplicity, prior_frame, has_imu, has_wheel;
    int32_t imu_stride;             // 466: ABI blobs; 190: the compact records of lvio2d_scan_wire::imu_compact
    const uint8_t* const_mask;      // [B*n]
    const uint8_t* frame_active;    // [B*n] laser block active
    const int32_t* ref_frame;       // [B*n] or nullptr
    const double* imu;              // [B*(n-1)][466]
    const double* wheel;            // [B*(n-1)][15]
    const double* prior_X0;         // [B][15]
    const double* prior_J;          // [B][225]
    const double* prior_H;          // [B][225]  J^T J of the prior, computed once per upload (prior_info_kernel)
    const double* partial;          // [B*n][tiles][pad]   scan-match output at the candidate
    double* x;                      // [B*n][15] accepted point
    double* xc;                     // [B*n][15] candidate
    double* scale;                  // [B*n][15] Jacobi scaling
    double* laser_blocks;           // [2][B*n][pad]  parity-buffered laser blocks
    double* items;                  // [2][B*n][kItem] parity-buffered factor items
    double* frame_tab;              // [B*n][24] tables of the candidate
    double* fac;                    // [B*n][3][225]  T_i | T'_i | -
    double* vec;                    // [B][2][n*15]  gradient scratch
    LMState* state;                 // [B]
    int32_t* win_status;            // [B] mirror of state.status for the other kernels
    // optional dense outputs (lvio2d_linearize): H [B][15n][15n], g [B][15n], cost [B]
    double* dense_H;
    double* dense_g;
    double* dense_cost;
    // optional marginalisation outputs
    double* marg_H;                 // [B][225] Schur complement on the last frame
    double* marg_g;                 // [B][15]
    int cr_scratch;                 // cyclic-reduction shape (NT = 512): scratch blocks behind the fixed arrays (0..13)
};

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
__device__ __forceinline__ bool col_const(uint8_t mask, int c) {  // c: column 0..14 of [p q v bs]
    const int blk = c < 3 ? 0 : (c < 6 ? 1 : (c < 9 ? 2 : 3));
    return (mask >> blk) & 1;
}

// asynchronous global -> shared copies (LDGSTS): inputs are fetched while independent arithmetic runs (the IMU blob
// behind the two exponentials in factor_kernel, the raw blocks of the next frame behind the Gauss-Jordan inverse of
// the current pivot in window_kernel)
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// fp64 tensor-core product D(8x8) += A(8x4) B(4x8) (DMMA.8x8x4 on sm_100a, measured at 37 TFLOP/s against 33.8 for the
// vector pipe, at one eighth of the issue slots per FMA).  Lane l = 4 g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// =====================================================================================================
// The factor item: the residual blocks hanging on frame i of a window, evaluated at the candidate point.
//
// Shared-memory matrix W of one item, [kWRows][kWS] doubles (row stride 34: 16-byte aligned rows for the DMMA result
// stores, at most 2-way bank conflicts on the fragment loads):
//   columns      0..14 frame i-1 | 15 zero | 16..30 frame i | 31 the residual
//   rows  0..14  IMU factor: Jacobian | residual, un-whitened after phase A, whitened in place by item_finish
//   rows 15..17  wheel factor (already weighted by diag(sqrt_inverse_P), wheel_factor.h:58-70)
//   rows 18..19  the two ground factors times sqrt(multiplicity) (the reference adds them n times, solver.cpp:727-743)
// Constant parameter blocks have zero columns (Ceres removes them from the reduced program, solver.cpp:787-794);
// residual blocks whose parameter blocks are all constant have zero rows (their cost is Ceres' fixed_cost).
constexpr int kWS = 34, kWRows = 20, kWSize = kWRows * kWS;
constexpr int kCB = 72;   // compact IMU blob: X[15] | pad | J(0..8, 9..14) [9][6] at 16 | Dt at 70
__device__ __forceinline__ void stage_compact_blob(double* sC, const double* blob, int sl, int nsl, bool compact) {
    for (int k = sl; k < 70; k += nsl) {
        const int src = compact ? (k < 69 ? k : 189) : (k < 15 ? k : (k < 69 ? 15 + ((k - 15) / 6) * 15 + 9 + (k - 15) % 6 : 465));
        cp_async8(sC + (k < 15 ? k : (k < 69 ? k + 1 : 70)), blob + src);
    }
}
// entry (r, k), k >= r, of the upper-triangular sqrt_inverse_P: row-major 15 x 15 in the ABI blob, packed rows in the compact record
__device__ __forceinline__ double sqrt_info_at(const double* sq, bool compact, int r, int k) {
    return __ldg(sq + (compact ? r * 15 - (r * (r - 1)) / 2 + (k - r) : r * 15 + k));
}

// Whole warp.  Whitening (when sq != nullptr: rows 0..14 <- sq rows 0..14, upper triangular sqrt_inverse_P read from
// global memory as DMMA A fragments), W^T W on the tensor pipe, the marginalisation prior of frame i (r = J (x - X0),
// marginalization_factor.h:50: H += J^T J, g += J^T r, cost += r^T r from the precomputed J^T J), one 16-byte store per
// lane and tile.
__device__ __forceinline__ void item_finish(const WindowArgs& a, double* W, const double* sq, const bool sq_packed, const int w, const int i, const uint8_t mb,
                                            const bool prior_on, double* sPg /* 16 doubles */, double* out, const int lane) {
    const int g = lane >> 2, t = lane & 3;
    if (sq) {
        double af[6], bf[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int k = 4 * ks + t;
            af[ks] = (k < 15 && k >= g) ? sqrt_info_at(sq, sq_packed, g, k) : 0.0;
        }
#pragma unroll
        for (int ks = 2; ks < 4; ++ks) {
            const int r = 8 + g, k = 4 * ks + t;
            af[2 + ks] = (r < 15 && k < 15 && k >= r) ? sqrt_info_at(sq, sq_packed, r, k) : 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) bf[ks][ni] = (4 * ks + t < 15) ? W[(4 * ks + t) * kWS + 8 * ni + g] : 0.0;
        __syncwarp();
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) dmma884(c0, c1, af[ks], bf[ks][ni]);
#pragma unroll
            for (int ks = 2; ks < 4; ++ks) dmma884(d0, d1, af[2 + ks], bf[ks][ni]);
            *reinterpret_cast<double2*>(W + g * kWS + 8 * ni + 2 * t) = make_double2(c0, c1);
            if (8 + g < 15) *reinterpret_cast<double2*>(W + (8 + g) * kWS + 8 * ni + 2 * t) = make_double2(d0, d1);
        }
        __syncwarp();
    }
    double acc[10][2];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k][0] = acc[k][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < kWRows / 4; ++ks) {
        double f[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) f[c] = W[(4 * ks + t) * kWS + 8 * c + g];
        int tile = 0;
#pragma unroll
        for (int ti = 0; ti < 4; ++ti)
#pragma unroll
            for (int tj = ti; tj < 4; ++tj) { dmma884(acc[tile][0], acc[tile][1], f[ti], f[tj]); ++tile; }
    }
    if (prior_on) {
        const double* PH = a.prior_H + (size_t)w * kBlk;
        const double* X0 = a.prior_X0 + (size_t)w * 15;
        const double* xb = a.xc + ((size_t)w * a.n_frames + i) * 15;
        double gp = 0.0, dl = 0.0;
        if (lane < 15) {
#pragma unroll
            for (int k = 0; k < 15; ++k) gp += PH[lane * 15 + k] * (xb[k] - X0[k]);
            dl = xb[lane] - X0[lane];
            sPg[lane] = gp;
        }
        const double cost_p = warp_sum(dl * gp);
        __syncwarp();
        int tile = 0;
#pragma unroll
        for (int ti = 0; ti < 4; ++ti)
#pragma unroll
            for (int tj = ti; tj < 4; ++tj) {
                const int r = 8 * ti + g - 16;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = 8 * tj + 2 * t + q - 16;
                    if (r >= 0 && r < 15 && !col_const(mb, r)) {
                        if (c >= 0 && c < 15 && !col_const(mb, c)) acc[tile][q] += PH[r * 15 + c];
                        else if (c == 15) acc[tile][q] += sPg[r];
                    } else if (r == 15 && c == 15) acc[tile][q] += cost_p;
                }
                ++tile;
            }
        __syncwarp();   // sPg may be reused by the next item
    }
    {
        int tile = 0;
#pragma unroll
        for (int ti = 0; ti < 4; ++ti)
#pragma unroll
            for (int tj = ti; tj < 4; ++tj) {
                *reinterpret_cast<double2*>(out + item_tile_slot(ti, tj) * 64 + g * 8 + 2 * t) = make_double2(acc[tile][0], acc[tile][1]);
                ++tile;
            }
    }
}

// J^T J of the marginalisation prior, once per upload: prior_H[w] = prior_J[w]^T prior_J[w]
__global__ void prior_info_kernel(const double* prior_J, double* prior_H, int n_windows) {
    const int w = blockIdx.x;
    if (w >= n_windows) return;
    const double* J = prior_J + (size_t)w * kBlk;
    for (int e = threadIdx.x; e < kBlk; e += blockDim.x) {
        const int r = e / 15, c = e - 15 * r;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 15; ++k) s += J[k * 15 + r] * J[k * 15 + c];
        prior_H[(size_t)w * kBlk + e] = s;
    }
}

// =====================================================================================================
// factor_kernel: one item per warp, every one of the 30 state columns in dual arithmetic (lane c = column c, lane 30 the
// values), whitened per lane.  The all-dual cross-check of factor_pair_kernel's closed-form columns (LVIO2D_FACTOR_PAIRED=0).
// Per-warp shared memory: blob 480 | W 680 | wheel blob 16 | prior scratch 16
constexpr int kFactorSmem = 480 + kWSize + 16 + 16;
__global__ void __launch_bounds__(128, LV_FACTOR_MINB) factor_kernel(WindowArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = a.n_frames;
    const int item = blockIdx.x * (blockDim.x >> 5) + warp;
    if (item >= a.n_windows * n) return;
    const int w = item / n, i = item - w * n;
    if (a.win_status[w] != 0) return;
    const int mode = a.mode;
    const int parity = 1 - a.state[w].cur;
    double* out = a.items + ((size_t)parity * a.n_windows * n + item) * kItem;
    double* sblob = smem + (size_t)warp * kFactorSmem;
    double* W = sblob + 480;
    double* sWb = W + kWSize;   // wheel blob [15]
    double* sPg = sWb + 16;
    const double* X = a.xc + (size_t)w * n * 15;
    const uint8_t* cm = a.const_mask + (size_t)w * n;
    const uint8_t mb = mode == 1 ? 0 : cm[i];
    const uint8_t ma = (mode == 1 || i == 0) ? 0 : cm[i - 1];
    const double* xb = X + 15 * i;
    const double* xa = X + 15 * (i > 0 ? i - 1 : 0);
    const bool imu_on = i > 0 && a.has_imu && ((ma & 15) != 15 || (mb & 15) != 15);
    const bool wheel_on = i > 0 && a.has_wheel && ((ma & 3) != 3 || (mb & 3) != 3);
    const bool ground_on = a.ground_multiplicity > 0 && (mb & 3) != 3;
    const bool prior_on = a.prior_frame == i && (mb & 15) != 15;
    if (imu_on) {
        const double* blob = a.imu + ((size_t)w * (n - 1) + (i - 1)) * a.imu_stride;
        if (a.imu_stride != 466) {
            // compact record -> the blob layout item_imu reads (entries imu_factor never reads stay zero)
            for (int k = lane; k < 466; k += 32) {
                double v = 0.0;
                if (k < 15) v = blob[k];
                else if (k < 240) { const int r = (k - 15) / 15, q = (k - 15) % 15; if (r < 9 && q >= 9) v = blob[15 + r * 6 + (q - 9)]; }
                else if (k < 465) { const int r = (k - 240) / 15, q = (k - 240) % 15; if (q >= r) v = blob[69 + r * 15 - (r * (r - 1)) / 2 + (q - r)]; }
                else v = blob[189];
                sblob[k] = v;
            }
        } else if ((reinterpret_cast<uintptr_t>(blob) & 15) == 0) {
            for (int k = lane; k < 233; k += 32) cp_async16(sblob + 2 * k, blob + 2 * k);
        } else {
            for (int k = lane; k < 466; k += 32) cp_async8(sblob + k, blob + k);
        }
    }
    if (wheel_on && lane < 15) cp_async8(sWb + lane, a.wheel + ((size_t)w * (n - 1) + (i - 1)) * 15 + lane);
    for (int k = lane; k < kWSize; k += 32) W[k] = 0.0;
    {
        const FrameState<Dual> fa_ = seed_frame_state(xa, lane < 15 ? lane : -1);
        const FrameState<Dual> fb_ = seed_frame_state(xb, (lane >= 15 && lane < 30) ? lane - 15 : -1);
        const M3<Dual> Rj = exp_so3(fb_.th);
        M3<Dual> Ri;
        if (imu_on || wheel_on) Ri = exp_so3(fa_.th);
        cp_async_wait_all();
        __syncwarp();
        const bool value_lane = lane == 30;
        const bool dead = lane < 30 && col_const(lane < 15 ? ma : mb, lane % 15);
        const double sm = sqrt((double)a.ground_multiplicity);
        const int wcol = lane < 15 ? lane : (lane < 30 ? lane + 1 : (lane == 30 ? 31 : 15));   // column of W this lane owns
        if (imu_on) {
            Dual ri[15];
            item_imu<Dual>(a.C, sblob, fa_, fb_, Ri, Rj, ri);
            const double* Sq = sblob + 240;  // sqrt_inverse_P = L^T: upper triangular
#pragma unroll
            for (int r = 0; r < 15; ++r) {
                double s = 0.0;
#pragma unroll
                for (int k = r; k < 15; ++k) s += Sq[r * 15 + k] * (value_lane ? ri[k].a : ri[k].d);
                W[r * kWS + wcol] = (dead || lane == 31) ? 0.0 : s;
            }
        }
        const int fl = lane % 15;
        if (wheel_on) {
            Dual rw[3];
            item_wheel<Dual>(a.C, sWb, fa_.p, fb_.p, Ri, Rj, rw);
            if (lane < 30 && fl < 6) {
#pragma unroll
                for (int k = 0; k < 3; ++k) W[(15 + k) * kWS + wcol] = dead ? 0.0 : rw[k].d;
            } else if (value_lane) {
#pragma unroll
                for (int k = 0; k < 3; ++k) W[(15 + k) * kWS + 31] = rw[k].a;
            }
        }
        if (ground_on) {
            Dual rg[2];
            item_ground<Dual>(a.C, fb_.p, Rj, rg);
            if (lane >= 15 && lane < 21) {
                W[18 * kWS + wcol] = dead ? 0.0 : sm * rg[0].d;
                W[19 * kWS + wcol] = dead ? 0.0 : sm * rg[1].d;
            } else if (value_lane) {
                W[18 * kWS + 31] = sm * rg[0].a;
                W[19 * kWS + 31] = sm * rg[1].a;
            }
        }
    }
    __syncwarp();
    item_finish(a, W, nullptr, false, w, i, mb, prior_on, sPg, out, lane);
}

// =====================================================================================================
// factor_pair_kernel: the same items, TWO per warp.
//
// Of the 30 state columns of an item only 12 need dual-number arithmetic: theta_a, theta_b, bw_a (the rotation chain
// of the IMU residual) and p_b (the norm-type wheel residuals; d/dp_a = -d/dp_b exactly, both enter through
// t_oj - t_oi).  The other 15 columns (v_a, ba_a, v_b, ba_b, bw_b) only touch the IMU factor, linearly:
//   d r_alpha / d v_a = R_i^T Dt,  d r_beta / d v_a = R_i^T,  d r_beta / d v_b = -R_i^T,
//   d r_{alpha,beta} / d ba_a = J[:, 9..11],  d r_ba / d ba_{a,b} = -+I,  d r_bw / d bw_b = I
// — exactly what the dual path produces for them (its products with the zero dual parts add exact zeros).
// Phase A, a half-warp per item: sub-lane 0 the values, 1..3 p_b, 4..6 theta_a, 7..9 theta_b, 10..12 bw_a evaluate the
// residuals in dual arithmetic and drop their UN-whitened column into W; 15 sub-lanes drop the closed-form columns.
// Phase B, the whole warp, one item after the other (item_finish): whitening by sqrt_inverse_P and the 32 x 32 product
// W^T W on the fp64 tensor pipe — 24 + 50 DMMA per item where round 1 issued ~700 DFMA + ~450 shared-memory loads.
// Per-warp shared memory: 2 x (compact blob 72 | W 680 | wheel blob 16) + prior scratch 16 = 12.3 KB (round 1: 18.7 KB).
constexpr int kPairHalf = kCB + kWSize + 16;
constexpr int kPairSmem = 2 * kPairHalf + 16;
// items `first` and `first + 1` (the second only when count == 2) on one warp; `base` = kPairSmem doubles of shared memory
__device__ __forceinline__ void factor_pair_item(const WindowArgs& a, const int first, const int count, const int lane, double* base) {
    const int h = lane >> 4, sl = lane & 15;
    const int n = a.n_frames, total = a.n_windows * n;
    const int item = first + h;
    int w = 0, i = 0;
    bool live = h < count && item < total;
    if (live) { w = item / n; i = item - w * n; live = a.win_status[w] == 0; }
    if (!__any_sync(0xffffffffu, live)) return;
    const int mode = a.mode;
    double* sC = base + h * kPairHalf;   // compact IMU blob
    double* W = sC + kCB;
    double* sWb = W + kWSize;            // wheel blob [15]
    double* sPg = base + 2 * kPairHalf;
    const double* X = a.xc + (size_t)w * n * 15;
    const uint8_t* cm = a.const_mask + (size_t)w * n;
    const uint8_t mb = (mode == 1 || !live) ? 0 : cm[i];
    const uint8_t ma = (mode == 1 || i == 0 || !live) ? 0 : cm[i - 1];
    const double* xb = X + 15 * i;
    const double* xa = X + 15 * (i > 0 ? i - 1 : 0);
    const bool imu_on = live && i > 0 && a.has_imu && ((ma & 15) != 15 || (mb & 15) != 15);
    const bool wheel_on = live && i > 0 && a.has_wheel && ((ma & 3) != 3 || (mb & 3) != 3);
    const bool ground_on = live && a.ground_multiplicity > 0 && (mb & 3) != 3;
    const bool prior_on = live && a.prior_frame == i && (mb & 15) != 15;
    const bool compact = a.imu_stride != 466;
    const double* blob = a.imu + ((size_t)w * (n - 1) + (i > 0 ? i - 1 : 0)) * a.imu_stride;
    // ---- phase A: both items at once.  Constant inputs travel to shared memory while the exponentials are evaluated.
    if (imu_on) stage_compact_blob(sC, blob, sl, 16, compact);
    if (wheel_on && sl < 15) cp_async8(sWb + sl, a.wheel + ((size_t)w * (n - 1) + (i - 1)) * 15 + sl);
    // rows 15..19 (and the IMU rows of an item without IMU factor) start from zero
    for (int k = (imu_on ? 15 * kWS : 0) + sl; k < kWSize; k += 16) W[k] = 0.0;
    // sub-lane roles
    const bool value_lane = sl == 0;
    const bool l_pb = sl >= 1 && sl <= 3, l_ta = sl >= 4 && sl <= 6, l_tb = sl >= 7 && sl <= 9, l_bw = sl >= 10 && sl <= 12;
    const int kk = l_pb ? sl - 1 : (l_ta ? sl - 4 : (l_tb ? sl - 7 : (l_bw ? sl - 10 : 0)));   // component 0..2
    const int seed_a = l_ta ? 3 + kk : (l_bw ? 12 + kk : -1);
    const int seed_b = l_pb ? kk : (l_tb ? 3 + kk : -1);
    // frame-a / frame-b column this lane's derivative belongs to (p_b lanes also own the negated p_a column)
    const int col_a = l_ta ? 3 + kk : (l_bw ? 12 + kk : (l_pb ? kk : -1));
    const int col_b = l_pb ? kk : (l_tb ? 3 + kk : -1);
    const bool dead_a = col_a >= 0 && col_const(ma, col_a);
    const bool dead_b = col_b >= 0 && col_const(mb, col_b);
    {
        const FrameState<Dual> fa_ = seed_frame_state(xa, seed_a);
        const FrameState<Dual> fb_ = seed_frame_state(xb, seed_b);
        const M3<Dual> Rj = exp_so3(fb_.th);
        M3<Dual> Ri;
        if (imu_on || wheel_on) Ri = exp_so3(fa_.th);
        cp_async_wait_all();
        __syncwarp();
        if (imu_on) {
            Dual ri[15];
            item_imu_t<Dual, 6, 0>(a.C, sC, sC + 16, sC[70], fa_, fb_, Ri, Rj, ri);
            if (value_lane) {
#pragma unroll
                for (int r = 0; r < 15; ++r) { W[r * kWS + 31] = ri[r].a; W[r * kWS + 15] = 0.0; }
            } else if (sl <= 12) {
#pragma unroll
                for (int r = 0; r < 15; ++r) {
                    const double s = ri[r].d;
                    if (col_a >= 0) W[r * kWS + col_a] = dead_a ? 0.0 : (l_pb ? -s : s);
                    if (col_b >= 0) W[r * kWS + 16 + col_b] = dead_b ? 0.0 : s;
                }
            }
            // the 15 closed-form columns, one per sub-lane: t = 0 v_a, 1 ba_a, 2 v_b, 3 ba_b, 4 bw_b; component k
            if (sl < 15) {
                const int t = sl / 3, k = sl - 3 * t;
                const double Dt = sC[70];
                const double* Jc = sC + 16;
                double rik[3];   // row k of R_i (= column k of R_i^T)
#pragma unroll
                for (int j = 0; j < 3; ++j) rik[j] = k == 0 ? Ri.m[j].a : (k == 1 ? Ri.m[3 + j].a : Ri.m[6 + j].a);
                // column of W: v_a 6.., ba_a 9.., v_b 16 + 6.., ba_b 16 + 9.., bw_b 16 + 12..
                const int c = (t == 0 ? 6 : (t == 1 ? 9 : (t == 2 ? 22 : (t == 3 ? 25 : 28)))) + k;
                const bool dead = col_const(c < 15 ? ma : mb, c < 15 ? c : c - 16);
#pragma unroll
                for (int r = 0; r < 15; ++r) {
                    double v = 0.0;
                    if (r < 3) v = t == 0 ? rik[r] * Dt : (t == 1 ? Jc[r * 6 + k] : 0.0);
                    else if (r < 6) v = t == 0 ? rik[r - 3] : (t == 1 ? Jc[r * 6 + k] : (t == 2 ? -rik[r - 3] : 0.0));
                    else if (r >= 9 && r < 12) v = (r - 9 == k) ? (t == 1 ? -1.0 : (t == 3 ? 1.0 : 0.0)) : 0.0;
                    else if (r >= 12) v = (r - 12 == k && t == 4) ? 1.0 : 0.0;
                    W[r * kWS + c] = dead ? 0.0 : v;
                }
            }
        }
        if (wheel_on) {
            Dual rw[3];
            item_wheel<Dual>(a.C, sWb, fa_.p, fb_.p, Ri, Rj, rw);
            if (value_lane) {
#pragma unroll
                for (int k = 0; k < 3; ++k) W[(15 + k) * kWS + 31] = rw[k].a;
            } else if (l_pb || l_ta || l_tb) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (col_a >= 0) W[(15 + k) * kWS + col_a] = dead_a ? 0.0 : (l_pb ? -rw[k].d : rw[k].d);
                    if (col_b >= 0) W[(15 + k) * kWS + 16 + col_b] = dead_b ? 0.0 : rw[k].d;
                }
            }
        }
        if (ground_on) {
            Dual rg[2];
            item_ground<Dual>(a.C, fb_.p, Rj, rg);
            const double sm = sqrt((double)a.ground_multiplicity);
            if (value_lane) {
                W[18 * kWS + 31] = sm * rg[0].a;
                W[19 * kWS + 31] = sm * rg[1].a;
            } else if (col_b >= 0) {
                W[18 * kWS + 16 + col_b] = dead_b ? 0.0 : sm * rg[0].d;
                W[19 * kWS + 16 + col_b] = dead_b ? 0.0 : sm * rg[1].d;
            }
        }
    }
    __syncwarp();
    // ---- phase B: whitening and W^T W of the two items, one after the other on the full warp
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
        const int src = 16 * hh;
        if (!__shfl_sync(0xffffffffu, (int)live, src)) continue;
        const int wv = __shfl_sync(0xffffffffu, w, src), iv = __shfl_sync(0xffffffffu, i, src);
        const uint8_t mbv = (uint8_t)__shfl_sync(0xffffffffu, (int)mb, src);
        const bool prior_v = __shfl_sync(0xffffffffu, (int)prior_on, src) != 0;
        const bool imu_v = __shfl_sync(0xffffffffu, (int)imu_on, src) != 0;
        const int parity = 1 - a.state[wv].cur;
        double* out = a.items + ((size_t)parity * total + (size_t)wv * n + iv) * kItem;
        const double* sq = imu_v ? a.imu + ((size_t)wv * (n - 1) + (iv - 1)) * a.imu_stride + (compact ? 69 : 240) : nullptr;
        item_finish(a, base + hh * kPairHalf + kCB, sq, compact, wv, iv, mbv, prior_v, sPg, out, lane);
    }
}

__global__ void __launch_bounds__(32 * LV_PAIR_WPC, LV_PAIR_MINB * 4 / LV_PAIR_WPC) factor_pair_kernel(WindowArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = blockIdx.x * (blockDim.x >> 5) + warp;
    factor_pair_item(a, 2 * pair, 2, lane, smem + (size_t)warp * kPairSmem);
}

// =====================================================================================================
// warp-cooperative 15x15 kernels on shared memory (lane r owns row r; lanes >= 15 idle)

// In-place inverse of an SPD block by Gauss-Jordan elimination without pivoting.  30 lanes: lane = (row r = lane & 15,
// column half h = lane >> 4) keeps 8 (h = 0: columns 0..7) or 7 (h = 1: columns 8..14) entries of its row in registers.
// Per pivot k the two lanes of row k publish the scaled pivot row, every lane updates its half row.
// Returns false (uniformly) when a pivot is not positive (the block is not positive definite) or not finite.
__device__ __noinline__ bool spd_inverse15(double* A, double* piv /* 16 doubles scratch */, int lane) {
    const int r = lane & 15, h = lane >> 4, c0 = h * 8;
    const bool act = r < 15;
    double row[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) row[j] = (act && c0 + j < 15) ? A[r * 15 + c0 + j] : 0.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        const int kh = k >> 3, kj = k & 7;                    // which half / register holds column k
        // the pivot and this row's multiplier live in the half kh: broadcast them
        const double p = __shfl_sync(0xffffffffu, row[kj], k + 16 * kh);
        const double f = __shfl_sync(0xffffffffu, row[kj], r + 16 * kh);
        if (!(p > 0.0) || !isfinite(p)) ok = false;
        const double pinv = 1.0 / p;
        if (r == k && act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { row[j] = (c0 + j == k) ? pinv : row[j] * pinv; piv[c0 + j] = row[j]; }
        }
        __syncwarp();
        if (r != k && act) {
            const double g = f;
#pragma unroll
            for (int j = 0; j < 8; ++j) row[j] = (c0 + j == k) ? -g * pinv : row[j] - g * piv[c0 + j];
        }
        __syncwarp();
    }
    if (act) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (c0 + j < 15) A[r * 15 + c0 + j] = row[j];
    }
    __syncwarp();
    return ok;
}
// Thread groups: one window is solved by NT threads — one warp (NT = 32, the batched shape: as many windows in flight as
// possible) or four warps (NT = 128, small batches: the same work spread four ways to cut the latency of one solve).
template <int NT> __device__ __forceinline__ void grp_sync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}
// reductions over the group; `red` = 16 doubles of shared scratch (NT > 32 only); every thread gets the result, and the
// cross-warp part is summed in a fixed order (deterministic)
template <int NT> __device__ __forceinline__ double grp_sum(double v, double* red, int tid) {
    v = warp_sum(v);
    if (NT == 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) s += red[k];
    return s;
}
template <int NT> __device__ __forceinline__ double grp_max(double v, double* red, int tid) {
    v = warp_max(v);
    if (NT == 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int k = 1; k < NT / 32; ++k) s = fmax(s, red[k]);
    return s;
}
template <int NT> __device__ __forceinline__ bool grp_all(bool v, double* red, int tid) {
    if (NT == 32) return __all_sync(0xffffffffu, v);
    return __syncthreads_and(v) != 0;
}
// SPD inverse by the first warp of the group, result flag broadcast through `red[8]`
template <int NT> __device__ __forceinline__ bool grp_inverse15(double* A, double* piv, double* red, int tid) {
    if (NT == 32) return spd_inverse15(A, piv, tid);
    if (tid < 32) {
        const bool ok = spd_inverse15(A, piv, tid);
        if (tid == 0) red[8] = ok ? 1.0 : 0.0;
    }
    __syncthreads();
    return red[8] != 0.0;
}
// The two 15 x 15 x 15 products run on the fp64 tensor pipe: padded to 16 x 16 x 16 = 2 x 2 tiles x 4 k-steps of
// DMMA.8x8x4, operands read straight from the packed 15-stride blocks (index 15 reads as zero).  With more than one
// warp per window the four tiles are dealt to the warps.  16 DMMA + 16 shared-memory loads per product where the
// vector-pipe version issued 120 DFMA + 135 loads.
// C = A B
template <int NT>
__device__ __noinline__ void gemm_ab15(double* Cm, const double* A, const double* B, int tid) {
    constexpr int NW = NT / 32;
    const int lane = tid & 31, wi = tid >> 5, g = lane >> 2, t = lane & 3;
    if (NW == 1) {
        double af[2][4], bf[4][2];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int k = 4 * ks + t;
            af[0][ks] = k < 15 ? A[g * 15 + k] : 0.0;
            af[1][ks] = (k < 15 && g < 7) ? A[(8 + g) * 15 + k] : 0.0;
            bf[ks][0] = k < 15 ? B[k * 15 + g] : 0.0;
            bf[ks][1] = (k < 15 && g < 7) ? B[k * 15 + 8 + g] : 0.0;
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) dmma884(c0, c1, af[mi][ks], bf[ks][ni]);
                const int r = 8 * mi + g, c = 8 * ni + 2 * t;
                if (r < 15) { Cm[r * 15 + c] = c0; if (c + 1 < 15) Cm[r * 15 + c + 1] = c1; }
            }
    } else {
#pragma unroll
        for (int tl = 0; tl < 4; ++tl) {
            if ((tl % NW) != wi) continue;
            const int mi = tl >> 1, ni = tl & 1, r = 8 * mi + g;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int k = 4 * ks + t;
                dmma884(c0, c1, (r < 15 && k < 15) ? A[r * 15 + k] : 0.0, (k < 15 && 8 * ni + g < 15) ? B[k * 15 + 8 * ni + g] : 0.0);
            }
            const int c = 8 * ni + 2 * t;
            if (r < 15) { Cm[r * 15 + c] = c0; if (c + 1 < 15) Cm[r * 15 + c + 1] = c1; }
        }
    }
    grp_sync<NT>();
}
// C -= A B^T
template <int NT>
__device__ __noinline__ void gemm_sub_abt15(double* Cm, const double* A, const double* B, int tid) {
    constexpr int NW = NT / 32;
    const int lane = tid & 31, wi = tid >> 5, g = lane >> 2, t = lane & 3;
    if (NW == 1) {
        double af[2][4], bf[4][2];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int k = 4 * ks + t;
            af[0][ks] = k < 15 ? -A[g * 15 + k] : 0.0;
            af[1][ks] = (k < 15 && g < 7) ? -A[(8 + g) * 15 + k] : 0.0;
            bf[ks][0] = k < 15 ? B[g * 15 + k] : 0.0;
            bf[ks][1] = (k < 15 && g < 7) ? B[(8 + g) * 15 + k] : 0.0;
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int r = 8 * mi + g, c = 8 * ni + 2 * t;
                double c0 = (r < 15) ? Cm[r * 15 + c] : 0.0, c1 = (r < 15 && c + 1 < 15) ? Cm[r * 15 + c + 1] : 0.0;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) dmma884(c0, c1, af[mi][ks], bf[ks][ni]);
                if (r < 15) { Cm[r * 15 + c] = c0; if (c + 1 < 15) Cm[r * 15 + c + 1] = c1; }
            }
    } else {
#pragma unroll
        for (int tl = 0; tl < 4; ++tl) {
            if ((tl % NW) != wi) continue;
            const int mi = tl >> 1, ni = tl & 1, r = 8 * mi + g, c = 8 * ni + 2 * t;
            double c0 = (r < 15) ? Cm[r * 15 + c] : 0.0, c1 = (r < 15 && c + 1 < 15) ? Cm[r * 15 + c + 1] : 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int k = 4 * ks + t;
                dmma884(c0, c1, (r < 15 && k < 15) ? -A[r * 15 + k] : 0.0, (k < 15 && 8 * ni + g < 15) ? B[(8 * ni + g) * 15 + k] : 0.0);
            }
            if (r < 15) { Cm[r * 15 + c] = c0; if (c + 1 < 15) Cm[r * 15 + c + 1] = c1; }
        }
    }
    grp_sync<NT>();
}
// out[r] (-)= sum_k M[r][k] v[k]   (TRANS: M[k][r])
template <int NT, bool TRANS, bool SUB>
__device__ __forceinline__ void gemv15(double* out, const double* M, const double* v, int tid) {
    if (tid < 15) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 15; ++k) s += (TRANS ? M[k * 15 + tid] : M[tid * 15 + k]) * v[k];
        if (SUB) out[tid] -= s; else out[tid] = s;
    }
    grp_sync<NT>();
}
template <int NT> __device__ __forceinline__ void copy_blk(double* dst, const double* src, int tid) {
    for (int i = tid; i < kBlk; i += NT) dst[i] = src[i];
}
// per-warp shared memory of window_kernel (doubles): 4 blocks (+3 for the arrow topology) + pivot row + one frame's
// laser block + rhs [n][15]
__host__ __device__ inline size_t window_smem_doubles(int n, bool arrow) { return (arrow ? 7 : 4) * kBlk + 16 + 48 + 32 + 16 + (size_t)n * 15 + 16; }
// the fast path: Dp | Cp | Ut | Ha | piv | slb | ssc | rolling rhs
#ifndef LV_WINDOW_BULK
#define LV_WINDOW_BULK 0   // 1: the raw blocks of an elimination step arrive by cp.async.bulk + mbarrier (4 bulk copies issued by one
                           // lane) instead of ten 16-byte cp.async per lane.  Measured on the B200: see DESIGN.md section 3.3.
#endif
constexpr int kWindowFastSmem = 3 * 256 + 192 + 16 + 32 + 32 + 48 + (LV_WINDOW_BULK ? 2 : 0);
__device__ __forceinline__ void bulk_g2s(double* smem_dst, const double* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// =====================================================================================================
// Tile-form 16 x 16 blocks (the batched fast path of window_step: one warp per window, tracking topology).
// A 15 x 15 block lives padded to 16 x 16 as four row-major 8 x 8 tiles [t00 | t01 | t10 | t11] — the layout the DMMA
// fragments want (no guards, 16-byte result stores), the layout the factor kernels write the item record in (so the
// raw blocks of an elimination step arrive as contiguous 16-byte cp.async chunks, no gather), and a layout whose
// element index is shifts and masks of the lane id (round 1's packed 15-stride blocks spent half of the kernel's
// instructions on e / 15, e % 15 and bounds predicates).  Row / column 15 of every block stay zero.
constexpr int kTB = 256;
__device__ __forceinline__ int t_idx(int r, int c) { return ((((r >> 3) << 1) + (c >> 3)) << 6) + ((r & 7) << 3) + (c & 7); }

// Gauss-Jordan inverse of an SPD block in tile form (same algorithm and arithmetic as spd_inverse15)
__device__ __noinline__ bool spd_inverse15_t(double* A, double* piv /* 16 doubles scratch */, int lane) {
    const int r = lane & 15, h = lane >> 4, c0 = h * 8;
    const bool act = r < 15;
    double* rowp = A + ((((r >> 3) << 1) + h) << 6) + ((r & 7) << 3);   // 8 contiguous entries of row r, half h
    double row[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const double2 v = act ? *reinterpret_cast<const double2*>(rowp + j) : make_double2(0.0, 0.0);
        row[j] = v.x; row[j + 1] = v.y;
    }
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        const int kh = k >> 3, kj = k & 7;
        const double p = __shfl_sync(0xffffffffu, row[kj], k + 16 * kh);
        const double f = __shfl_sync(0xffffffffu, row[kj], r + 16 * kh);
        if (!(p > 0.0) || !isfinite(p)) ok = false;
        const double pinv = 1.0 / p;
        if (r == k && act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { row[j] = (c0 + j == k) ? pinv : row[j] * pinv; piv[c0 + j] = row[j]; }
        }
        __syncwarp();
        if (r != k && act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) row[j] = (c0 + j == k) ? -f * pinv : row[j] - f * piv[c0 + j];
        }
        __syncwarp();
    }
    if (act) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(rowp + j) = make_double2(row[j], row[j + 1]);
    }
    __syncwarp();
    return ok;
}
// C = A B (tile form): 16 DMMA, 16 fragment loads, 4 16-byte stores
__device__ __forceinline__ void gemm_ab_t(double* Cm, const double* A, const double* B, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double af[2][4], bf[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            af[m][ks] = A[(((m << 1) + (ks >> 1)) << 6) + (g << 3) + ((ks & 1) << 2) + t];            // A[8m + g][4ks + t]
            bf[ks][m] = B[((((ks >> 1) << 1) + m) << 6) + ((((ks & 1) << 2) + t) << 3) + g];            // B[4ks + t][8m + g]
        }
    }
    __syncwarp();   // C may alias A or B: every fragment is in registers before the first store
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) dmma884(c0, c1, af[mi][ks], bf[ks][ni]);
            *reinterpret_cast<double2*>(Cm + (((mi << 1) + ni) << 6) + (g << 3) + (t << 1)) = make_double2(c0, c1);
        }
    __syncwarp();
}
// C -= A B^T (tile form)
__device__ __forceinline__ void gemm_sub_abt_t(double* Cm, const double* A, const double* B, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double af[2][4], bf[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            af[m][ks] = -A[(((m << 1) + (ks >> 1)) << 6) + (g << 3) + ((ks & 1) << 2) + t];           // -A[8m + g][4ks + t]
            bf[ks][m] = B[(((m << 1) + (ks >> 1)) << 6) + (g << 3) + ((ks & 1) << 2) + t];            // B[8m + g][4ks + t] = B^T[4ks + t][8m + g]
        }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            double2* cp = reinterpret_cast<double2*>(Cm + (((mi << 1) + ni) << 6) + (g << 3) + (t << 1));
            double2 c = *cp;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) dmma884(c.x, c.y, af[mi][ks], bf[ks][ni]);
            *cp = c;
        }
    __syncwarp();
}
// out[r] (-)= sum_k M(r, k) v[k]  (TRANS: M(k, r)), M in tile form, r, k < 15
template <bool TRANS, bool SUB>
__device__ __forceinline__ void gemv_t(double* out, const double* M, const double* v, int lane) {
    if (lane < 15) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 15; ++k) s += M[TRANS ? t_idx(k, lane) : t_idx(lane, k)] * v[k];
        if (SUB) out[lane] -= s; else out[lane] = s;
    }
    __syncwarp();
}

// C -= A^T B (tile form)
__device__ __forceinline__ void gemm_sub_atb_t(double* Cm, const double* A, const double* B, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double af[2][4], bf[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            af[m][ks] = -A[((((ks >> 1) << 1) + m) << 6) + ((((ks & 1) << 2) + t) << 3) + g];           // -A[4ks + t][8m + g] = -A^T[8m + g][4ks + t]
            bf[ks][m] = B[((((ks >> 1) << 1) + m) << 6) + ((((ks & 1) << 2) + t) << 3) + g];            // B[4ks + t][8m + g]
        }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            double2* cp = reinterpret_cast<double2*>(Cm + (((mi << 1) + ni) << 6) + (g << 3) + (t << 1));
            double2 c = *cp;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) dmma884(c.x, c.y, af[mi][ks], bf[ks][ni]);
            *cp = c;
        }
    __syncwarp();
}
// C = -A B (tile form); C may alias A or B
__device__ __forceinline__ void gemm_neg_ab_t(double* Cm, const double* A, const double* B, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double af[2][4], bf[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            af[m][ks] = -A[(((m << 1) + (ks >> 1)) << 6) + (g << 3) + ((ks & 1) << 2) + t];
            bf[ks][m] = B[((((ks >> 1) << 1) + m) << 6) + ((((ks & 1) << 2) + t) << 3) + g];
        }
    }
    __syncwarp();
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) dmma884(c0, c1, af[mi][ks], bf[ks][ni]);
            *reinterpret_cast<double2*>(Cm + (((mi << 1) + ni) << 6) + (g << 3) + (t << 1)) = make_double2(c0, c1);
        }
    __syncwarp();
}
// shared memory of the cyclic-reduction shape (doubles) behind the generic per-window area: D [n] | U [1..n-1] | rhs [n][16]
// | pivot rows [16][16] | flag [2] | extra scratch blocks
__host__ __device__ inline size_t window_cr_doubles(int n, int extra_scratch) {
    return (size_t)n * kTB + (size_t)(n > 1 ? n - 1 : 0) * kTB + (size_t)n * 16 + 256 + 2 + (size_t)extra_scratch * kTB;
}

// item offsets of the three raw blocks of an elimination step, per packed 15 x 15 entry e = 15 r + c:
// H(i-1, i)(r, c) | hbb(r, c) << 10 | haa(r, c) << 20 — built once per CTA (window_offsets_init), read by every frame
__device__ __forceinline__ void window_offsets_init(uint32_t* tab, int tid, int nthreads) {
    for (int e = tid; e < kBlk; e += nthreads) {
        const int r = e / 15, c = e - r * 15;
        tab[e] = (uint32_t)item_hab(r, c) | ((uint32_t)item_hbb(r, c) << 10) | ((uint32_t)item_haa(r, c) << 20);
    }
}

template <bool ARROW, int NT>
__device__ void window_step(const WindowArgs& a, int w, int lane, double* ws, const uint32_t* otab) {
    constexpr int NPAD = ARROW ? kPadFree : kPadTrack;
    constexpr int ICOST = ARROW ? 44 : 20;
    constexpr int IGJ = ARROW ? 36 : 15;
    const int n = a.n_frames;
    LMState st = a.state[w];
    if (st.status != 0) return;
    const LMOptions& opt = a.opt;
    const int mode = a.mode;

    double* Dm = ws;                  // pivot block / its inverse
    double* Cy = ws + kBlk;           // next diagonal block being assembled
    double* Um = ws + 2 * kBlk;       // coupling H(i-1, i)
    double* Tm = ws + 3 * kBlk;       // T = U Dinv
    double* Wm = ws + 4 * kBlk;       // arrow block H(0, i)          (arrow topology only)
    double* Tp = ws + 5 * kBlk;       // T' = W Dinv
    double* D0 = ws + 6 * kBlk;       // accumulated updates of D_0
    double* piv = ws + (ARROW ? 7 : 4) * kBlk;  // 16
    double* slb = piv + 16;           // laser block of the frame being assembled [NPAD]
    double* ssc = slb + 48;           // Jacobi scaling of frames i-1 and i [30]
    double* red = ssc + 32;           // cross-warp reduction scratch [16]
    // the fast path (one warp per window, tracking topology, solver program) keeps the right-hand side in global memory
    // and only a rolling two-frame window of it in shared memory: 8.6 KB per warp instead of 12.9
    constexpr bool FASTC = !ARROW && NT == 32;
    const bool fast = FASTC && mode == 0 && !a.dense_H;
    double* sb = fast ? a.vec + ((size_t)w * 3 + 2) * n * 15 : red + 16;   // rhs / solution [n][15]

    double* x = a.x + (size_t)w * n * 15;
    double* xc = a.xc + (size_t)w * n * 15;
    const uint8_t* cm = a.const_mask + (size_t)w * n;
    const uint8_t* fa = a.frame_active + (size_t)w * n;
    const size_t F = (size_t)a.n_windows * n;
    auto is_const = [&](int f, int c) -> bool { return mode == 1 ? false : col_const(cm[f], c); };
    const double lsq = a.C.laser_sqrt_info * a.C.laser_sqrt_info;
    const int cand = 1 - st.cur;
    double* lb_c = a.laser_blocks + ((size_t)cand * F + (size_t)w * n) * NPAD;
    const double* it_c = a.items + ((size_t)cand * F + (size_t)w * n) * kItem;

    // ---- (1) candidate laser blocks: sum the tiles in a fixed order; candidate cost
    double csum = 0.0;
    if (a.tiles == 1) {
        // one tile per frame (the batched shape): 4 independent loads per lane in flight
        const double* pw = a.partial + (size_t)w * n * NPAD;
        for (int i0 = lane; i0 < n * NPAD; i0 += 4 * NT) {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = i0 + NT * q;
                v[q] = (idx < n * NPAD && fa[idx / NPAD]) ? pw[idx] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = i0 + NT * q;
                if (idx < n * NPAD) {
                    const double s = 0.0 + v[q];
                    lb_c[idx] = s;
                    if (idx % NPAD == ICOST) csum += lsq * s;
                }
            }
        }
    } else {
        for (int idx = lane; idx < n * NPAD; idx += NT) {
            const int f = idx / NPAD, k = idx - f * NPAD;
            double s = 0.0;
            if (fa[f]) {
                const double* p = a.partial + ((size_t)(w * n + f) * a.tiles) * NPAD + k;
                for (int t = 0; t < a.tiles; ++t) s += p[(size_t)t * NPAD];
            }
            lb_c[idx] = s;
            if (k == ICOST) csum += lsq * s;
        }
    }
    for (int f = lane; f < n; f += NT) csum += it_c[(size_t)f * kItem + kItemCost];
    double cand_cost = 0.5 * grp_sum<NT>(csum, red, lane);
    grp_sync<NT>();

    // ---- decide on the candidate
    bool accepted = false;
    if (!st.started) {
        st.started = 1;   // iteration 0: the "candidate" is the initial point
        st.cost = cand_cost;
        st.initial_cost = cand_cost;
        accepted = true;
        double s = 0.0;
        for (int i0 = lane; i0 < n * 15; i0 += 4 * NT) {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = (i0 + NT * q < n * 15) ? xc[i0 + NT * q] : 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + NT * q;
                if (i < n * 15) {
                    x[i] = v[q];
                    if (!is_const(i / 15, i % 15)) s += v[q] * v[q];
                }
            }
        }
        st.x_norm = sqrt(grp_sum<NT>(s, red, lane));
        st.last_success = 1;
    } else {
        if (!isfinite(cand_cost)) cand_cost = 1.7976931348623157e308;
        double s = 0.0;
        for (int i0 = lane; i0 < n * 15; i0 += 4 * NT) {
            double va[4], vb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool in = i0 + NT * q < n * 15;
                va[q] = in ? x[i0 + NT * q] : 0.0;
                vb[q] = in ? xc[i0 + NT * q] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + NT * q;
                if (i < n * 15 && !is_const(i / 15, i % 15)) { const double d = va[q] - vb[q]; s += d * d; }
            }
        }
        const double step_norm = sqrt(grp_sum<NT>(s, red, lane));
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
                    for (int i0 = lane; i0 < n * 15; i0 += 4 * NT) {
                        double v[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[q] = (i0 + NT * q < n * 15) ? xc[i0 + NT * q] : 0.0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int i = i0 + NT * q;
                            if (i < n * 15) {
                                x[i] = v[q];
                                if (!is_const(i / 15, i % 15)) s2 += v[q] * v[q];
                            }
                        }
                    }
                    st.x_norm = sqrt(grp_sum<NT>(s2, red, lane));
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
    if (accepted) st.cur = cand;
    grp_sync<NT>();
    if (mode == 0 && st.iteration >= opt.max_iters) {
        st.status = 1; st.termination = 0;
        if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
        return;
    }
    // the linearisation of the accepted point
    const double* lb = a.laser_blocks + ((size_t)st.cur * F + (size_t)w * n) * NPAD;
    const double* itm = a.items + ((size_t)st.cur * F + (size_t)w * n) * kItem;

    // entry (r, c) of the laser 6x6 own-pose block of frame f, unscaled
    // (blk: the frame's laser block, lb + f * NPAD or its shared-memory copy)
    auto laser_own_at = [&](const double* blk, int f, int r, int c) -> double {
        if (r == 2 || c == 2 || r > 5 || c > 5 || !fa[f]) return 0.0;
        if (is_const(f, r) || is_const(f, c)) return 0.0;
        int ri = r < 2 ? r : r - 1, ci = c < 2 ? c : c - 1;
        if (ri > ci) { const int t = ri; ri = ci; ci = t; }
        // upper 5x5 in the order aa(3) | a x bj(6) | bj bj(6)
        const int src = (ri < 2 && ci < 2) ? (ri + ci) : (ri < 2 ? 3 + ri * 3 + (ci - 2) : 9 + (ri == 2 ? ci - 2 : (ri == 3 ? ci : 5)));
        return lsq * blk[src];
    };
    auto laser_own = [&](int f, int r, int c) -> double { return laser_own_at(lb + f * NPAD, f, r, c); };
    // the same with the frame's const mask and laser flag already in registers
    auto laser_own_reg = [&](const double* blk, uint8_t mask, bool active, int r, int c) -> double {
        if (r == 2 || c == 2 || r > 5 || c > 5 || !active) return 0.0;
        if (col_const(mask, r) || col_const(mask, c)) return 0.0;
        int ri = r < 2 ? r : r - 1, ci = c < 2 ? c : c - 1;
        if (ri > ci) { const int t = ri; ri = ci; ci = t; }
        const int src = (ri < 2 && ci < 2) ? (ri + ci) : (ri < 2 ? 3 + ri * 3 + (ci - 2) : 9 + (ri == 2 ? ci - 2 : (ri == 3 ? ci : 5)));
        return lsq * blk[src];
    };
    auto has_cross = [&](int j) -> bool {
        return ARROW && mode == 0 && j >= 1 && fa[j] && a.ref_frame && a.ref_frame[(size_t)w * n + j] == 0;
    };
    // reference-frame side (arrow): 6x6 block of frame 0 summed over all frames that hang under it
    auto laser_ref_own = [&](int r, int c) -> double {
        if (!ARROW || mode != 0) return 0.0;
        if (r == 2 || c == 2 || r > 5 || c > 5) return 0.0;
        if (is_const(0, r) || is_const(0, c)) return 0.0;
        int ri = r < 2 ? r : r - 1, ci = c < 2 ? c : c - 1;
        if (ri > ci) { const int t = ri; ri = ci; ci = t; }
        double s = 0.0;
        for (int j = 1; j < n; ++j) {
            if (!has_cross(j)) continue;
            const double* cb = lb + j * NPAD;
            double v;
            if (ri < 2 && ci < 2) v = cb[ri + ci];
            else if (ri < 2) v = -cb[15 + ri * 3 + (ci - 2)];
            else v = cb[21 + (ri == 2 ? ci - 2 : (ri == 3 ? ci : 5))];
            s += v;
        }
        return lsq * s;
    };
    // laser cross block between frame j and its reference frame 0 as H(0, j) (rows 0, cols j), 6x6 pose part
    auto cross_entry = [&](int j, int r, int c) -> double {
        if (r == 2 || c == 2) return 0.0;
        const double* cb = lb + j * NPAD;
        const int ri = r < 2 ? r : r - 1, ci = c < 2 ? c : c - 1;
        double v;
        if (ri < 2 && ci < 2) v = -cb[ri + ci];                      // -(a a^T)
        else if (ri < 2) v = -cb[3 + ri * 3 + (ci - 2)];             // rows p_0 (-a), cols theta_j (bj)
        else if (ci < 2) v = cb[15 + ci * 3 + (ri - 2)];             // rows theta_0 (bi), cols p_j (a)
        else v = cb[27 + (ci - 2) * 3 + (ri - 2)];                   // bj x bi stored [bj][bi]
        return lsq * v;
    };
    // gradient entry c of frame f
    auto grad = [&](int f, int c) -> double {
        if (is_const(f, c)) return 0.0;
        double g = itm[(size_t)f * kItem + item_gb(c)];
        if (f + 1 < n) g += itm[(size_t)(f + 1) * kItem + item_ga(c)];
        if (c < 6 && c != 2) {
            const int k = c < 2 ? c : c - 1;
            if (fa[f]) g += lsq * lb[f * NPAD + IGJ + k];
            if (ARROW && mode == 0 && f == 0) {
                for (int j = 1; j < n; ++j) {
                    if (!has_cross(j)) continue;
                    const double* cb = lb + j * NPAD;
                    g += lsq * (k < 2 ? -cb[36 + k] : cb[41 + k - 2]);
                }
            }
        }
        return g;
    };
    // unscaled diagonal block of frame i into dst (shared memory)
    auto assemble_D = [&](int i, double* dst) {
        for (int e = lane; e < kBlk; e += NT) {
            const int r = e / 15, c = e - r * 15;
            double v = itm[(size_t)i * kItem + item_hbb(r, c)];
            if (i + 1 < n) v += itm[(size_t)(i + 1) * kItem + item_haa(r, c)];
            v += laser_own(i, r, c);
            if (i == 0) v += laser_ref_own(r, c);
            dst[e] = v;
        }
        grp_sync<NT>();
    };
    auto diag_H = [&](int f, int c) -> double {
        double v = itm[(size_t)f * kItem + item_hbb(c, c)];
        if (f + 1 < n) v += itm[(size_t)(f + 1) * kItem + item_haa(c, c)];
        v += laser_own(f, c, c);
        if (f == 0) v += laser_ref_own(c, c);
        return v;
    };

    // ---- dense outputs for lvio2d_linearize
    if (a.dense_H) {
        const int dim = 15 * n;
        double* H = a.dense_H + (size_t)w * dim * dim;
        for (size_t e = lane; e < (size_t)dim * dim; e += NT) H[e] = 0.0;
        grp_sync<NT>();
        for (int i = 0; i < n; ++i) {
            assemble_D(i, Dm);
            for (int e = lane; e < kBlk; e += NT) H[(size_t)(15 * i + e / 15) * dim + 15 * i + e % 15] = Dm[e];
            if (i >= 1)
                for (int e = lane; e < kBlk; e += NT) {
                    const double v = itm[(size_t)i * kItem + item_hab(e / 15, e % 15)];  // H(i-1, i)
                    H[(size_t)(15 * (i - 1) + e / 15) * dim + 15 * i + e % 15] += v;
                    H[(size_t)(15 * i + e % 15) * dim + 15 * (i - 1) + e / 15] += v;
                }
            grp_sync<NT>();
            if (has_cross(i))
                for (int e = lane; e < 36; e += NT) {
                    const int r = e / 6, c = e % 6;
                    if (is_const(0, r) || is_const(i, c)) continue;
                    const double v = cross_entry(i, r, c);
                    H[(size_t)r * dim + 15 * i + c] += v;
                    H[(size_t)(15 * i + c) * dim + r] += v;
                }
            grp_sync<NT>();
        }
        for (int i = lane; i < dim; i += NT) a.dense_g[(size_t)w * dim + i] = grad(i / 15, i % 15);
        if (lane == 0) a.dense_cost[w] = st.cost;
        st.status = 1;
        if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
        return;
    }

    // ---- marginalisation program: forward elimination of frames 0..n-2 (solver.cpp:4-40)
    if (mode == 1) {
        for (int i = lane; i < n * 15; i += NT) sb[i] = -grad(i / 15, i % 15);   // g = -J^T R
        grp_sync<NT>();
        assemble_D(0, Dm);
        for (int i = 0; i + 1 < n; ++i) {
            assemble_D(i + 1, Cy);
            // Um = H(i+1, i) = H(i, i+1)^T
            for (int e = lane; e < kBlk; e += NT) Um[e] = itm[(size_t)(i + 1) * kItem + item_hab(e % 15, e / 15)];
            grp_sync<NT>();
            if (!grp_inverse15<NT>(Dm, piv, red, lane)) st.termination = 5;
            gemm_ab15<NT>(Tm, Um, Dm, lane);                 // T = H(i+1,i) Hii^-1
            gemm_sub_abt15<NT>(Cy, Tm, Um, lane);            // H(i+1,i+1) -= T H(i+1,i)^T
            gemv15<NT, false, true>(sb + 15 * (i + 1), Tm, sb + 15 * i, lane);
            copy_blk<NT>(Dm, Cy, lane);
            grp_sync<NT>();
        }
        for (int e = lane; e < kBlk; e += NT) a.marg_H[(size_t)w * kBlk + e] = Dm[e];
        if (lane < 15) a.marg_g[(size_t)w * 15 + lane] = sb[15 * (n - 1) + lane];
        st.status = 1;
        if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
        return;
    }

    // ---- gradient tolerance: |x - Plus(x, -g)|_inf (only after a successful step); gradient kept for the step
    double* gvec = a.vec + (size_t)w * 3 * n * 15;
    double* dvec = gvec + n * 15;   // unscaled diagonal of H at the accepted point
    {
        double mx = 0.0;
        for (int i0 = lane; i0 < n * 15; i0 += 4 * NT) {
            double g[4], dg[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + NT * q;
                const bool in = i < n * 15;
                g[q] = in ? grad(i / 15, i % 15) : 0.0;
                dg[q] = in ? diag_H(i / 15, i % 15) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + NT * q;
                if (i < n * 15) { gvec[i] = g[q]; dvec[i] = dg[q]; }
            }
        }
        grp_sync<NT>();
        if (st.last_success) {
            for (int f = lane; f < n; f += NT) {
#pragma unroll
                for (int c = 0; c < 15; ++c) {
                    if (is_const(f, c) || (c >= 3 && c < 6)) continue;
                    mx = fmax(mx, fabs(gvec[f * 15 + c]));
                }
                if (!is_const(f, 3)) {
                    double neg[3] = {-gvec[f * 15 + 3], -gvec[f * 15 + 4], -gvec[f * 15 + 5]}, outv[3];
                    so3_plus(x + 15 * f + 3, neg, outv);
#pragma unroll
                    for (int c = 0; c < 3; ++c) mx = fmax(mx, fabs(x[15 * f + 3 + c] - outv[c]));
                }
            }
            mx = grp_max<NT>(mx, red, lane);
            if (mx <= opt.gradient_tolerance) {
                st.status = 1; st.termination = 3;
                if (lane == 0) { a.state[w] = st; a.win_status[w] = 1; }
                return;
            }
        }
    }

    // ---- Jacobi scaling (fixed at iteration 0)
    double* scw = a.scale + (size_t)w * n * 15;
    if (st.iteration == 0) {
        for (int i = lane; i < n * 15; i += NT)
            scw[i] = is_const(i / 15, i % 15) ? 1.0 : 1.0 / (1.0 + sqrt(dvec[i]));
        grp_sync<NT>();
    }

    // ---- (3) trust-region step; invalid steps shrink the radius and retry without a new evaluation
    double* facw = a.fac + (size_t)w * n * 3 * kBlk;
    bool have_step = false;
    for (;;) {
        if (st.radius < opt.min_radius) { st.status = 1; st.termination = 4; break; }
        ++st.iteration;
        bool ok = true;
        for (int i = lane; i < n * 15; i += NT) sb[i] = is_const(i / 15, i % 15) ? 0.0 : gvec[i] * scw[i];
        grp_sync<NT>();
        const double inv_radius = 1.0 / st.radius;
        auto scale_damp = [&](int i, double* blkp) {  // A = S H S + diag(clamp(diag(S H S)) / radius); const entries -> identity
            for (int e = lane; e < kBlk; e += NT) {
                const int r = e / 15, c = e - r * 15;
                double v = blkp[e] * scw[i * 15 + r] * scw[i * 15 + c];
                if (r == c) v = is_const(i, r) ? 1.0 : v + fmin(fmax(v, opt.min_lm_diagonal), opt.max_lm_diagonal) * inv_radius;
                blkp[e] = v;
            }
            grp_sync<NT>();
        };
        // step = -y, model cost change; sets ok
        auto step_and_model = [&]() {
            // step = -y ; model_cost_change = -1/2 step.gs + 1/2 sum lm_diag step^2
            double sg = 0.0, lq = 0.0;
            bool finite = true;
            for (int i = lane; i < n * 15; i += NT) {
                if (is_const(i / 15, i % 15)) { sb[i] = 0.0; continue; }
                const double stp = -sb[i];
                sb[i] = stp;
                if (!isfinite(stp)) finite = false;
                const double sc = scw[i];
                const double hs = dvec[i] * sc * sc;
                sg += stp * gvec[i] * sc;
                lq += fmin(fmax(hs, opt.min_lm_diagonal), opt.max_lm_diagonal) * inv_radius * stp * stp;
            }
            sg = grp_sum<NT>(sg, red, lane);
            lq = grp_sum<NT>(lq, red, lane);
            finite = grp_all<NT>(finite, red, lane);
            st.model_cost_change = -0.5 * sg + 0.5 * lq;
            ok = finite && st.model_cost_change > 0.0;
                };
        if constexpr (!ARROW && NT == 512) {
            // ---- cyclic reduction (small batches: one CTA of 16 warps per window).  The block-tridiagonal system is
            // eliminated level by level — level l removes the frames i = s (mod 2s), s = 2^l, all of them concurrently,
            // one warp each — so the sequential depth is ceil(log2 n) + 1 pivot inversions instead of n.  For an
            // eliminated frame i with live neighbours l = i - s, r = i + s:
            //     Dinv = D_i^-1, c_i = Dinv b_i, TL = U_i Dinv, Y = Dinv U_r          (U_k = coupling H(left of k, k))
            //     D_l -= TL U_i^T, b_l -= U_i c_i          |  D_r -= U_r^T Y, b_r -= U_r^T c_i,  U_r <- -TL U_r
            // and on the way back x_i = c_i - TL^T x_l - Y x_r.  TL and Y stay in the dead slots of U_i and D_i.
            // Same LM system as the other shapes; the elimination ORDER differs (nested dissection instead of n-1 .. 0),
            // so results agree with them to rounding, not bit for bit.
            const int wi = lane >> 5, wl = lane & 31;
            constexpr int NW = NT / 32;
            double* crb = ws + ((window_smem_doubles(n, false) + 1) & ~(size_t)1);
            double* Db = crb;                                   // [n][256]
            double* Ub = Db + (size_t)(n - 1) * kTB;            // U_f at Ub + f * 256, f >= 1 (no slot for frame 0)
            double* rb = Db + (size_t)(2 * n - 1) * kTB;        // [n][16]
            double* pv = rb + (size_t)n * 16;                   // [16][16]
            double* flg = pv + 256;                             // [2]
            double* Sx = flg + 2;                               // extra scratch blocks
            const int nscr = min(NW, 3 + a.cr_scratch);         // warps that can hold a TL block at the same time
            double* Sw = wi < 3 ? ws + wi * kTB : Sx + (size_t)(wi - 3) * kTB;   // the first three live in the generic area
            const int lr = wl >> 3, lc = wl & 7;
            if (lane == 0) flg[0] = 0.0;
            // level-0 blocks: D_f = S (hbb_f + haa_{f+1} + laser_f) S + damping, U_f = S_{f-1} H(f-1, f) S_f, b_f
            for (int f = wi; f < n; f += NW) {
                const double* it_f = itm + (size_t)f * kItem;
                const bool nxt = f + 1 < n;
                const double* it_n = itm + (size_t)(nxt ? f + 1 : f) * kItem;
                const uint8_t cm_f = cm[f];
                const bool fa_f = fa[f] != 0;
                const double* blk = lb + f * NPAD;
                double hb[8], ha[8], hu[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int tile = q >> 1;
                    const int win = ((q & 1) << 5) + wl, wtr = (lc << 3) + ((q & 1) << 2) + lr;
                    const int off = tile == 0 ? win : (tile == 1 ? 64 + win : (tile == 2 ? 64 + wtr : 128 + win));
                    hb[q] = it_f[kItemHbb + off];
                    ha[q] = nxt ? it_n[off] : 0.0;
                    hu[q] = f >= 1 ? it_f[kItemHab + (q << 5) + wl] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int tile = q >> 1;
                    const int r = ((q >> 2) << 3) + ((q & 1) << 2) + lr, c = ((tile & 1) << 3) + lc;
                    const bool in = r < 15 && c < 15;
                    const double sr = scw[f * 15 + (r < 15 ? r : 0)], sc = scw[f * 15 + (c < 15 ? c : 0)];
                    double d = hb[q] + ha[q];
                    if (tile == 0) d += laser_own_reg(blk, cm_f, fa_f, r, c);
                    d = d * sr * sc;
                    if ((tile == 0 || tile == 3) && r == c)
                        d = col_const(cm_f, r) ? 1.0 : d + fmin(fmax(d, opt.min_lm_diagonal), opt.max_lm_diagonal) * inv_radius;
                    Db[(size_t)f * kTB + (q << 5) + wl] = in ? d : 0.0;
                    if (f >= 1) Ub[(size_t)f * kTB + (q << 5) + wl] = in ? hu[q] * scw[(f - 1) * 15 + r] * sc : 0.0;
                }
                if (wl < 16) rb[f * 16 + wl] = wl < 15 ? sb[f * 15 + wl] : 0.0;
            }
            __syncthreads();
            int s = 1;
            for (; s < n; s <<= 1) {
                const int cnt = (n - s + 2 * s - 1) / (2 * s);           // frames s, 3s, 5s, ... < n
                for (int base = 0; base < cnt; base += nscr) {
                    const int k = base + wi;
                    const bool act = wi < nscr && k < cnt;
                    const int i = s + 2 * s * k, l = i - s, r = i + s;
                    const bool has_r = act && r < n;
                    if (act) {
                        double* Di = Db + (size_t)i * kTB;
                        const double* Ui = Ub + (size_t)i * kTB;
                        if (!spd_inverse15_t(Di, pv + wi * 16, wl) && wl == 0) flg[0] = 1.0;
                        gemv_t<false, false>(pv + wi * 16, Di, rb + i * 16, wl);                 // c_i = Dinv b_i
                        if (wl < 15) rb[i * 16 + wl] = pv[wi * 16 + wl];
                        gemm_ab_t(Sw, Ui, Di, wl);                                               // TL = U_i Dinv
                        if (has_r) gemm_ab_t(Di, Di, Ub + (size_t)r * kTB, wl);                  // Y = Dinv U_r, over Dinv
                        gemm_sub_abt_t(Db + (size_t)l * kTB, Sw, Ui, wl);                        // D_l -= TL U_i^T
                        gemv_t<false, true>(rb + l * 16, Ui, rb + i * 16, wl);                   // b_l -= U_i c_i
                    }
                    __syncthreads();
                    if (has_r) {
                        double* Ur = Ub + (size_t)r * kTB;
                        gemm_sub_atb_t(Db + (size_t)r * kTB, Ur, Db + (size_t)i * kTB, wl);      // D_r -= U_r^T Y
                        gemv_t<true, true>(rb + r * 16, Ur, rb + i * 16, wl);                    // b_r -= U_r^T c_i
                        gemm_neg_ab_t(Ur, Sw, Ur, wl);                                           // U_r <- H(l, r) = -TL U_r
                    }
                    if (act) {
                        double* Ui = Ub + (size_t)i * kTB;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<double2*>(Ui + 2 * (wl + 32 * q)) = *reinterpret_cast<const double2*>(Sw + 2 * (wl + 32 * q));
                    }
                    __syncthreads();
                }
            }
            if (wi == 0) {
                if (!spd_inverse15_t(Db, pv, wl) && wl == 0) flg[0] = 1.0;
                gemv_t<false, false>(pv, Db, rb, wl);                                            // x_0 = D_0^-1 b_0
                if (wl < 15) rb[wl] = pv[wl];
            }
            __syncthreads();
            ok = flg[0] == 0.0;
            if (ok) {
                for (s >>= 1; s >= 1; s >>= 1) {
                    const int cnt = (n - s + 2 * s - 1) / (2 * s);
                    for (int k = wi; k < cnt; k += NW) {
                        const int i = s + 2 * s * k, l = i - s, r = i + s;
                        gemv_t<true, true>(rb + i * 16, Ub + (size_t)i * kTB, rb + l * 16, wl);          // x_i = c_i - TL^T x_l
                        if (r < n) gemv_t<false, true>(rb + i * 16, Db + (size_t)i * kTB, rb + r * 16, wl);   //     - Y x_r
                    }
                    __syncthreads();
                }
                for (int e = lane; e < n * 15; e += NT) sb[e] = rb[(e / 15) * 16 + e % 15];
                __syncthreads();
                step_and_model();
            }
        } else if constexpr (FASTC) {
            // shared memory of the fast path (doubles): Dp 256 | Cp 256 | Ut 256 | Ha 192 | piv 16 | slb 32 | ssc 32 | rhs 3 x 16
            double* Dp = ws;
            double* Cp = ws + kTB;
            double* Ut = ws + 2 * kTB;
            double* Ha = ws + 3 * kTB;          // raw haa tiles of item i (slots t00 t01 t11), added into D_{i-1}
            double* pivf = Ha + 192;
            double* slbf = pivf + 16;
            double* sscf = slbf + 32;
            double* bcur = sscf + 32;           // b_i, later c_i / y_i  (rolling: the full right-hand side lives in global memory)
            double* bprv = bcur + 16;           // b_{i-1}
            double* facf = a.fac + (size_t)w * n * kTB;
            // element (r, c) this lane owns in pass q of a block: tile q >> 1, row 4 (q & 1) + (lane >> 3), column lane & 7
            const int lr = lane >> 3, lc = lane & 7;
#if LV_WINDOW_BULK
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(ws + kWindowFastSmem - 2);
            uint32_t bar_phase = 0;
            if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
#endif
            {   // D_{n-1}, scaled and damped
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int r = ((q >> 2) << 3) + ((q & 1) << 2) + lr, c = (((q >> 1) & 1) << 3) + lc;
                    double v = 0.0;
                    if (r < 15 && c < 15) {
                        v = itm[(size_t)(n - 1) * kItem + item_hbb(r, c)] + laser_own(n - 1, r, c);
                        v = v * scw[(n - 1) * 15 + r] * scw[(n - 1) * 15 + c];
                        if (r == c) v = is_const(n - 1, r) ? 1.0 : v + fmin(fmax(v, opt.min_lm_diagonal), opt.max_lm_diagonal) * inv_radius;
                    }
                    Dp[(q << 5) + lane] = v;
                }
                if (lane < 16) bcur[lane] = lane < 15 ? sb[15 * (n - 1) + lane] : 0.0;
                __syncwarp();
            }
            for (int i = n - 1; i >= 1; --i) {
                // raw blocks of this elimination step as contiguous 16-byte chunks, fetched while the pivot is inverted:
                // Ut <- hab of item i (4 tiles), Ha / Cp slots t00 t01 t11 <- haa of item i / hbb of item i-1
                {
                    const double* it_i = itm + (size_t)i * kItem;
                    const double* it_p = itm + (size_t)(i - 1) * kItem + kItemHbb;
#if LV_WINDOW_BULK
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2048u + 1536u + 1024u + 512u) : "memory");
                        bulk_g2s(Ut, it_i + kItemHab, 2048u, bar);
                        bulk_g2s(Ha, it_i, 1536u, bar);
                        bulk_g2s(Cp, it_p, 1024u, bar);
                        bulk_g2s(Cp + 192, it_p + 128, 512u, bar);
                    }
#else
#pragma unroll
                    for (int k = 0; k < 4; ++k) cp_async16(Ut + 2 * (lane + 32 * k), it_i + kItemHab + 2 * (lane + 32 * k));
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        cp_async16(Ha + (k << 6) + 2 * lane, it_i + (k << 6) + 2 * lane);
                        cp_async16(Cp + ((k == 2 ? 3 : k) << 6) + 2 * lane, it_p + (k << 6) + 2 * lane);
                    }
#endif
                    if (lane < NPAD) cp_async8(slbf + lane, lb + (i - 1) * NPAD + lane);
                    if (lane < 30) cp_async8(sscf + lane, scw + (i - 1) * 15 + lane);
                    if (lane < 15) cp_async8(bprv + lane, sb + 15 * (i - 1) + lane);
                }
                const uint8_t cm_p = cm[i - 1];
                const bool fa_p = fa[i - 1] != 0;
                __syncwarp();
                const bool inv_ok = spd_inverse15_t(Dp, pivf, lane);
                cp_async_wait_all();
#if LV_WINDOW_BULK
                {
                    uint32_t done = 0;
                    while (!done)
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(done) : "r"(bar), "r"(bar_phase) : "memory");
                    bar_phase ^= 1u;
                }
#endif
                __syncwarp();
                if (!inv_ok) { ok = false; break; }
                // U = S_{i-1} H(i-1, i) S_i;  D_{i-1} = S (hbb + haa + laser) S + damping.  Read everything, then write in place
                // (the lower tile t10 of D is the transpose of t01).
                double uv[8], dv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int tile = q >> 1;
                    const int r = ((q >> 2) << 3) + ((q & 1) << 2) + lr, c = ((tile & 1) << 3) + lc;
                    const int e = (q << 5) + lane;
                    const int win = ((q & 1) << 5) + lane;                                   // offset inside the tile
                    const int wtr = (lc << 3) + ((q & 1) << 2) + lr;                         // ... of the transposed entry
                    const bool in = r < 15 && c < 15;
                    const double sr = sscf[r < 15 ? r : 0], sci = sscf[15 + (c < 15 ? c : 0)], scp = sscf[c < 15 ? c : 0];
                    uv[q] = in ? Ut[e] * sr * sci : 0.0;
                    double d = tile == 2 ? Cp[64 + wtr] : Cp[e];
                    d += tile == 0 ? Ha[win] : (tile == 3 ? Ha[128 + win] : (tile == 1 ? Ha[64 + win] : Ha[64 + wtr]));
                    if (tile == 0) d += laser_own_reg(slbf, cm_p, fa_p, r, c);
                    d = d * sr * scp;
                    if ((tile == 0 || tile == 3) && r == c)
                        d = col_const(cm_p, r) ? 1.0 : d + fmin(fmax(d, opt.min_lm_diagonal), opt.max_lm_diagonal) * inv_radius;
                    dv[q] = in ? d : 0.0;
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; ++q) { Ut[(q << 5) + lane] = uv[q]; Cp[(q << 5) + lane] = dv[q]; }
                __syncwarp();
                gemv_t<false, false>(pivf, Dp, bcur, lane);                         // c_i = Dinv b_i
                if (lane < 15) { const double ci = pivf[lane]; bcur[lane] = ci; sb[15 * i + lane] = ci; }
                gemm_ab_t(Dp, Ut, Dp, lane);                                        // T = U Dinv, over Dinv (loads, barrier, stores)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    *reinterpret_cast<double2*>(facf + (size_t)i * kTB + 2 * (lane + 32 * k)) = *reinterpret_cast<const double2*>(Dp + 2 * (lane + 32 * k));
                gemv_t<false, true>(bprv, Ut, bcur, lane);                          // b_{i-1} -= U c_i
                gemm_sub_abt_t(Cp, Dp, Ut, lane);                                   // D_{i-1} -= T U^T
                { double* tsw = Dp; Dp = Cp; Cp = tsw; }
                { double* tsw = bcur; bcur = bprv; bprv = tsw; }
            }
            if (ok) ok = spd_inverse15_t(Dp, pivf, lane);
            if (ok) {
                gemv_t<false, false>(pivf, Dp, bcur, lane);   // y_0 = D0inv b_0
                if (lane < 15) { const double y = pivf[lane]; bcur[lane] = y; sb[lane] = y; }
                __syncwarp();
                // T_i and c_i come back from global memory one frame ahead of their use (Ut / Cp alternate)
                auto fetch_T = [&](int i, double* dstT, double* dstc) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) cp_async16(dstT + 2 * (lane + 32 * k), facf + (size_t)i * kTB + 2 * (lane + 32 * k));
                    if (lane < 15) cp_async8(dstc + lane, sb + 15 * i + lane);
                };
                double* ycur = bcur;                          // y_{i-1}
                double* cA = bprv;                            // c_i buffers alternate with the T buffers
                double* cB = pivf;
                if (n > 1) fetch_T(1, Ut, cA);
                for (int i = 1; i < n; ++i) {
                    double* Tc = (i & 1) ? Ut : Cp;
                    double* cc = (i & 1) ? cA : cB;
                    cp_async_wait_all();
                    __syncwarp();
                    if (i + 1 < n) fetch_T(i + 1, (i & 1) ? Cp : Ut, (i & 1) ? cB : cA);
                    gemv_t<true, true>(cc, Tc, ycur, lane);   // y_i = c_i - T^T y_{i-1}
                    if (lane < 15) { const double y = cc[lane]; ycur[lane] = y; sb[15 * i + lane] = y; }
                    __syncwarp();
                }
                step_and_model();
            }
        } else {
        // Dm/Cy swap roles every frame (pivot block <-> block being assembled)
            double* Dp = Dm;
            double* Cp = Cy;
            assemble_D(n - 1, Dp);
            scale_damp(n - 1, Dp);
            bool have_wc = false, have_d0 = false;
            for (int i = n - 1; i >= 1; --i) {
                // raw blocks of this elimination step, fetched asynchronously while the pivot is inverted:
                // Um <- H(i-1, i) of item i, Cp <- hbb of item i-1, Tm <- haa of item i  (D_{i-1} = hbb + haa + laser)
                {
                    const double* it_i = itm + (size_t)i * kItem;
                    const double* it_p = itm + (size_t)(i - 1) * kItem;
                    for (int e = lane; e < kBlk; e += NT) {
                        const uint32_t pk = otab[e];
                        cp_async8(Um + e, it_i + (pk & 1023u));
                        cp_async8(Cp + e, it_p + ((pk >> 10) & 1023u));
                        cp_async8(Tm + e, it_i + (pk >> 20));
                    }
                    for (int e = lane; e < NPAD; e += NT) cp_async8(slb + e, lb + (i - 1) * NPAD + e);
                    if (lane < 30) cp_async8(ssc + lane, scw + (i - 1) * 15 + lane);
                }
                // per-frame flags, requested now and consumed after the inverse
                const uint8_t cm_p = mode == 1 ? 0 : cm[i - 1];
                const bool fa_p = fa[i - 1] != 0;
                const bool cross_i = has_cross(i);
                bool arrow_i = false;
                if (ARROW && i >= 2) {
                    arrow_i = have_wc || cross_i;
                    if (arrow_i)
                        for (int e = lane; e < kBlk; e += NT) {
                            const int r = e / 15, c = e - r * 15;
                            double v = have_wc ? Wm[e] : 0.0;
                            if (cross_i && r < 6 && c < 6 && !is_const(0, r) && !is_const(i, c)) v += cross_entry(i, r, c) * scw[r] * scw[i * 15 + c];
                            Wm[e] = v;
                        }
                }
                grp_sync<NT>();
                const bool inv_ok = grp_inverse15<NT>(Dp, piv, red, lane);
                cp_async_wait_all();
                grp_sync<NT>();
                if (!inv_ok) { ok = false; break; }
                // coupling U = H(i-1, i), scaled (+ laser cross block / carried fill-in when i == 1: H(0,1) is also the arrow
                // block); D_{i-1} assembled, scaled and damped
                for (int e = lane; e < kBlk; e += NT) {
                    const int r = e / 15, c = e - r * 15;
                    double v = Um[e] * ssc[r] * ssc[15 + c];
                    if (ARROW && i == 1) {
                        if (have_wc) v += Wm[e];
                        if (cross_i && r < 6 && c < 6 && !is_const(0, r) && !is_const(1, c)) v += cross_entry(1, r, c) * scw[r] * scw[15 + c];
                    }
                    Um[e] = v;
                    double dv = Cp[e];
                    dv += Tm[e];
                    dv += laser_own_reg(slb, cm_p, fa_p, r, c);
                    if (i - 1 == 0) dv += laser_ref_own(r, c);
                    dv = dv * ssc[r] * ssc[c];
                    if (r == c) dv = col_const(cm_p, r) ? 1.0 : dv + fmin(fmax(dv, opt.min_lm_diagonal), opt.max_lm_diagonal) * inv_radius;
                    Cp[e] = dv;
                }
                grp_sync<NT>();
                gemm_ab15<NT>(Tm, Um, Dp, lane);                                   // T = U Dinv
                if (ARROW && arrow_i) gemm_ab15<NT>(Tp, Wm, Dp, lane);             // T' = W Dinv
                gemv15<NT, false, false>(piv, Dp, sb + 15 * i, lane);              // c_i = Dinv b_i
                if (lane < 15) sb[15 * i + lane] = piv[lane];
                grp_sync<NT>();
                for (int e = lane; e < kBlk; e += NT) {
                    facw[(size_t)i * 3 * kBlk + e] = Tm[e];
                    if (ARROW) facw[(size_t)i * 3 * kBlk + kBlk + e] = arrow_i ? Tp[e] : 0.0;
                }
                gemv15<NT, false, true>(sb + 15 * (i - 1), Um, sb + 15 * i, lane);  // b_{i-1} -= U c_i
                if (ARROW && arrow_i) gemv15<NT, false, true>(sb, Wm, sb + 15 * i, lane);
                gemm_sub_abt15<NT>(Cp, Tm, Um, lane);                               // D_{i-1} -= T U^T
                if (ARROW && arrow_i) {
                    if (!have_d0) { for (int e = lane; e < kBlk; e += NT) D0[e] = 0.0; grp_sync<NT>(); have_d0 = true; }
                    gemm_sub_abt15<NT>(D0, Tp, Wm, lane);                           // D_0 -= T' W^T
                    // fill-in for frame i-1: H(0, i-1) = -T' U^T   (Wm is free again after this; the old pivot block is scratch)
                    for (int e = lane; e < kBlk; e += NT) Dp[e] = 0.0;
                    grp_sync<NT>();
                    gemm_sub_abt15<NT>(Dp, Tp, Um, lane);
                    copy_blk<NT>(Wm, Dp, lane);
                    grp_sync<NT>();
                    have_wc = true;
                } else if (ARROW) {
                    have_wc = false;
                }
                if (ARROW && i - 1 == 0 && have_d0) {
                    for (int e = lane; e < kBlk; e += NT) Cp[e] += D0[e];
                    grp_sync<NT>();
                }
                { double* t = Dp; Dp = Cp; Cp = t; }
            }
            if (ok) ok = grp_inverse15<NT>(Dp, piv, red, lane);
            if (ok) {
                gemv15<NT, false, false>(piv, Dp, sb, lane);   // y_0 = D0inv b_0
                if (lane < 15) sb[lane] = piv[lane];
                grp_sync<NT>();
                // T_i (and T'_i) come back from global memory one frame ahead of their use (Tm / Um alternate)
                auto fetch_T = [&](int i, double* dstT) {
                    for (int e = lane; e < kBlk; e += NT) cp_async8(dstT + e, facw + (size_t)i * 3 * kBlk + e);
                };
                if (n > 1) fetch_T(1, Tm);
                for (int i = 1; i < n; ++i) {
                    double* Tc = (i & 1) ? Tm : Um;
                    if (ARROW && i >= 2) for (int e = lane; e < kBlk; e += NT) cp_async8(Tp + e, facw + (size_t)i * 3 * kBlk + kBlk + e);
                    cp_async_wait_all();
                    grp_sync<NT>();
                    if (i + 1 < n) fetch_T(i + 1, (i & 1) ? Um : Tm);
                    gemv15<NT, true, true>(sb + 15 * i, Tc, sb + 15 * (i - 1), lane);   // y_i = c_i - T^T y_{i-1} - T'^T y_0
                    if (ARROW && i >= 2) gemv15<NT, true, true>(sb + 15 * i, Tp, sb, lane);
                }
                step_and_model();
            }
        }
        grp_sync<NT>();
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
    if (!have_step) {
        if (lane == 0) { a.state[w] = st; a.win_status[w] = st.status; }
        return;
    }

    // ---- (4) candidate = Plus(x, step * scale) and its laser frame tables
    for (int f = lane; f < n; f += NT) {
        double d[15];
#pragma unroll
        for (int c = 0; c < 15; ++c) d[c] = sb[f * 15 + c] * scw[f * 15 + c];
#pragma unroll
        for (int c = 0; c < 15; ++c) xc[15 * f + c] = is_const(f, c) ? x[15 * f + c] : x[15 * f + c] + d[c];
        if (!is_const(f, 3)) so3_plus(x + 15 * f + 3, d + 3, xc + 15 * f + 3);
        laser_frame_table(a.C, xc + 15 * f, a.frame_tab + ((size_t)w * n + f) * kFrameTab);
    }
    if (lane == 0) { a.state[w] = st; a.win_status[w] = 0; }
}

template <bool ARROW, int NT>
__global__ void __launch_bounds__(NT == 32 ? 32 * LV_WINDOW_WPC : NT, NT == 32 ? LV_WINDOW_WARPS / LV_WINDOW_WPC : 512 / NT) window_kernel(WindowArgs a, int per_window_doubles) {
    extern __shared__ __align__(16) double smem[];
    constexpr bool FASTC = !ARROW && NT == 32;   // this instantiation never runs the generic elimination loop
    __shared__ uint32_t otab[FASTC ? 1 : kBlk];
    if (!FASTC) {
        window_offsets_init(otab, threadIdx.x, blockDim.x);
        __syncthreads();
    }
    if (NT == 32) {
        // batched shape: one warp per window, LV_WINDOW_WPC windows per CTA
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int w = blockIdx.x * (blockDim.x >> 5) + warp;
        if (w >= a.n_windows) return;
        window_step<ARROW, NT>(a, w, lane, smem + (size_t)warp * per_window_doubles, otab);
    } else {
        // small batches: one CTA of NT threads per window
        window_step<ARROW, NT>(a, blockIdx.x, threadIdx.x, smem, otab);
    }
}

// =====================================================================================================
// solve_small_kernel: the WHOLE minimiser loop of a small batch in one launch, one CTA of 8 warps per window.
// The reference's steady-state problem is tiny (2 frames, a few matched segment pairs, up to 50 iterations): launched
// as three kernels per trip it is bound by ~150 launch + drain latencies.  Here a trip is scan-match items (a warp per
// frame), factor item pairs (a warp per pair), the window step on all 256 threads, separated by __syncthreads(); the
// three phases alias the same shared memory.  Same device functions, same arithmetic as the batched kernels.
template <bool ARROW, bool HAS_WEIGHT>
__global__ void __launch_bounds__(256, 1) solve_small_kernel(ScanMatchArgs sa, WindowArgs wa, int trips) {
    extern __shared__ __align__(16) double smem[];
    constexpr int ROW = ARROW ? kRowFree : kRowTrack;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = blockIdx.x;
    const int n = wa.n_frames;
    __shared__ uint32_t otab[kBlk];
    window_offsets_init(otab, tid, 256);
    __syncthreads();
    for (int trip = 0; trip < trips; ++trip) {
        if (wa.win_status[w] != 0) break;                       // uniform: written before the last __syncthreads()
        if (sa.points)
            for (int k = warp; k < n * sa.tiles; k += 8)
                scan_match_item<ARROW, HAS_WEIGHT, false, false>(sa, w * n * sa.tiles + k, lane, smem + (size_t)warp * sa.line_cap * ROW);
        __syncthreads();                                        // the factor phase reuses the line-table memory
        for (int k = 2 * warp; k < n; k += 16)
            factor_pair_item(wa, w * n + k, (k + 1 < n) ? 2 : 1, lane, smem + (size_t)warp * kPairSmem);
        __syncthreads();                                        // partials and items are complete (block-visible)
        window_step<ARROW, 256>(wa, w, tid, smem, otab);
        __syncthreads();
    }
}

__global__ void init_state_kernel(LMState* st, int32_t* status, int n, double radius) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    LMState s;
    s.radius = radius; s.decrease_factor = 2.0; s.cost = 0.0; s.model_cost_change = 0.0; s.x_norm = 0.0; s.initial_cost = 0.0;
    s.iteration = 0; s.num_invalid = 0; s.status = 0; s.termination = 0; s.n_success = 0; s.n_unsuccess = 0; s.last_success = 0; s.started = 0;
    s.cur = 0; s.pad0 = 0;
    st[w] = s;
    status[w] = 0;
}

// frame tables of arbitrary poses ([F][stride] -> [F][24]); used for the initial point and the external reference poses
__global__ void frame_table_kernel(Consts C, const double* poses, int stride, double* tabs, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    laser_frame_table(C, poses + (size_t)f * stride, tabs + (size_t)f * kFrameTab);
}

}  // namespace lv
