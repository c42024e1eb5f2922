// pose_graph.cuh — back-end pose-graph solve on the device.
//
// What it replaces (reference file:line):
//   keyframe_manager::solve   src/trajectory/keyframe_manager.cpp:722-838   sequential + loop edges, ground factors on every
//                                                                          key frame, first key frame constant, ceres::Solve
//                                                                          with default options (LM, SPARSE_SCHUR)
//   edge_factor               src/factor/edge_factor.h:79-126             r = w * J_noise * log_SE3(T_j^-1 T_i T_12)
//   ground_factor_p / q       src/factor/ground_factor.h:27-82            (lv_math.cuh: ground_residuals)
//
// Structure of the normal equations.  Key frames are 6-DoF nodes; an edge between neighbours (|i - j| = 1, the
// reference's seq_edges) puts a 6x6 block on the first off-diagonal, every other edge (loop_edges) is a rank-6 update:
//     H + D_lm = T + U U^T,   T block-tridiagonal (band edges + ground factors + LM diagonal),   U = [J_e^T]_{e in loops}
// so one LM step is
//   (1) a block-tridiagonal factorisation of T (one CTA, 36 threads = one 6x6 block, sequential in k),
//   (2) T^-1 [-g | U] for 1 + 6L right-hand sides (one thread each),
//   (3) the 6L x 6L capacitance system  (I + U^T T^-1 U) w = U^T T^-1 (-g)  (one CTA),
//   (4) delta = x0 - Z w.
// Nothing dense of size 6K is ever formed; the solve is exact (as the reference's SPARSE_SCHUR is), not iterative.
// pose_graph_segments.cuh (opt-in) cuts the two sequential chains of (1) and (2) into concurrently processed segments.
//
// How the code is organised.  Every kernel is `thread t of n runs pg_thread<KID>(args, t)` with no intra-block
// communication; the cooperative kernels (the block-tridiagonal factorisation, the dense capacitance Cholesky and the
// reduction of the per-item partial sums) are written as phases separated by __syncthreads.  The bodies are
// __host__ __device__, so the CPU test suite (tests/native/pose_graph_host.cpp) runs the SAME bodies and the SAME
// minimiser loop (pg_minimize) thread by thread and checks them against the oracle — a formula / control-flow check without a GPU, not a fallback: the product
// entry point (lvio2d_pose_graph_solve, lvio2d_api.cu) only ever launches the kernels.
//
// The minimiser is Ceres 1.14's TrustRegionMinimizer + LevenbergMarquardtStrategy (jacobi scaling, monotonic steps), the
// same restatement the sliding window uses (window.cuh), with the scalar accept / reject logic on the host: a back-end
// solve is a single problem (nothing to batch), runs on the back-end thread, and needs two small read-backs per iteration.
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstddef>

#include "../../include/lvio2d.h"
#include "lv_math.cuh"

namespace lv {
namespace pg {

enum Kernel { K_COLUMNS = 0, K_ASSEMBLE, K_SCALE, K_TRISOLVE, K_CAPACITANCE, K_COMBINE, K_MODEL, K_COST };
enum Scalar { S_COST = 0, S_YNORM, S_STEP, S_MODEL, S_DG, S_GRAD, S_COUNT };

struct Options {
    int max_iters = 50;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32;
    double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    int max_consecutive_invalid = 5;
};

struct Args {
    int K, E, L, ncol;      // key frames, edges, loop (non-band) edges, 1 + 6L right-hand sides
    int fixed;              // constant key frame (index1 of the first edge, keyframe_manager.cpp:744-748) or -1
    int ground_p, ground_q;
    double radius, min_lm, max_lm;
    Consts C;
    double Jn[36];          // edge_noise::J, row-major (edge_factor.h:14-26)
    // inputs
    const int32_t* edge_index;   // [E][2]
    const double* edge_tf;       // [E][12] row-major 3x4
    const double* edge_weight;   // [E]
    const int32_t* edge_band;    // [E] 1: |i - j| == 1
    const int32_t* loop_edge;    // [L] ids of the non-band edges
    const int32_t* inc_off;      // [K+1] incidence lists: entries 2 * edge + side (0: the pose is index1)
    const int32_t* inc;
    // state
    double* x;       // [K][6] (p, q)
    double* xc;      // candidate
    const double* y; // the state K_COST evaluates (x or xc)
    // work
    double* EJ;      // [E][13][6]: Jacobian columns 0..11 over (p_i q_i p_j q_j), column 12 = residual
    double* GJ;      // [K][7][2]: ground (p, q) Jacobian columns 0..5, column 6 = residuals
    double* g;       // [6K] gradient J^T r
    double* Hd;      // [6K] diag(J^T J)
    double* D;       // [K][36] band diagonal blocks (no LM term)
    double* O;       // [K][36] block (k+1, k)
    double* scale;   // [6K] jacobi scaling (iteration 0)
    double* Sinv;    // [K][36] inverse Schur pivots of the tridiagonal factorisation
    double* M;       // [K][36] O[k-1] * Sinv[k-1]
    double* Z;       // [6K][ncol] T^-1 [-g | U]
    double* Cm;      // [6L][ncol] column 0: U^T x0, columns 1..: I + U^T Z  (factored in place)
    double* w;       // [6L]
    double* delta;   // [6K]
    double* part;    // per-item partials, see part_offset
    double* scal;    // [S_COUNT]
    int32_t* flags;  // [2] non-positive pivot in the tridiagonal / capacitance factorisation
    // partitioned solve (pose_graph_segments.cuh; P <= 1: off, everything below unused)
    int P;                     // segments
    int stage;                 // != 0: segment solves stage their blocks in shared memory (LVIO2D_PG_STAGE)
    const int32_t* node_seg;   // [K] segment of an interior key frame, or -(i + 1) for separator i
    int32_t* segflag;          // [P]
    double* Zx;                // [6K][ncol + 12] segment solutions: right-hand sides + the two spikes
    double* Rd;                // [P-1][36] reduced (separator) system: diagonal blocks,
    double* Ro;                // [P-1][36] block (i + 1, i),
    double* RSinv;             // its factorisation
    double* RM;
    double* Zr;                // [6(P-1)][ncol] reduced right-hand sides / solutions
};

// read-only (non-coherent) load and L1 prefetch on the device, plain load / nothing on the host
LV_HD double lv_ldg(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
LV_HD void lv_prefetch(const void* p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

LV_HD int part_count(const Args& a, int q) { return (q == S_COST || q == S_MODEL) ? a.E + a.K : a.K; }
LV_HD int part_offset(const Args& a, int q) {
    int off = 0;
    for (int i = 0; i < q; ++i) off += part_count(a, i);
    return off;
}
LV_HD int part_total(const Args& a) { return part_offset(a, S_COUNT); }

// ------------------------------------------------------------------ edge_factor (edge_factor.h:93-119)
template <class T>
LV_HD void edge_residual(const Iso& T12, double weight, const double* Jn, const V3<T>& pi, const V3<T>& thi, const V3<T>& pj, const V3<T>& thj,
                         T* res /*[6]*/) {
    const M3<T> Ri = exp_so3(thi), Rj = exp_so3(thj);
    // error = tf_j^-1 * tf_i * tf12
    const M3<T> Re = mul_tn(Rj, mul(Ri, lift<T>(T12.R)));
    const V3<T> te = mul_t(Rj, mul(Ri, lift<T>(T12.t)) + pi - pj);
    const V3<T> rq = log_so3(Re);
    const T raw[6] = {te.x, te.y, te.z, rq.x, rq.y, rq.z};
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        T s = T(0.0);
#pragma unroll
        for (int k = 0; k < 6; ++k) s = s + Jn[r * 6 + k] * raw[k];
        res[r] = weight * s;
    }
}
LV_HD void seed(V3<Dual>* blocks, int c) {
    V3<Dual>& t = blocks[c / 3];
    (c % 3 == 0 ? t.x : (c % 3 == 1 ? t.y : t.z)).d = 1.0;
}
// column c of edge e: 0..11 = d r / d (p_i q_i p_j q_j)[c], 12 = r
LV_HD void edge_column(const Args& a, int e, int c, double* out /*[6]*/) {
    const int i = a.edge_index[2 * e], j = a.edge_index[2 * e + 1];
    const double* xi = a.x + 6 * i;
    const double* xj = a.x + 6 * j;
    const Iso T12 = load_iso(a.edge_tf + 12 * (size_t)e);
    const double w = a.edge_weight[e];
    if (c == 12) {
        edge_residual<double>(T12, w, a.Jn, load3(xi), load3(xi + 3), load3(xj), load3(xj + 3), out);
        return;
    }
    if ((c < 6 ? i : j) == a.fixed) {
        for (int r = 0; r < 6; ++r) out[r] = 0.0;
        return;
    }
    V3<Dual> v[4] = {lift<Dual>(load3(xi)), lift<Dual>(load3(xi + 3)), lift<Dual>(load3(xj)), lift<Dual>(load3(xj + 3))};
    seed(v, c);
    Dual r[6];
    edge_residual<Dual>(T12, w, a.Jn, v[0], v[1], v[2], v[3], r);
    for (int k = 0; k < 6; ++k) out[k] = r[k].d;
}
// column c of the ground factors of key frame k: 0..5 = d (r_p, r_q) / d (p q)[c], 6 = (r_p, r_q)
LV_HD void ground_column(const Args& a, int k, int c, double* out /*[2]*/) {
    const double* xk = a.x + 6 * k;
    double rp = 0.0, rq = 0.0;
    if (c == 6) {
        ground_residuals<double>(a.C, load3(xk), load3(xk + 3), &rp, &rq);
    } else if (k != a.fixed) {
        V3<Dual> v[2] = {lift<Dual>(load3(xk)), lift<Dual>(load3(xk + 3))};
        seed(v, c);
        Dual dp, dq;
        ground_residuals<Dual>(a.C, v[0], v[1], &dp, &dq);
        rp = dp.d; rq = dq.d;
    }
    out[0] = a.ground_p ? rp : 0.0;
    out[1] = a.ground_q ? rq : 0.0;
}

// ------------------------------------------------------------------ kernel bodies: thread t of n
// K_COLUMNS, n = 13 E + 7 K: every Jacobian column (one dual-number pass each) and the residuals at x
LV_HD void body_columns(const Args& a, int t) {
    if (t < 13 * a.E) {
        edge_column(a, t / 13, t % 13, a.EJ + (size_t)t * 6);
    } else {
        const int u = t - 13 * a.E;
        ground_column(a, u / 7, u % 7, a.GJ + (size_t)u * 2);
    }
}
// K_ASSEMBLE, n = K: gradient, diag(H), band blocks of key frame k, and its term of the gradient max-norm
LV_HD void body_assemble(const Args& a, int k) {
    double g[6], hd[6], D[36], O[36];
    for (int q = 0; q < 6; ++q) { g[q] = 0.0; hd[q] = 0.0; }
    for (int q = 0; q < 36; ++q) { D[q] = 0.0; O[q] = 0.0; }
    for (int u = a.inc_off[k]; u < a.inc_off[k + 1]; ++u) {
        const int e = a.inc[u] >> 1, side = a.inc[u] & 1;
        const double* J = a.EJ + (size_t)e * 78;
        const double* Jk = J + side * 36;   // Jk[q * 6 + r] = d r[r] / d x_k[q]
        const double* r = J + 72;
        const bool band = a.edge_band[e] != 0;
        for (int q = 0; q < 6; ++q) {
            double s = 0.0;
            for (int i = 0; i < 6; ++i) s += Jk[q * 6 + i] * r[i];
            g[q] += s;
        }
        for (int q = 0; q < 6; ++q)
            for (int q2 = 0; q2 < 6; ++q2) {
                double s = 0.0;
                for (int i = 0; i < 6; ++i) s += Jk[q * 6 + i] * Jk[q2 * 6 + i];
                if (q == q2) hd[q] += s;
                if (band) D[q * 6 + q2] += s;
            }
        if (band && a.edge_index[2 * e + (1 - side)] == k + 1) {
            const double* Jo = J + (1 - side) * 36;
            for (int q2 = 0; q2 < 6; ++q2)     // row: key frame k + 1
                for (int q = 0; q < 6; ++q) {  // column: key frame k
                    double s = 0.0;
                    for (int i = 0; i < 6; ++i) s += Jo[q2 * 6 + i] * Jk[q * 6 + i];
                    O[q2 * 6 + q] += s;
                }
        }
    }
    const double* Jg = a.GJ + (size_t)k * 14;
    for (int q = 0; q < 6; ++q) {
        g[q] += Jg[q * 2] * Jg[12] + Jg[q * 2 + 1] * Jg[13];
        for (int q2 = 0; q2 < 6; ++q2) {
            const double s = Jg[q * 2] * Jg[q2 * 2] + Jg[q * 2 + 1] * Jg[q2 * 2 + 1];
            D[q * 6 + q2] += s;
            if (q == q2) hd[q] += s;
        }
    }
    double gmax = 0.0;
    if (k == a.fixed) {
        for (int q = 0; q < 36; ++q) D[q] = (q % 7 == 0) ? 1.0 : 0.0;
        for (int q = 0; q < 6; ++q) { g[q] = 0.0; hd[q] = 0.0; }
    } else {
        // |x - Plus(x, -g)|_inf (TrustRegionMinimizer::EvaluateGradientAndJacobian)
        const double* xk = a.x + 6 * k;
        const double neg[3] = {-g[3], -g[4], -g[5]};
        double qn[3];
        so3_plus(xk + 3, neg, qn);
        for (int q = 0; q < 3; ++q) {
            gmax = fmax(gmax, fabs(g[q]));
            gmax = fmax(gmax, fabs(xk[3 + q] - qn[q]));
        }
    }
    for (int q = 0; q < 6; ++q) { a.g[6 * k + q] = g[q]; a.Hd[6 * k + q] = hd[q]; }
    for (int q = 0; q < 36; ++q) { a.D[(size_t)k * 36 + q] = D[q]; a.O[(size_t)k * 36 + q] = O[q]; }
    a.part[part_offset(a, S_GRAD) + k] = gmax;
}
// K_SCALE, n = 6K: jacobi scaling 1 / (1 + sqrt(diag(J^T J))), taken once at iteration 0
LV_HD void body_scale(const Args& a, int c) { a.scale[c] = 1.0 / (1.0 + sqrt(a.Hd[c])); }
// the block-tridiagonal factorisation of T = band(H) + LM diagonal, one CTA of 36 working threads (one per entry of a
// 6x6 block), sequential over the key frames, as phases with a __syncthreads between them:
//     S_k = D_k + lm_k - M_k B_k^T,   M_k = B_k S_{k-1}^-1,   B_k = block (k, k-1);   stores S_k^-1 and M_k.
// The inverse is a Gauss-Jordan sweep without pivoting (S_k is SPD; its pivots are the squares of the Cholesky ones, so
// "pivot <= 0" is the same not-positive-definite test), ping-ponging between two copies of [S | Inv] so that a sweep
// step is one phase.  Every phase reads what earlier phases wrote and writes only its own entry t (of arrays no thread
// reads in that phase), so running the 36 threads of a phase one after the other (CPU check) is equivalent.
// The operands of step k + 1 (D, B, diag(H), scale) are fetched while step k runs (factor_fetch / factor_stage).
struct FactorTile { double S[2][36], Inv[2][36], Sp[36], Mk[36], B[36], nD[36], nO[36], nHd[6], nSc[6]; int ok; };
struct FactorNext { double D, O, Hd, Sc; };
enum { FACTOR_PHASES = 9, FACTOR_THREADS = 36 };
LV_HD FactorNext factor_fetch(const Args& a, int k, int t) {
    FactorNext n;
    n.D = a.D[(size_t)k * 36 + t];
    n.O = k > 0 ? a.O[(size_t)(k - 1) * 36 + t] : 0.0;
    n.Hd = t < 6 ? a.Hd[6 * k + t] : 0.0;
    n.Sc = t < 6 ? a.scale[6 * k + t] : 1.0;
    return n;
}
LV_HD void factor_stage(FactorTile& T, int t, const FactorNext& n) {
    T.nD[t] = n.D;
    T.nO[t] = n.O;
    if (t < 6) { T.nHd[t] = n.Hd; T.nSc[t] = n.Sc; }
}
LV_HD void factor_phase(const Args& a, FactorTile& T, int k, int phase, int t) {
    const int r = t / 6, c = t % 6;
    if (phase == 0) {
        double v = T.nD[t];
        if (r == c && k != a.fixed) {
            const double s = T.nSc[r], hs = T.nHd[r] * s * s;
            const double d = fmin(fmax(hs, a.min_lm), a.max_lm);
            v += d / a.radius / (s * s);   // the scaled system's diagonal / radius, back in unscaled columns
        }
        T.S[0][t] = v;
        T.B[t] = T.nO[t];
        T.Inv[0][t] = (r == c) ? 1.0 : 0.0;
        if (t == 0 && k == 0) T.ok = 1;
    } else if (phase == 1) {
        double m = 0.0;
        if (k > 0)
            for (int i = 0; i < 6; ++i) m += T.B[r * 6 + i] * T.Sp[i * 6 + c];
        T.Mk[t] = m;
        a.M[(size_t)k * 36 + t] = m;
    } else if (phase == 2) {
        if (k > 0) {
            double s = 0.0;
            for (int i = 0; i < 6; ++i) s += T.Mk[r * 6 + i] * T.B[c * 6 + i];
            T.S[0][t] -= s;
        }
    } else {
        const int j = phase - 3, src = j & 1, dst = src ^ 1;
        double p = T.S[src][j * 6 + j];
        if (!(p > 0.0)) { if (t == 0) T.ok = 0; p = 1.0; }
        const double ip = 1.0 / p;
        const double sj = T.S[src][j * 6 + c] * ip, ij = T.Inv[src][j * 6 + c] * ip;
        double ns, ni;
        if (r == j) { ns = sj; ni = ij; }
        else { const double f = T.S[src][r * 6 + j]; ns = T.S[src][t] - f * sj; ni = T.Inv[src][t] - f * ij; }
        T.S[dst][t] = ns;
        T.Inv[dst][t] = ni;
        if (j == 5) {   // dst == 0: the inverse is complete
            T.Sp[t] = ni;
            a.Sinv[(size_t)k * 36 + t] = ni;
        }
    }
}
// K_TRISOLVE, n = ncol: T z = b for one right-hand side (0: -g, 1 + 6l + rho: row rho of loop edge l's Jacobian).
// Sequential in k per thread; the chain only carries 6 doubles, so the cost of a step is the latency of fetching
// M_k / Sinv_k: they are read through the read-only path (free to move above the Z stores) and prefetched PG_AHEAD
// steps ahead.  Measured (profiles/r1_pose_graph.md): still 1.5 us per step and sweep — ptxas keeps only 8 of the 36
// loads of a block in flight and they are served by L2; round 2 stages the blocks of 16 steps in shared memory.
#define PG_AHEAD 4
LV_HD void body_trisolve(const Args& a, int col) {
    int e = -1, rho = 0, ie = -1, je = -1;
    if (col > 0) {
        e = a.loop_edge[(col - 1) / 6];
        rho = (col - 1) % 6;
        ie = a.edge_index[2 * e];
        je = a.edge_index[2 * e + 1];
    }
    const double* J = e >= 0 ? a.EJ + (size_t)e * 78 : nullptr;
    double yp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, b[6];
    for (int k = 0; k < a.K; ++k) {
        if (k + PG_AHEAD < a.K) {
            const double* Mp = a.M + (size_t)(k + PG_AHEAD) * 36;
            lv_prefetch(Mp); lv_prefetch(Mp + 16); lv_prefetch(Mp + 32);
            if (col == 0) lv_prefetch(a.g + 6 * (k + PG_AHEAD));
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            if (col == 0) b[q] = -lv_ldg(a.g + 6 * k + q);
            else b[q] = (k == ie) ? J[q * 6 + rho] : ((k == je) ? J[36 + q * 6 + rho] : 0.0);
        }
        if (k > 0) {
            const double* Mk = a.M + (size_t)k * 36;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i) s += lv_ldg(Mk + q * 6 + i) * yp[i];
                b[q] -= s;
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) { yp[q] = b[q]; a.Z[(size_t)(6 * k + q) * a.ncol + col] = b[q]; }
    }
    double xn[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int k = a.K - 1; k >= 0; --k) {
        if (k - PG_AHEAD >= 0) {
            const double* Sp = a.Sinv + (size_t)(k - PG_AHEAD) * 36;
            lv_prefetch(Sp); lv_prefetch(Sp + 16); lv_prefetch(Sp + 32);
            const double* Mp = a.M + (size_t)(k - PG_AHEAD + 1) * 36;
            lv_prefetch(Mp); lv_prefetch(Mp + 16); lv_prefetch(Mp + 32);
#pragma unroll
            for (int q = 0; q < 6; ++q) lv_prefetch(a.Z + (size_t)(6 * (k - PG_AHEAD) + q) * a.ncol + col);
        }
        const double* Si = a.Sinv + (size_t)k * 36;
#pragma unroll
        for (int q = 0; q < 6; ++q) yp[q] = a.Z[(size_t)(6 * k + q) * a.ncol + col];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) s += lv_ldg(Si + q * 6 + i) * yp[i];
            b[q] = s;
        }
        if (k < a.K - 1) {
            const double* Mn = a.M + (size_t)(k + 1) * 36;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i) s += lv_ldg(Mn + i * 6 + q) * xn[i];
                b[q] -= s;
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) { xn[q] = b[q]; a.Z[(size_t)(6 * k + q) * a.ncol + col] = b[q]; }
    }
}
// K_CAPACITANCE, n = 6L * ncol: [U^T x0 | I + U^T T^-1 U]
LV_HD void body_capacitance(const Args& a, int t) {
    const int row = t / a.ncol, col = t % a.ncol;
    const int e = a.loop_edge[row / 6], rho = row % 6;
    const int ie = a.edge_index[2 * e], je = a.edge_index[2 * e + 1];
    const double* J = a.EJ + (size_t)e * 78;
    double s = (col == 1 + row) ? 1.0 : 0.0;
    for (int q = 0; q < 6; ++q)
        s += J[q * 6 + rho] * a.Z[(size_t)(6 * ie + q) * a.ncol + col] + J[36 + q * 6 + rho] * a.Z[(size_t)(6 * je + q) * a.ncol + col];
    a.Cm[(size_t)row * a.ncol + col] = s;
}
// the capacitance system, one CTA: right-looking Cholesky in phases (a __syncthreads between them), then the two
// triangular solves on thread 0.  Cm(i, j) = Cm[i * ncol + 1 + j].
LV_HD void dense_phase_pivot(const Args& a, int j, int tid) {
    if (tid != 0) return;
    double d = a.Cm[(size_t)j * a.ncol + 1 + j];
    if (!(d > 0.0)) { a.flags[1] = 1; d = 1.0; }
    a.Cm[(size_t)j * a.ncol + 1 + j] = sqrt(d);
}
LV_HD void dense_phase_column(const Args& a, int j, int tid, int nt) {
    const int n = 6 * a.L;
    const double piv = a.Cm[(size_t)j * a.ncol + 1 + j];
    for (int i = j + 1 + tid; i < n; i += nt) a.Cm[(size_t)i * a.ncol + 1 + j] /= piv;
}
LV_HD void dense_phase_update(const Args& a, int j, int tid, int nt) {
    const int n = 6 * a.L;
    for (int i = j + 1 + tid; i < n; i += nt) {
        const double lij = a.Cm[(size_t)i * a.ncol + 1 + j];
        for (int k = j + 1; k <= i; ++k) a.Cm[(size_t)i * a.ncol + 1 + k] -= lij * a.Cm[(size_t)k * a.ncol + 1 + j];
    }
}
LV_HD void dense_phase_solve(const Args& a, int tid) {
    if (tid != 0) return;
    const int n = 6 * a.L;
    for (int i = 0; i < n; ++i) {
        double s = a.Cm[(size_t)i * a.ncol];
        for (int k = 0; k < i; ++k) s -= a.Cm[(size_t)i * a.ncol + 1 + k] * a.w[k];
        a.w[i] = s / a.Cm[(size_t)i * a.ncol + 1 + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = a.w[i];
        for (int k = i + 1; k < n; ++k) s -= a.Cm[(size_t)k * a.ncol + 1 + i] * a.w[k];
        a.w[i] = s / a.Cm[(size_t)i * a.ncol + 1 + i];
    }
}
// K_COMBINE, n = K: delta = x0 - Z w, candidate = Plus(x, delta) (so3_parameterization on q), and delta^T g
LV_HD void body_combine(const Args& a, int k) {
    double d[6];
    const int n = 6 * a.L;
    for (int q = 0; q < 6; ++q) {
        const double* z = a.Z + (size_t)(6 * k + q) * a.ncol;
        double s = z[0];
        for (int i = 0; i < n; ++i) s -= z[1 + i] * a.w[i];
        d[q] = (k == a.fixed) ? 0.0 : s;
    }
    const double* xk = a.x + 6 * k;
    double* ck = a.xc + 6 * k;
    double dg = 0.0;
    for (int q = 0; q < 6; ++q) { a.delta[6 * k + q] = d[q]; dg += d[q] * a.g[6 * k + q]; }
    for (int q = 0; q < 3; ++q) ck[q] = (k == a.fixed) ? xk[q] : xk[q] + d[q];
    if (k == a.fixed) { for (int q = 3; q < 6; ++q) ck[q] = xk[q]; }
    else so3_plus(xk + 3, d + 3, ck + 3);
    a.part[part_offset(a, S_DG) + k] = dg;
}
// K_MODEL, n = E + K: |J delta|^2 per residual block (delta^T H delta without forming H)
LV_HD void body_model(const Args& a, int t) {
    double s2 = 0.0;
    if (t < a.E) {
        const double* J = a.EJ + (size_t)t * 78;
        const double* di = a.delta + 6 * a.edge_index[2 * t];
        const double* dj = a.delta + 6 * a.edge_index[2 * t + 1];
        for (int r = 0; r < 6; ++r) {
            double v = 0.0;
            for (int q = 0; q < 6; ++q) v += J[q * 6 + r] * di[q] + J[36 + q * 6 + r] * dj[q];
            s2 += v * v;
        }
    } else {
        const int k = t - a.E;
        const double* Jg = a.GJ + (size_t)k * 14;
        const double* dk = a.delta + 6 * k;
        double vp = 0.0, vq = 0.0;
        for (int q = 0; q < 6; ++q) { vp += Jg[q * 2] * dk[q]; vq += Jg[q * 2 + 1] * dk[q]; }
        s2 = vp * vp + vq * vq;
    }
    a.part[part_offset(a, S_MODEL) + t] = s2;
}
// K_COST, n = E + K: squared residuals at y, |y|^2 and |x - y|^2 over the free key frames
LV_HD void body_cost(const Args& a, int t) {
    if (t < a.E) {
        const double* yi = a.y + 6 * a.edge_index[2 * t];
        const double* yj = a.y + 6 * a.edge_index[2 * t + 1];
        double r[6];
        edge_residual<double>(load_iso(a.edge_tf + 12 * (size_t)t), a.edge_weight[t], a.Jn, load3(yi), load3(yi + 3), load3(yj), load3(yj + 3), r);
        double s2 = 0.0;
        for (int q = 0; q < 6; ++q) s2 += r[q] * r[q];
        a.part[part_offset(a, S_COST) + t] = s2;
    } else {
        const int k = t - a.E;
        const double* yk = a.y + 6 * k;
        double rp, rq;
        ground_residuals<double>(a.C, load3(yk), load3(yk + 3), &rp, &rq);
        // (the ground factors of the constant key frame have no variable block: Ceres counts them as fixed cost)
        a.part[part_offset(a, S_COST) + t] = k == a.fixed ? 0.0 : (a.ground_p ? rp * rp : 0.0) + (a.ground_q ? rq * rq : 0.0);
        double n2 = 0.0, s2 = 0.0;
        if (k != a.fixed)
            for (int q = 0; q < 6; ++q) { n2 += yk[q] * yk[q]; const double d = a.x[6 * k + q] - yk[q]; s2 += d * d; }
        a.part[part_offset(a, S_YNORM) + k] = n2;
        a.part[part_offset(a, S_STEP) + k] = s2;
    }
}

template <int KID> LV_HD void pg_thread(const Args& a, int t) {
    if (KID == K_COLUMNS) body_columns(a, t);
    else if (KID == K_ASSEMBLE) body_assemble(a, t);
    else if (KID == K_SCALE) body_scale(a, t);
    else if (KID == K_TRISOLVE) body_trisolve(a, t);
    else if (KID == K_CAPACITANCE) body_capacitance(a, t);
    else if (KID == K_COMBINE) body_combine(a, t);
    else if (KID == K_MODEL) body_model(a, t);
    else if (KID == K_COST) body_cost(a, t);
}
LV_HD int kernel_threads(const Args& a, int kid) {
    switch (kid) {
        case K_COLUMNS: return 13 * a.E + 7 * a.K;
        case K_ASSEMBLE: case K_COMBINE: return a.K;
        case K_SCALE: return 6 * a.K;
        case K_TRISOLVE: return a.ncol;
        case K_CAPACITANCE: return 6 * a.L * a.ncol;
        default: return a.E + a.K;
    }
}

#if defined(__CUDACC__)
template <int KID> __global__ void __launch_bounds__(128) pg_kernel(Args a, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) pg_thread<KID>(a, t);
}
__global__ void __launch_bounds__(64) pg_factor_kernel(Args a) {
    __shared__ FactorTile T;
    const int t = threadIdx.x;
    const bool work = t < FACTOR_THREADS;
    if (work) factor_stage(T, t, factor_fetch(a, 0, t));
    __syncthreads();
    for (int k = 0; k < a.K; ++k) {
        FactorNext nx;
        const bool more = work && k + 1 < a.K;
        if (more) nx = factor_fetch(a, k + 1, t);   // in flight while the phases of step k run
#pragma unroll
        for (int phase = 0; phase < FACTOR_PHASES; ++phase) {
            if (work) factor_phase(a, T, k, phase, t);
            if (phase == FACTOR_PHASES - 1 && more) factor_stage(T, t, nx);
            __syncthreads();
        }
    }
    if (t == 0) a.flags[0] = T.ok ? 0 : 1;
}
__global__ void __launch_bounds__(256) pg_dense_kernel(Args a) {
    const int n = 6 * a.L, tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) a.flags[1] = 0;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        dense_phase_pivot(a, j, tid);
        __syncthreads();
        dense_phase_column(a, j, tid, nt);
        __syncthreads();
        dense_phase_update(a, j, tid, nt);
        __syncthreads();
    }
    dense_phase_solve(a, tid);
}
// one CTA per scalar: sum (max for the gradient norm) of its partials, fixed order
__global__ void __launch_bounds__(256) pg_reduce_kernel(Args a) {
    __shared__ double sh[256];
    const int q = blockIdx.x, tid = threadIdx.x;
    const double* p = a.part + part_offset(a, q);
    const int cnt = part_count(a, q);
    const bool is_max = q == S_GRAD;
    double v = 0.0;
    for (int i = tid; i < cnt; i += 256) v = is_max ? fmax(v, p[i]) : v + p[i];
    sh[tid] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) sh[tid] = is_max ? fmax(sh[tid], sh[tid + s]) : sh[tid] + sh[tid + s];
        __syncthreads();
    }
    if (tid == 0) a.scal[q] = sh[0];
}
#endif

// ------------------------------------------------------------------ the minimiser (host control flow)
// `Launcher` runs the kernels: bool run(int kernel_id, const Args&), bool factor(const Args&), bool dense(const Args&), bool reduce(const Args&),
// bool partitioned(const Args&) (only called when a.P > 1, pose_graph_segments.cuh),
// bool read(const Args&, double* scal /*[S_COUNT]*/, int32_t* flags /*[2]*/) (the read is the synchronisation point).
// Returns false on a launcher (CUDA) error.  The optimised poses end up in a.x (x and xc are swapped on acceptance).
template <class Launcher>
inline bool pg_minimize(Launcher& Lr, Args& a, const Options& opt, lvio2d_summary* out) {
    lvio2d_summary S;
    S.iterations = 0; S.termination = LVIO2D_TERM_NO_CONVERGENCE; S.num_successful_steps = 0; S.num_unsuccessful_steps = 0;
    S.initial_cost = 0.0; S.final_cost = 0.0; S.final_radius = 0.0; S.reserved = 0.0;
    double scal[S_COUNT];
    int32_t flags[2];
    a.min_lm = opt.min_lm_diagonal; a.max_lm = opt.max_lm_diagonal;
    double radius = opt.initial_radius, decrease_factor = 2.0;
    a.radius = radius;
    // cost, |x| and the linearisation at the start point
    a.y = a.x;
    if (!Lr.run(K_COST, a) || !Lr.run(K_COLUMNS, a) || !Lr.run(K_ASSEMBLE, a) || !Lr.run(K_SCALE, a) || !Lr.reduce(a) || !Lr.read(a, scal, flags)) return false;
    double x_cost = 0.5 * scal[S_COST], x_norm = sqrt(scal[S_YNORM]), grad_max = scal[S_GRAD];
    S.initial_cost = x_cost;
    const int free_poses = a.K - ((a.fixed >= 0 && a.fixed < a.K) ? 1 : 0);
    if (free_poses == 0) {
        S.final_cost = x_cost; S.final_radius = radius; S.termination = LVIO2D_TERM_CONVERGENCE_GRADIENT;
        *out = S;
        return true;
    }
    int num_consecutive_invalid = 0, iteration = 0;
    bool last_successful = true;
    for (;;) {
        if (iteration >= opt.max_iters) { S.termination = LVIO2D_TERM_NO_CONVERGENCE; break; }
        if (last_successful && grad_max <= opt.gradient_tolerance) { S.termination = LVIO2D_TERM_CONVERGENCE_GRADIENT; break; }
        if (radius < opt.min_radius) { S.termination = LVIO2D_TERM_CONVERGENCE_RADIUS; break; }
        ++iteration;
        // ComputeTrustRegionStep + candidate evaluation, one round trip
        a.radius = radius;
        a.y = a.xc;
        bool ok = a.P > 1 ? Lr.partitioned(a) : (Lr.factor(a) && Lr.run(K_TRISOLVE, a));
        if (ok && a.L > 0) ok = Lr.run(K_CAPACITANCE, a) && Lr.dense(a);
        ok = ok && Lr.run(K_COMBINE, a) && Lr.run(K_MODEL, a) && Lr.run(K_COST, a) && Lr.reduce(a) && Lr.read(a, scal, flags);
        if (!ok) return false;
        bool valid = flags[0] == 0 && (a.L == 0 || flags[1] == 0);
        // model_cost_change = -step^T g - 1/2 step^T H step (the same number in scaled and unscaled columns)
        const double model_cost_change = -scal[S_DG] - 0.5 * scal[S_MODEL];
        if (valid) valid = std::isfinite(model_cost_change) && std::isfinite(scal[S_STEP]) && model_cost_change > 0.0;
        if (!valid) {
            ++num_consecutive_invalid;
            last_successful = false;
            ++S.num_unsuccessful_steps;
            if (num_consecutive_invalid >= opt.max_consecutive_invalid) { S.termination = LVIO2D_TERM_FAILURE; break; }
            radius = radius / decrease_factor; decrease_factor *= 2.0;
            continue;
        }
        num_consecutive_invalid = 0;
        double candidate_cost = 0.5 * scal[S_COST];
        if (!std::isfinite(candidate_cost)) candidate_cost = 1.7976931348623157e308;
        const double step_norm = sqrt(scal[S_STEP]);
        if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { S.termination = LVIO2D_TERM_CONVERGENCE_PARAMETER; break; }
        const double cost_change = x_cost - candidate_cost;
        if (fabs(cost_change) <= opt.function_tolerance * x_cost) { S.termination = LVIO2D_TERM_CONVERGENCE_FUNCTION; break; }
        const double relative_decrease = cost_change / model_cost_change;
        if (relative_decrease > opt.min_relative_decrease) {
            double* tmp = a.x; a.x = a.xc; a.xc = tmp;
            x_norm = sqrt(scal[S_YNORM]);
            x_cost = candidate_cost;
            if (!Lr.run(K_COLUMNS, a) || !Lr.run(K_ASSEMBLE, a) || !Lr.reduce(a) || !Lr.read(a, scal, flags)) return false;
            grad_max = scal[S_GRAD];
            const double t = 2.0 * relative_decrease - 1.0;
            radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
            radius = fmin(opt.max_radius, radius);
            decrease_factor = 2.0;
            last_successful = true;
            ++S.num_successful_steps;
        } else {
            radius = radius / decrease_factor; decrease_factor *= 2.0;
            last_successful = false;
            ++S.num_unsuccessful_steps;
        }
    }
    S.iterations = iteration;
    S.final_cost = x_cost;
    S.final_radius = radius;
    *out = S;
    return true;
}

// host-side preparation shared by the product entry point and the CPU check: band / loop classification, the incidence
// lists, and the carving of the two arenas (one of int32, one of double) every Args pointer lives in.
// int arena:    [edge_index 2E][edge_band E][loop_edge L][inc_off K+1][inc 2E][flags 2]  (+ [node_seg K][segflag P] when P > 1)
// double arena: [edge_tf 12E][edge_weight E][x 6K][xc 6K][EJ 78E][GJ 14K][g 6K][Hd 6K][D 36K][O 36K][scale 6K][Sinv 36K][M 36K]
//               [Z 6K*ncol][Cm 6L*ncol][w 6L][delta 6K][part][scal S_COUNT]  (+ [Zx 6K(ncol+12)][Rd Ro RSinv RM 36(P-1)][Zr 6(P-1)ncol])
// Returns false on an invalid graph (index out of range, self edge).  `ints` receives the int arena's host image.
template <class VecI>
inline bool pg_topology(int K, int E, const int32_t* edge_index, VecI& ints, int* n_loops, int P = 0) {
    VecI band(E, 0), loops, inc_off(K + 1, 0), inc(2 * (size_t)E, 0);
    for (int e = 0; e < E; ++e) {
        const int i = edge_index[2 * e], j = edge_index[2 * e + 1];
        if (i < 0 || j < 0 || i >= K || j >= K || i == j) return false;
        const int d = i > j ? i - j : j - i;
        band[e] = d == 1 ? 1 : 0;
        if (d != 1) loops.push_back(e);
        ++inc_off[i + 1];
        ++inc_off[j + 1];
    }
    for (int k = 0; k < K; ++k) inc_off[k + 1] += inc_off[k];
    VecI fill(inc_off.begin(), inc_off.end() - 1);
    for (int e = 0; e < E; ++e) {
        inc[fill[edge_index[2 * e]]++] = 2 * e;
        inc[fill[edge_index[2 * e + 1]]++] = 2 * e + 1;
    }
    ints.clear();
    ints.insert(ints.end(), edge_index, edge_index + 2 * (size_t)E);
    ints.insert(ints.end(), band.begin(), band.end());
    ints.insert(ints.end(), loops.begin(), loops.end());
    ints.insert(ints.end(), inc_off.begin(), inc_off.end());
    ints.insert(ints.end(), inc.begin(), inc.end());
    ints.push_back(0);
    ints.push_back(0);
    if (P > 1) {   // node_seg [K] + segflag [P]; separator i sits at key frame (i + 1) K / P
        VecI node_seg(K, 0);
        int c = 0;
        for (int k = 0; k < K; ++k) {
            const int next_sep = c < P - 1 ? (int)(((long long)(c + 1) * K) / P) : K;
            if (k == next_sep) { node_seg[k] = -(c + 1); ++c; }
            else node_seg[k] = c;
        }
        ints.insert(ints.end(), node_seg.begin(), node_seg.end());
        ints.insert(ints.end(), (size_t)P, 0);
    }
    *n_loops = (int)loops.size();
    return true;
}
// a.K, a.E, a.L, a.P must be set; assigns ncol and every pointer.  Returns the number of doubles the double arena needs.
inline size_t pg_bind(Args& a, int32_t* ints, double* dbl) {
    a.ncol = 1 + 6 * a.L;
    const size_t K = (size_t)a.K, E = (size_t)a.E, L = (size_t)a.L, nc = (size_t)a.ncol;
    size_t ioff = 0;
    auto itake = [&](size_t n) { int32_t* p = ints ? ints + ioff : nullptr; ioff += n; return p; };
    a.edge_index = itake(2 * E);
    a.edge_band = itake(E);
    a.loop_edge = itake(L);
    a.inc_off = itake(K + 1);
    a.inc = itake(2 * E);
    a.flags = itake(2);
    a.node_seg = a.P > 1 ? itake(K) : nullptr;
    a.segflag = a.P > 1 ? itake((size_t)a.P) : nullptr;
    size_t off = 0;
    auto take = [&](size_t n) { double* p = dbl ? dbl + off : nullptr; off += n; return p; };
    a.edge_tf = take(12 * E);
    a.edge_weight = take(E);
    a.x = take(6 * K);
    a.xc = take(6 * K);
    a.EJ = take(78 * E);
    a.GJ = take(14 * K);
    a.g = take(6 * K);
    a.Hd = take(6 * K);
    a.D = take(36 * K);
    a.O = take(36 * K);
    a.scale = take(6 * K);
    a.Sinv = take(36 * K);
    a.M = take(36 * K);
    a.Z = take(6 * K * nc);
    a.Cm = take(6 * L * nc);
    a.w = take(6 * L);
    a.delta = take(6 * K);
    a.part = take((size_t)part_total(a));
    a.scal = take(S_COUNT);
    if (a.P > 1) {
        const size_t R = (size_t)a.P - 1;
        a.Zx = take(6 * K * (nc + 12));
        a.Rd = take(36 * R); a.Ro = take(36 * R); a.RSinv = take(36 * R); a.RM = take(36 * R);
        a.Zr = take(6 * R * nc);
    } else {
        a.Zx = a.Rd = a.Ro = a.RSinv = a.RM = a.Zr = nullptr;
    }
    a.y = a.x;
    return off;
}
// the segment count that minimises the sequential depth 3 K / P + 2 P (LVIO2D_PG_SEGMENTS=auto); below 64 key frames: none
inline int pg_auto_segments(int K) { return K < 64 ? 0 : (int)std::lround(std::sqrt(1.5 * K)); }
// the largest useful number of segments for K key frames (0: use the plain path): every segment keeps >= 3 interior key frames
inline int pg_segments(int K, int requested) {
    int P = requested;
    if (P > K / 4) P = K / 4;
    return P >= 2 ? P : 0;
}

}  // namespace pg
}  // namespace lv
