// pose_graph_segments.cuh — the block-tridiagonal solve of pose_graph.cuh with its sequential chains cut into P segments.
//
// STATUS: the default path since round 2 (P = sqrt(1.5 K) segments, segment solves staged in shared memory;
// LVIO2D_PG_SEGMENTS=0 selects the single-chain path of pose_graph.cuh, LVIO2D_PG_STAGE=0 the unstaged segment solves).
// Written from the launch list in profiles/r1_pose_graph.md (the two sequential chains over the key frames were 94 % of
// an LM iteration); verified on the CPU by the thread-by-thread host run against the oracle
// (tests/test_pose_graph_host.py) and on a B200 (tests/test_zz_gpu_pose_graph.py, compute-sanitizer clean,
// profiles/r2_pose_graph.md): 0.41 / 0.55 / 1.21 ms per LM iteration at 200 / 1000 / 4000 key frames against
// 1.11 / 4.67 / 18.7 ms on the single chain and 0.91 / 4.54 / 32.3 ms for the same algorithm on one host core.
//
// Substructuring of T x = b (T block-tridiagonal SPD, 6x6 blocks, K key frames): P - 1 separator key frames
// s_i = (i + 1) K / P split the chain into P segments of interior key frames.
//   (1) every segment (one CTA each, concurrently): factorise its interior block-tridiagonal matrix T_c;
//   (2) every segment: T_c^-1 [ b_c | E_left | E_right ] — the 1 + 6L right-hand sides plus the two 6-column "spikes"
//       that couple the segment to its separators (E_left = O_{lo-1} at the first row, E_right = O_hi^T at the last);
//   (3) the Schur complement on the separators is again block-tridiagonal (P - 1 blocks):
//         R_i      = T_ss - O_{s-1} W_i[s-1] - O_s^T V_{i+1}[s+1]          (V / W: left / right spike solutions)
//         R_{i,i-1} = - O_{s-1} V_i[s-1]
//         r_i      = b_s - O_{s-1} z_i[s-1] - O_s^T z_{i+1}[s+1]
//       factorised and solved by one CTA (P - 1 sequential steps);
//   (4) interior back-substitution x = z - V x_left - W x_right, all key frames concurrently.
// Sequential depth: 3 K / P + 2 P block steps instead of 3 K.  Same arithmetic class as the plain path (exact; the
// elimination order differs, results agree to rounding).
#pragma once
#include "pose_graph.cuh"

namespace lv {
namespace pg {

enum SegKernel { KS_TRISOLVE = 100, KS_REDUCED_BLOCKS, KS_REDUCED_RHS, KS_REDUCED_TRISOLVE, KS_BACKSUB };

LV_HD int sep_node(const Args& a, int i) { return (int)(((long long)(i + 1) * a.K) / a.P); }
LV_HD int seg_lo(const Args& a, int c) { return c == 0 ? 0 : sep_node(a, c - 1) + 1; }
LV_HD int seg_hi(const Args& a, int c) { return c == a.P - 1 ? a.K - 1 : sep_node(a, c) - 1; }
LV_HD int ncol_x(const Args& a) { return a.ncol + 12; }

// one block-tridiagonal system over consecutive entries lo..hi of node-indexed arrays; O[i] = block (i + 1, i)
struct Chain {
    const double* D;
    const double* O;
    const double* Hd;   // nullptr: D already carries the LM diagonal (reduced system)
    const double* Sc;
    double* Sinv;
    double* M;
    int lo, hi, fixed;
};
LV_HD Chain segment_chain(const Args& a, int c) {
    Chain ch;
    ch.D = a.D; ch.O = a.O; ch.Hd = a.Hd; ch.Sc = a.scale; ch.Sinv = a.Sinv; ch.M = a.M;
    ch.lo = seg_lo(a, c); ch.hi = seg_hi(a, c); ch.fixed = a.fixed;
    return ch;
}
LV_HD Chain reduced_chain(const Args& a) {
    Chain ch;
    ch.D = a.Rd; ch.O = a.Ro; ch.Hd = nullptr; ch.Sc = nullptr; ch.Sinv = a.RSinv; ch.M = a.RM;
    ch.lo = 0; ch.hi = a.P - 2; ch.fixed = -1;
    return ch;
}
// the LM term of key frame k, column r, in unscaled columns (as factor_phase of pose_graph.cuh)
LV_HD double lm_term(const Args& a, double hd, double sc) {
    const double hs = hd * sc * sc;
    const double d = fmin(fmax(hs, a.min_lm), a.max_lm);
    return d / a.radius / (sc * sc);
}

// ---- factorisation of a chain: the phases of pose_graph.cuh's factor_phase with a range and optional damping
LV_HD FactorNext chain_fetch(const Chain& ch, int k, int t) {
    FactorNext n;
    n.D = ch.D[(size_t)k * 36 + t];
    n.O = k > ch.lo ? ch.O[(size_t)(k - 1) * 36 + t] : 0.0;
    n.Hd = (ch.Hd && t < 6) ? ch.Hd[6 * k + t] : 0.0;
    n.Sc = (ch.Sc && t < 6) ? ch.Sc[6 * k + t] : 1.0;
    return n;
}
LV_HD void chain_phase(const Args& a, const Chain& ch, FactorTile& T, int k, int phase, int t) {
    const int r = t / 6, c = t % 6;
    if (phase == 0) {
        double v = T.nD[t];
        if (ch.Hd && r == c && k != ch.fixed) v += lm_term(a, T.nHd[r], T.nSc[r]);
        T.S[0][t] = v;
        T.B[t] = T.nO[t];
        T.Inv[0][t] = (r == c) ? 1.0 : 0.0;
        if (t == 0 && k == ch.lo) T.ok = 1;
    } else if (phase == 1) {
        double m = 0.0;
        if (k > ch.lo)
            for (int i = 0; i < 6; ++i) m += T.B[r * 6 + i] * T.Sp[i * 6 + c];
        T.Mk[t] = m;
        ch.M[(size_t)k * 36 + t] = m;
    } else if (phase == 2) {
        if (k > ch.lo) {
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
        if (j == 5) {
            T.Sp[t] = ni;
            ch.Sinv[(size_t)k * 36 + t] = ni;
        }
    }
}

// ---- right-hand sides
struct ColumnInfo { int e, rho, ie, je; };
LV_HD ColumnInfo column_info(const Args& a, int col) {
    ColumnInfo ci;
    ci.e = -1; ci.rho = 0; ci.ie = -1; ci.je = -1;
    if (col > 0 && col < a.ncol) {
        ci.e = a.loop_edge[(col - 1) / 6];
        ci.rho = (col - 1) % 6;
        ci.ie = a.edge_index[2 * ci.e];
        ci.je = a.edge_index[2 * ci.e + 1];
    }
    return ci;
}
// entry q of key frame k of global right-hand side `col` (0: -g, 1 + 6l + rho: loop edge l, residual row rho)
LV_HD double global_rhs(const Args& a, const ColumnInfo& ci, int col, int k, int q) {
    if (col == 0) return -a.g[6 * k + q];
    const double* J = a.EJ + (size_t)ci.e * 78;
    return (k == ci.ie) ? J[q * 6 + ci.rho] : ((k == ci.je) ? J[36 + q * 6 + ci.rho] : 0.0);
}
// what a thread needs to know about its right-hand side column
struct ColumnCtx { ColumnInfo ci; bool spike, left; int sj; };
LV_HD ColumnCtx column_ctx(const Args& a, int col, int kind) {
    ColumnCtx cc;
    cc.ci = column_info(a, kind == 0 ? col : 0);
    cc.spike = kind == 0 && col >= a.ncol;
    cc.left = cc.spike && (col - a.ncol) < 6;
    cc.sj = cc.spike ? (col - a.ncol) % 6 : 0;
    return cc;
}
// right-hand side block of key frame k.  kind 0: global columns and (col >= ncol) the two spikes of the segment;
// kind 1: the right-hand side is in Z already.
LV_HD void chain_rhs(const Args& a, const Chain& ch, const double* Z, int ncz, int col, int kind, const ColumnCtx& cc, int k, double* b) {
    for (int q = 0; q < 6; ++q) {
        if (kind == 1) b[q] = Z[(size_t)(6 * k + q) * ncz + col];
        else if (!cc.spike) b[q] = global_rhs(a, cc.ci, col, k, q);
        else if (cc.left) b[q] = (k == ch.lo && ch.lo > 0) ? a.O[(size_t)(ch.lo - 1) * 36 + q * 6 + cc.sj] : 0.0;    // O_{lo-1}[:, sj]
        else b[q] = (k == ch.hi && ch.hi < a.K - 1) ? a.O[(size_t)ch.hi * 36 + cc.sj * 6 + q] : 0.0;                 // O_hi^T[:, sj]
    }
}
// n forward steps from key frame k0; Mblk = M[k0 .. k0 + n - 1] (global or staged in shared memory); yp carries y_{k0-1}
LV_HD void forward_steps(const Args& a, const Chain& ch, double* Z, int ncz, int col, int kind, const ColumnCtx& cc, int k0, int n,
                         const double* Mblk, double* yp) {
    double b[6];
    for (int s = 0; s < n; ++s) {
        const int k = k0 + s;
        chain_rhs(a, ch, Z, ncz, col, kind, cc, k, b);
        if (k > ch.lo) {
            const double* Mk = Mblk + (size_t)s * 36;
            for (int q = 0; q < 6; ++q) {
                double t = 0.0;
                for (int i = 0; i < 6; ++i) t += Mk[q * 6 + i] * yp[i];
                b[q] -= t;
            }
        }
        for (int q = 0; q < 6; ++q) { yp[q] = b[q]; Z[(size_t)(6 * k + q) * ncz + col] = b[q]; }
    }
}
// n backward steps from key frame k0 + n - 1 down to k0; Sblk = Sinv[k0 ..], Mnext = M[k0 + 1 ..]; xn carries x_{k0+n}
LV_HD void backward_steps(const Chain& ch, double* Z, int ncz, int col, int k0, int n, const double* Sblk, const double* Mnext, double* xn) {
    double y[6], b[6];
    for (int s = n - 1; s >= 0; --s) {
        const int k = k0 + s;
        const double* Si = Sblk + (size_t)s * 36;
        for (int q = 0; q < 6; ++q) y[q] = Z[(size_t)(6 * k + q) * ncz + col];
        for (int q = 0; q < 6; ++q) {
            double t = 0.0;
            for (int i = 0; i < 6; ++i) t += Si[q * 6 + i] * y[i];
            b[q] = t;
        }
        if (k < ch.hi) {
            const double* Mn = Mnext + (size_t)s * 36;
            for (int q = 0; q < 6; ++q) {
                double t = 0.0;
                for (int i = 0; i < 6; ++i) t += Mn[i * 6 + q] * xn[i];
                b[q] -= t;
            }
        }
        for (int q = 0; q < 6; ++q) { xn[q] = b[q]; Z[(size_t)(6 * k + q) * ncz + col] = b[q]; }
    }
}
// forward / backward block substitution of one right-hand side over a whole chain, blocks straight from global memory
LV_HD void chain_trisolve(const Args& a, const Chain& ch, double* Z, int ncz, int col, int kind) {
    const ColumnCtx cc = column_ctx(a, col, kind);
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    forward_steps(a, ch, Z, ncz, col, kind, cc, ch.lo, ch.hi - ch.lo + 1, ch.M + (size_t)ch.lo * 36, v);
    for (int q = 0; q < 6; ++q) v[q] = 0.0;
    backward_steps(ch, Z, ncz, col, ch.lo, ch.hi - ch.lo + 1, ch.Sinv + (size_t)ch.lo * 36, ch.M + (size_t)(ch.lo + 1) * 36, v);
}
// the same with the blocks of STAGE_STEPS steps staged in shared memory by the CTA (a.stage != 0): one L2 round trip per
// STAGE_STEPS steps instead of 4.5 per step (profiles/r1_pose_graph.md).  stage_copy is the cooperative load of a phase.
enum { STAGE_STEPS = 16 };
struct StageTile { double M[STAGE_STEPS * 36], S[STAGE_STEPS * 36]; };
LV_HD void stage_copy(double* dst, const double* src, int count, int tid, int nt) {
    for (int i = tid; i < count; i += nt) dst[i] = src[i];
}
// the blocks of the backward chunk [k0, k0 + n): Sinv[k0 ..] and M[k0 + 1 ..] (the last one only if it exists)
LV_HD void stage_backward(const Chain& ch, StageTile& T, int k0, int n, int tid, int nt) {
    stage_copy(T.S, ch.Sinv + (size_t)k0 * 36, n * 36, tid, nt);
    const int nm = (k0 + n <= ch.hi) ? n : n - 1;
    stage_copy(T.M, ch.M + (size_t)(k0 + 1) * 36, nm * 36, tid, nt);
}

// ---- kernel bodies
// KS_TRISOLVE, n = P * (ncol + 12): segment c, column col
LV_HD void body_seg_trisolve(const Args& a, int t) {
    const int ncx = ncol_x(a);
    chain_trisolve(a, segment_chain(a, t / ncx), a.Zx, ncx, t % ncx, 0);
}
// KS_REDUCED_BLOCKS, n = (P - 1) * 36: entry (r, c) of R_i and of R_{i, i-1}
LV_HD void body_reduced_blocks(const Args& a, int t) {
    const int i = t / 36, r = (t % 36) / 6, c = t % 6;
    const int s = sep_node(a, i), ncx = ncol_x(a);
    const double* Ol = a.O + (size_t)(s - 1) * 36;   // block (s, s-1)
    const double* Or = a.O + (size_t)s * 36;         // block (s+1, s)
    double v = a.D[(size_t)s * 36 + r * 6 + c];
    if (r == c && s != a.fixed) v += lm_term(a, a.Hd[6 * s + r], a.scale[6 * s + r]);
    double off = 0.0;
    for (int m = 0; m < 6; ++m) {
        const double W = a.Zx[(size_t)(6 * (s - 1) + m) * ncx + a.ncol + 6 + c];   // right spike of segment i at its last key frame
        const double V = a.Zx[(size_t)(6 * (s + 1) + m) * ncx + a.ncol + c];       // left spike of segment i+1 at its first key frame
        v -= Ol[r * 6 + m] * W + Or[m * 6 + r] * V;
        off -= Ol[r * 6 + m] * a.Zx[(size_t)(6 * (s - 1) + m) * ncx + a.ncol + c]; // left spike of segment i at its last key frame
    }
    a.Rd[(size_t)i * 36 + r * 6 + c] = v;
    if (i > 0) a.Ro[(size_t)(i - 1) * 36 + r * 6 + c] = off;
}
// KS_REDUCED_RHS, n = (P - 1) * 6 * ncol
LV_HD void body_reduced_rhs(const Args& a, int t) {
    const int col = t % a.ncol, q = (t / a.ncol) % 6, i = t / (6 * a.ncol);
    const int s = sep_node(a, i), ncx = ncol_x(a);
    const ColumnInfo ci = column_info(a, col);
    const double* Ol = a.O + (size_t)(s - 1) * 36;
    const double* Or = a.O + (size_t)s * 36;
    double v = global_rhs(a, ci, col, s, q);
    for (int m = 0; m < 6; ++m)
        v -= Ol[q * 6 + m] * a.Zx[(size_t)(6 * (s - 1) + m) * ncx + col] + Or[m * 6 + q] * a.Zx[(size_t)(6 * (s + 1) + m) * ncx + col];
    a.Zr[(size_t)(6 * i + q) * a.ncol + col] = v;
}
// KS_REDUCED_TRISOLVE, n = ncol
LV_HD void body_reduced_trisolve(const Args& a, int col) { chain_trisolve(a, reduced_chain(a), a.Zr, a.ncol, col, 1); }
// KS_BACKSUB, n = K * ncol: the solution in the layout the plain path leaves it in (a.Z)
LV_HD void body_backsub(const Args& a, int t) {
    const int k = t / a.ncol, col = t % a.ncol, ncx = ncol_x(a);
    const int tag = a.node_seg[k];
    if (tag < 0) {
        const int i = -tag - 1;
        for (int q = 0; q < 6; ++q) a.Z[(size_t)(6 * k + q) * a.ncol + col] = a.Zr[(size_t)(6 * i + q) * a.ncol + col];
        return;
    }
    double xl[6], xr[6];
    for (int j = 0; j < 6; ++j) {
        xl[j] = tag > 0 ? a.Zr[(size_t)(6 * (tag - 1) + j) * a.ncol + col] : 0.0;
        xr[j] = tag < a.P - 1 ? a.Zr[(size_t)(6 * tag + j) * a.ncol + col] : 0.0;
    }
    for (int q = 0; q < 6; ++q) {
        const double* z = a.Zx + (size_t)(6 * k + q) * ncx;
        double s = z[col];
        for (int j = 0; j < 6; ++j) s -= z[a.ncol + j] * xl[j] + z[a.ncol + 6 + j] * xr[j];
        a.Z[(size_t)(6 * k + q) * a.ncol + col] = s;
    }
}
template <int KID> LV_HD void seg_thread(const Args& a, int t) {
    if (KID == KS_TRISOLVE) body_seg_trisolve(a, t);
    else if (KID == KS_REDUCED_BLOCKS) body_reduced_blocks(a, t);
    else if (KID == KS_REDUCED_RHS) body_reduced_rhs(a, t);
    else if (KID == KS_REDUCED_TRISOLVE) body_reduced_trisolve(a, t);
    else if (KID == KS_BACKSUB) body_backsub(a, t);
}
LV_HD int seg_kernel_threads(const Args& a, int kid) {
    switch (kid) {
        case KS_TRISOLVE: return a.P * ncol_x(a);
        case KS_REDUCED_BLOCKS: return (a.P - 1) * 36;
        case KS_REDUCED_RHS: return (a.P - 1) * 6 * a.ncol;
        case KS_REDUCED_TRISOLVE: return a.ncol;
        default: return a.K * a.ncol;
    }
}
// a whole chain by the 36 working threads of one CTA; `sync` is __syncthreads on the device, nothing on the host (where
// the caller runs the threads of a phase one after the other)
#if defined(__CUDACC__)
template <int KID> __global__ void __launch_bounds__(128) pg_seg_kernel(Args a, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) seg_thread<KID>(a, t);
}
// KS_TRISOLVE with staging: CTA (c, g) walks segment c for the columns [128 g, 128 g + 128)
__global__ void __launch_bounds__(128) pg_seg_trisolve_staged_kernel(Args a) {
    __shared__ StageTile T;
    const int tid = threadIdx.x, nt = blockDim.x, ncx = ncol_x(a);
    const Chain ch = segment_chain(a, blockIdx.x);
    const int col = blockIdx.y * nt + tid;
    const bool active = col < ncx;
    const ColumnCtx cc = column_ctx(a, active ? col : 0, 0);
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int k0 = ch.lo; k0 <= ch.hi; k0 += STAGE_STEPS) {
        const int n = min((int)STAGE_STEPS, ch.hi - k0 + 1);
        stage_copy(T.M, ch.M + (size_t)k0 * 36, n * 36, tid, nt);
        __syncthreads();
        if (active) forward_steps(a, ch, a.Zx, ncx, col, 0, cc, k0, n, T.M, v);
        __syncthreads();
    }
    for (int q = 0; q < 6; ++q) v[q] = 0.0;
    for (int k1 = ch.hi; k1 >= ch.lo; k1 -= STAGE_STEPS) {
        const int k0 = max(ch.lo, k1 - (int)STAGE_STEPS + 1), n = k1 - k0 + 1;
        stage_backward(ch, T, k0, n, tid, nt);
        __syncthreads();
        if (active) backward_steps(ch, a.Zx, ncx, col, k0, n, T.S, T.M, v);
        __syncthreads();
    }
}
// reduced == 0: CTA c factorises segment c (flag -> segflag[c]); reduced == 1: one CTA factorises the separator system
// and folds every flag into flags[0]
__global__ void __launch_bounds__(64) pg_chain_factor_kernel(Args a, int reduced) {
    __shared__ FactorTile T;
    const int t = threadIdx.x;
    const bool work = t < FACTOR_THREADS;
    const Chain ch = reduced ? reduced_chain(a) : segment_chain(a, blockIdx.x);
    if (work) factor_stage(T, t, chain_fetch(ch, ch.lo, t));
    __syncthreads();
    for (int k = ch.lo; k <= ch.hi; ++k) {
        FactorNext nx;
        const bool more = work && k + 1 <= ch.hi;
        if (more) nx = chain_fetch(ch, k + 1, t);
#pragma unroll
        for (int phase = 0; phase < FACTOR_PHASES; ++phase) {
            if (work) chain_phase(a, ch, T, k, phase, t);
            if (phase == FACTOR_PHASES - 1 && more) factor_stage(T, t, nx);
            __syncthreads();
        }
    }
    if (t == 0) {
        if (!reduced) {
            a.segflag[blockIdx.x] = T.ok ? 0 : 1;
        } else {
            int bad = T.ok ? 0 : 1;
            for (int c = 0; c < a.P; ++c) bad |= a.segflag[c];
            a.flags[0] = bad;
        }
    }
}
#endif

// the launch sequence of one partitioned solve (what Launcher::partitioned runs): Lr.chain_factor(a, reduced),
// Lr.seg(kernel id, a)
template <class Launcher> inline bool pg_partitioned_solve(Launcher& Lr, const Args& a) {
    return Lr.chain_factor(a, 0) && Lr.seg(KS_TRISOLVE, a) && Lr.seg(KS_REDUCED_BLOCKS, a) && Lr.seg(KS_REDUCED_RHS, a) && Lr.chain_factor(a, 1) &&
           Lr.seg(KS_REDUCED_TRISOLVE, a) && Lr.seg(KS_BACKSUB, a);
}

}  // namespace pg
}  // namespace lv
