// Test-only host build of the product's pose-graph header (2dliw-slam_b200/csrc/pose_graph.cuh): the kernel bodies are
// __host__ __device__ and free of intra-block communication, so the CPU test-suite runs every "kernel" thread by thread
// (and the two cooperative ones phase by phase) under the product's own minimiser loop and checks the result against the
// oracle without a GPU.  This is NOT a CPU fallback of the product: it is compiled by tests/test_pose_graph_host.py only.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../2dliw-slam_b200/csrc/pose_graph_segments.cuh"
using namespace lv;
using namespace lv::pg;

namespace {
struct HostLauncher {
    int launches = 0;
    template <int KID> static void loop(const Args& a) {
        const int n = kernel_threads(a, KID);
        for (int t = 0; t < n; ++t) pg_thread<KID>(a, t);
    }
    bool run(int kid, const Args& a) {
        ++launches;
        switch (kid) {
            case K_COLUMNS: loop<K_COLUMNS>(a); break;
            case K_ASSEMBLE: loop<K_ASSEMBLE>(a); break;
            case K_SCALE: loop<K_SCALE>(a); break;
            case K_TRISOLVE: loop<K_TRISOLVE>(a); break;
            case K_CAPACITANCE: loop<K_CAPACITANCE>(a); break;
            case K_COMBINE: loop<K_COMBINE>(a); break;
            case K_MODEL: loop<K_MODEL>(a); break;
            case K_COST: loop<K_COST>(a); break;
            default: return false;
        }
        return true;
    }
    bool factor(const Args& a) {
        ++launches;
        FactorTile T;
        for (int t = 0; t < FACTOR_THREADS; ++t) factor_stage(T, t, factor_fetch(a, 0, t));
        for (int k = 0; k < a.K; ++k) {
            for (int phase = 0; phase < FACTOR_PHASES; ++phase)
                for (int t = 0; t < FACTOR_THREADS; ++t) factor_phase(a, T, k, phase, t);
            if (k + 1 < a.K)
                for (int t = 0; t < FACTOR_THREADS; ++t) factor_stage(T, t, factor_fetch(a, k + 1, t));
        }
        a.flags[0] = T.ok ? 0 : 1;
        return true;
    }
    // the partitioned solve (pose_graph_segments.cuh): segment / reduced factorisations phase by phase, the rest thread by thread
    bool chain_factor(const Args& a, int reduced) {
        ++launches;
        int bad = 0;
        const int ctas = reduced ? 1 : a.P;
        for (int cta = 0; cta < ctas; ++cta) {
            const Chain ch = reduced ? reduced_chain(a) : segment_chain(a, cta);
            FactorTile T;
            for (int t = 0; t < FACTOR_THREADS; ++t) factor_stage(T, t, chain_fetch(ch, ch.lo, t));
            for (int k = ch.lo; k <= ch.hi; ++k) {
                for (int phase = 0; phase < FACTOR_PHASES; ++phase)
                    for (int t = 0; t < FACTOR_THREADS; ++t) chain_phase(a, ch, T, k, phase, t);
                if (k + 1 <= ch.hi)
                    for (int t = 0; t < FACTOR_THREADS; ++t) factor_stage(T, t, chain_fetch(ch, k + 1, t));
            }
            if (!reduced) a.segflag[cta] = T.ok ? 0 : 1;
            else bad = T.ok ? 0 : 1;
        }
        if (reduced) {
            for (int c = 0; c < a.P; ++c) bad |= a.segflag[c];
            a.flags[0] = bad;
        }
        return true;
    }
    template <int KID> static void seg_loop(const Args& a) {
        const int n = seg_kernel_threads(a, KID);
        for (int t = 0; t < n; ++t) seg_thread<KID>(a, t);
    }
    // pg_seg_trisolve_staged_kernel: per CTA (segment c, column group g) the chunk loads and the chunk steps as phases;
    // the carried 6-vector of every thread lives in `carry` between phases
    static void staged_trisolve(const Args& a) {
        const int nt = 128, ncx = ncol_x(a);
        for (int c = 0; c < a.P; ++c)
            for (int g = 0; g * nt < ncx; ++g) {
                const Chain ch = segment_chain(a, c);
                StageTile T;
                std::vector<double> carry(6 * nt, 0.0);
                for (int k0 = ch.lo; k0 <= ch.hi; k0 += STAGE_STEPS) {
                    const int n = std::min((int)STAGE_STEPS, ch.hi - k0 + 1);
                    for (int tid = 0; tid < nt; ++tid) stage_copy(T.M, ch.M + (size_t)k0 * 36, n * 36, tid, nt);
                    for (int tid = 0; tid < nt; ++tid) {
                        const int col = g * nt + tid;
                        if (col < ncx) forward_steps(a, ch, a.Zx, ncx, col, 0, column_ctx(a, col, 0), k0, n, T.M, &carry[6 * tid]);
                    }
                }
                std::fill(carry.begin(), carry.end(), 0.0);
                for (int k1 = ch.hi; k1 >= ch.lo; k1 -= STAGE_STEPS) {
                    const int k0 = std::max(ch.lo, k1 - (int)STAGE_STEPS + 1), n = k1 - k0 + 1;
                    for (int tid = 0; tid < nt; ++tid) stage_backward(ch, T, k0, n, tid, nt);
                    for (int tid = 0; tid < nt; ++tid) {
                        const int col = g * nt + tid;
                        if (col < ncx) backward_steps(ch, a.Zx, ncx, col, k0, n, T.S, T.M, &carry[6 * tid]);
                    }
                }
            }
    }
    bool seg(int kid, const Args& a) {
        ++launches;
        if (kid == KS_TRISOLVE && a.stage) { staged_trisolve(a); return true; }
        switch (kid) {
            case KS_TRISOLVE: seg_loop<KS_TRISOLVE>(a); break;
            case KS_REDUCED_BLOCKS: seg_loop<KS_REDUCED_BLOCKS>(a); break;
            case KS_REDUCED_RHS: seg_loop<KS_REDUCED_RHS>(a); break;
            case KS_REDUCED_TRISOLVE: seg_loop<KS_REDUCED_TRISOLVE>(a); break;
            case KS_BACKSUB: seg_loop<KS_BACKSUB>(a); break;
            default: return false;
        }
        return true;
    }
    bool partitioned(const Args& a) { return pg_partitioned_solve(*this, a); }
    bool dense(const Args& a) {
        ++launches;
        const int n = 6 * a.L, nt = 256;   // the CTA shape of pg_dense_kernel
        a.flags[1] = 0;
        for (int j = 0; j < n; ++j) {
            for (int tid = 0; tid < nt; ++tid) dense_phase_pivot(a, j, tid);
            for (int tid = 0; tid < nt; ++tid) dense_phase_column(a, j, tid, nt);
            for (int tid = 0; tid < nt; ++tid) dense_phase_update(a, j, tid, nt);
        }
        for (int tid = 0; tid < nt; ++tid) dense_phase_solve(a, tid);
        return true;
    }
    bool reduce(const Args& a) {
        ++launches;
        for (int q = 0; q < S_COUNT; ++q) {
            const double* p = a.part + part_offset(a, q);
            double v = 0.0;
            for (int i = 0; i < part_count(a, q); ++i) v = (q == S_GRAD) ? fmax(v, p[i]) : v + p[i];
            a.scal[q] = v;
        }
        return true;
    }
    bool read(const Args& a, double* scal, int32_t* flags) {
        std::memcpy(scal, a.scal, sizeof(double) * S_COUNT);
        flags[0] = a.flags[0]; flags[1] = a.flags[1];
        return true;
    }
};
}  // namespace

extern "C" {
// same argument meaning as lvio2d_pose_graph_solve (+ segments / stage: the LVIO2D_PG_SEGMENTS / LVIO2D_PG_STAGE knobs, 0 = plain path); opt5 = {max_iters, function_tol, gradient_tol, parameter_tol, initial_radius}
int pgh_solve(const Consts* C, const double* opt5, int32_t n_poses, double* poses, int32_t n_edges, const int32_t* edge_index, const double* edge_tf,
              const double* edge_weight, const double* sqrt_info, int32_t ground_p, int32_t ground_q, lvio2d_summary* summary, int32_t* launches,
              int32_t segments, int32_t stage) {
    std::vector<int32_t> ints;
    Args a;
    std::memset(&a, 0, sizeof(a));
    a.K = n_poses; a.E = n_edges;
    a.P = pg_segments(n_poses, segments);
    a.stage = stage;
    if (!pg_topology(n_poses, n_edges, edge_index, ints, &a.L, a.P)) return -1;
    a.fixed = n_edges > 0 ? edge_index[0] : -1;
    a.ground_p = ground_p; a.ground_q = ground_q;
    a.C = *C;
    std::memcpy(a.Jn, sqrt_info, sizeof(a.Jn));
    std::vector<double> dbl(pg_bind(a, ints.data(), nullptr), 0.0);
    pg_bind(a, ints.data(), dbl.data());
    std::memcpy(const_cast<double*>(a.edge_tf), edge_tf, sizeof(double) * 12 * n_edges);
    std::memcpy(const_cast<double*>(a.edge_weight), edge_weight, sizeof(double) * n_edges);
    std::memcpy(a.x, poses, sizeof(double) * 6 * n_poses);
    Options opt;
    opt.max_iters = (int)opt5[0]; opt.function_tolerance = opt5[1]; opt.gradient_tolerance = opt5[2]; opt.parameter_tolerance = opt5[3];
    opt.initial_radius = opt5[4];
    HostLauncher Lr;
    if (!pg_minimize(Lr, a, opt, summary)) return -3;
    std::memcpy(poses, a.x, sizeof(double) * 6 * n_poses);
    if (launches) *launches = Lr.launches;
    return 0;
}
// one edge: res[6], jac[6][12] over (p_i, q_i, p_j, q_j)
void pgh_edge(const double* tf12, double weight, const double* sqrt_info, const double* pose_i, const double* pose_j, double* res, double* jac) {
    Args a;
    std::memset(&a, 0, sizeof(a));
    const int32_t idx[2] = {0, 1};
    double x[12], EJ[78];
    std::memcpy(x, pose_i, 48); std::memcpy(x + 6, pose_j, 48);
    a.K = 2; a.E = 1; a.fixed = -1;
    a.edge_index = idx; a.edge_tf = tf12; a.edge_weight = &weight; a.x = x; a.EJ = EJ;
    std::memcpy(a.Jn, sqrt_info, sizeof(a.Jn));
    for (int c = 0; c < 13; ++c) edge_column(a, 0, c, EJ + 6 * c);
    for (int r = 0; r < 6; ++r) { res[r] = EJ[72 + r]; for (int c = 0; c < 12; ++c) jac[r * 12 + c] = EJ[c * 6 + r]; }
}
}
