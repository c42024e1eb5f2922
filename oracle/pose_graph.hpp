// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED BY THE REFERENCE (it ships no tests).
//
// pose_graph.hpp — CPU restatement of the back-end pose-graph solve (the next row; no device
// path exists for it yet):
//   keyframe_manager::solve   src/trajectory/keyframe_manager.cpp:722-838   (sequential + loop edges, ground factors on
//                             every key frame, first key frame constant, ceres::Solve with defaults, SPARSE_SCHUR)
//   edge_factor               src/factor/edge_factor.h:79-126
//   edge_noise                src/factor/edge_factor.h:4-27   (including its J(1,2) slip: the caller builds the matrix)
//   ground_factor_p / q       src/factor/ground_factor.h:27-82 (factors.hpp)
// The minimiser is the same Ceres 1.14 restatement as for the sliding window (solver.hpp: lm_minimize).
#pragma once
#include <vector>

#include "solver.hpp"

namespace oracle {
namespace posegraph {

struct Edge { int i1, i2; Iso3<double> tf12; double weight; };

struct edge_factor {
    Iso3<double> tf12;
    double weight;
    const double* Jn;   // edge_noise::J, row-major 6x6
    template <class T>
    bool operator()(const T* const p_w_i, const T* const theta_w_i, const T* const p_w_j, const T* const theta_w_j, T* res) const {
        const Iso3<T> tf_i = lie::make_tf<T>(map3(p_w_i), map3(theta_w_i));
        const Iso3<T> tf_j = lie::make_tf<T>(map3(p_w_j), map3(theta_w_j));
        const Iso3<T> error = inverse(tf_j) * tf_i * cast_iso<T>(tf12);
        Vec3<T> rp, rq;
        lie::log_SE3<T>(error, rp, rq);
        const T raw[6] = {rp.x, rp.y, rp.z, rq.x, rq.y, rq.z};
        for (int r = 0; r < 6; ++r) {
            T s = T(0.0);
            for (int k = 0; k < 6; ++k) s = s + T(Jn[r * 6 + k]) * raw[k];
            res[r] = T(weight) * s;
        }
        return true;
    }
};

struct Graph {
    const Params* P;
    int K;
    std::vector<Edge> edges;
    const double* Jn;
    bool ground_p, ground_q;
    int fixed = -2;   // constant key frame; -1: none; -2: index1 of the first edge (the reference's seq_edges[0], keyframe_manager.cpp:744-748)
    int fixed_pose() const { return fixed == -2 ? (edges.empty() ? -1 : edges[0].i1) : fixed; }
};

struct Problem {
    const Graph& G;
    explicit Problem(const Graph& g) : G(g) {}
    int dim() const { return 6 * G.K; }
    void plus(const double* x, const double* delta, const std::vector<uint8_t>& free_col, double* out) const {
        for (int i = 0; i < G.K; ++i) {
            for (int k = 0; k < 6; ++k) out[6 * i + k] = free_col[6 * i + k] ? x[6 * i + k] + delta[6 * i + k] : x[6 * i + k];
            if (free_col[6 * i + 3]) so3_plus(x + 6 * i + 3, delta + 6 * i + 3, out + 6 * i + 3);
        }
    }
    // H += J^T J, g += J^T r for the columns of the listed poses (pose 0 is constant: keyframe_manager.cpp:744-748)
    static void accumulate(Linearization* lin, const int* poses, int np, const double* r, const double* J, int nr) {
        const int dim = lin->dim, ncol = 6 * np;
        for (int a = 0; a < ncol; ++a) {
            const int ca = 6 * poses[a / 6] + a % 6;
            if (!lin->free_col[ca]) continue;
            double ga = 0.0;
            for (int i = 0; i < nr; ++i) ga += J[i * ncol + a] * r[i];
            lin->g[ca] += ga;
            for (int b = 0; b < ncol; ++b) {
                const int cb = 6 * poses[b / 6] + b % 6;
                if (!lin->free_col[cb]) continue;
                double h = 0.0;
                for (int i = 0; i < nr; ++i) h += J[i * ncol + a] * J[i * ncol + b];
                lin->H[(size_t)ca * dim + cb] += h;
            }
        }
    }
    double evaluate(const double* x, Linearization* lin) const {
        const int dimv = dim();
        if (lin) {
            lin->dim = dimv;
            lin->H.assign((size_t)dimv * dimv, 0.0);
            lin->g.assign(dimv, 0.0);
            lin->free_col.assign(dimv, 1);
            // the first sequential edge's index1 is held constant; the reference always builds edges from key frame 0
            if (G.fixed_pose() >= 0) for (int k = 0; k < 6; ++k) lin->free_col[6 * G.fixed_pose() + k] = 0;
        }
        double sumsq = 0.0;
        for (const Edge& e : G.edges) {
            edge_factor fac{e.tf12, e.weight, G.Jn};
            const double* xi = x + 6 * e.i1;
            const double* xj = x + 6 * e.i2;
            double r[6], J[6 * 12];
            if (lin) {
                autodiff<6, 12>(fac, {xi, xi + 3, xj, xj + 3}, {3, 3, 3, 3}, r, J,
                                [](const edge_factor& f, const Jet<12>* const* a, Jet<12>* res) { f(a[0], a[1], a[2], a[3], res); });
                const int poses[2] = {e.i1, e.i2};
                accumulate(lin, poses, 2, r, J, 6);
            } else {
                fac(xi, xi + 3, xj, xj + 3, r);
            }
            for (int k = 0; k < 6; ++k) sumsq += r[k] * r[k];
        }
        for (int i = 0; i < G.K; ++i) {
            const double* xi = x + 6 * i;
            if (i == G.fixed_pose()) continue;   // no variable block: Ceres moves these residual blocks to the fixed cost
            for (int which = 0; which < 2; ++which) {
                if (which == 0 ? !G.ground_p : !G.ground_q) continue;
                double r[1], J[6];
                if (which == 0) {
                    ground_factor_p fp(G.P);
                    if (lin) autodiff<1, 6>(fp, {xi, xi + 3}, {3, 3}, r, J, [](const ground_factor_p& f, const Jet<6>* const* a, Jet<6>* res) { f(a[0], a[1], res); });
                    else fp(xi, xi + 3, r);
                } else {
                    ground_factor_q fq(G.P);
                    if (lin) autodiff<1, 6>(fq, {xi, xi + 3}, {3, 3}, r, J, [](const ground_factor_q& f, const Jet<6>* const* a, Jet<6>* res) { f(a[0], a[1], res); });
                    else fq(xi, xi + 3, r);
                }
                if (lin) { const int poses[1] = {i}; accumulate(lin, poses, 1, r, J, 1); }
                sumsq += r[0] * r[0];
            }
        }
        if (lin) lin->cost = 0.5 * sumsq;
        return 0.5 * sumsq;
    }
};

inline lvio2d_summary solve(const Graph& G, const LMOptions& opt, double* poses /* [K][6] in/out */) {
    Problem prob(G);
    return lm_minimize(prob, opt, poses);
}

}  // namespace posegraph
}  // namespace oracle
