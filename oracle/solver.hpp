// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED.
//
// solver.hpp — restatement of lvio_2d::solver on one flattened window (include/lvio2d.h):
//   problem assembly / constness / multiplicities   src/factor/solver.cpp:631-794 (tracking), :50-159 (init)
//   ceres::Solve (Ceres 1.14 TrustRegionMinimizer + LevenbergMarquardtStrategy, un-vendored dependency;
//       defaults the reference keeps: solver.cpp:795-802, :161-168)
//   marginalization + marginalization_matrix        src/factor/solver.cpp:257-442, :4-40
// Dense normal equations stand in for Ceres' SPARSE_SCHUR / DENSE_SCHUR (the linear-solver type changes
// speed, not the step).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "factors.hpp"
#include "preint.hpp"

namespace oracle {

// one window of a lvio2d_window_batch
struct Window {
    const lvio2d_window_batch* B;
    int w;  // window index
    int n;
    Window(const lvio2d_window_batch* B_, int w_) : B(B_), w(w_), n(B_->n_frames) {}
    int f(int i) const { return w * n + i; }
    uint8_t cmask(int i) const { return B->const_mask ? B->const_mask[f(i)] : 0; }
    int ref_frame(int i) const { return B->ref_frame ? B->ref_frame[f(i)] : -1; }
    const double* ref_pose(int i) const { return B->ref_pose + (size_t)f(i) * 6; }
    int64_t p0(int i) const { return B->point_offset ? B->point_offset[f(i)] : 0; }
    int64_t p1(int i) const { return B->point_offset ? B->point_offset[f(i) + 1] : 0; }
    int64_t l0(int i) const { return B->line_offset ? B->line_offset[f(i)] : 0; }
    const double* imu(int i) const { return B->imu + ((size_t)w * (n - 1) + (i - 1)) * LVIO2D_IMU_BLOB; }      // i = 1..n-1
    const double* wheel(int i) const { return B->wheel + ((size_t)w * (n - 1) + (i - 1)) * LVIO2D_WHEEL_BLOB; }
    bool has_prior() const { return B->prior_frame >= 0 && B->prior_X0 && B->prior_J; }
    const double* prior_X0() const { return B->prior_X0 + (size_t)w * 15; }
    const double* prior_J() const { return B->prior_J + (size_t)w * 225; }
};

struct Linearization {
    int dim;                 // 15 n
    std::vector<double> H;   // dense [dim][dim] = J^T J
    std::vector<double> g;   // J^T r
    double cost;             // 1/2 sum r^2 over active residual blocks
    std::vector<uint8_t> free_col;  // 1 when the column belongs to a non-constant parameter block
};

// offsets of the parameter blocks inside the 15-vector [p q v bs]
static const int kOff[4] = {0, 3, 6, 9};
static const int kLen[4] = {3, 3, 3, 6};
static const uint8_t kBit[4] = {LVIO2D_CONST_P, LVIO2D_CONST_Q, LVIO2D_CONST_V, LVIO2D_CONST_BS};

struct BlockRef { int frame; int blk; };  // frame < 0: external constant block (laser_match p1/q1)

class Evaluator {
public:
    Evaluator(const Params& P, const Window& W, int mode) : P_(P), W_(W), mode_(mode) {}

    // cost only when lin == nullptr
    double evaluate(const double* x, Linearization* lin) const {
        const int n = W_.n, dim = 15 * n;
        if (lin) {
            lin->dim = dim;
            lin->H.assign((size_t)dim * dim, 0.0);
            lin->g.assign(dim, 0.0);
            lin->free_col.assign(dim, 0);
            for (int i = 0; i < n; ++i)
                for (int b = 0; b < 4; ++b)
                    if (!is_const(i, b)) for (int k = 0; k < kLen[b]; ++k) lin->free_col[15 * i + kOff[b] + k] = 1;
        }
        double sumsq = 0.0;
        // ---- laser (solver.cpp:669-698 / :87-113; marginalisation :448-478)
        for (int j = 0; j < n; ++j) {
            const int64_t p0 = W_.p0(j), p1 = W_.p1(j);
            if (p1 <= p0) continue;
            const int k = W_.ref_frame(j);
            const double* xi = (k >= 0) ? x + 15 * k : W_.ref_pose(j);
            const double* xj = x + 15 * j;
            BlockRef refs[4] = {{k, 0}, {k, 1}, {j, 0}, {j, 1}};
            if (mode_ == 1) { refs[0].frame = -1; refs[1].frame = -1; }  // only jacobians[2],[3] are kept (solver.cpp:471-472)
            if (!any_free(refs, 4)) continue;
            const double* lines = W_.B->lines + (size_t)W_.l0(j) * 4;
            const int64_t nl = (W_.B->line_offset ? W_.B->line_offset[W_.f(j) + 1] : 0) - W_.l0(j);
            // world pose of the laser for the point side and the line side (association only)
            const Iso3<double> Twj = lie::make_tf<double>(map3(xj), map3(xj + 3)) * P_.T_imu_to_laser;
            const Iso3<double> Twi = lie::make_tf<double>(map3(xi), map3(xi + 3)) * P_.T_imu_to_laser;
            for (int64_t p = p0; p < p1; ++p) {
                int li = W_.B->point_line[p];
                if (li < 0) continue;
                if (P_.assoc_mode == 1) {
                    // BASELINE config 3 (extension; the reference freezes correspondences, trajectory.cpp:210): nearest line
                    // by perpendicular distance among the lines whose extent (+gate) contains the foot of the point
                    Vec3<double> C = Twj * Vec3<double>(W_.B->points[2 * p], W_.B->points[2 * p + 1], 0.0);
                    C.z = 0.0;
                    int best = -1;
                    double best_d = P_.assoc_max_dist;
                    for (int64_t l = 0; l < nl; ++l) {
                        const double* Lr = lines + (size_t)l * 4;
                        Vec3<double> A1 = Twi * Vec3<double>(Lr[0], Lr[1], 0.0), A2 = Twi * Vec3<double>(Lr[2], Lr[3], 0.0);
                        A1.z = 0.0; A2.z = 0.0;
                        const Vec3<double> dlt = A2 - A1;
                        const double len = norm(dlt);
                        const Vec3<double> u = dlt / len;
                        const double t = dot(u, C - A1);
                        if (t < -P_.assoc_gate || t > len + P_.assoc_gate) continue;
                        const double d = std::fabs(-u.y * (C.x - A2.x) + u.x * (C.y - A2.y));
                        if (d < best_d) { best_d = d; best = (int)l; }
                    }
                    if (best < 0) continue;
                    li = best;
                }
                const double* L = lines + (size_t)li * 4;
                const double wgt = W_.B->point_weight ? W_.B->point_weight[p] : 1.0;
                laser_point_factor fac(&P_, Vec3<double>(L[0], L[1], 0.0), Vec3<double>(L[2], L[3], 0.0),
                                       Vec3<double>(W_.B->points[2 * p], W_.B->points[2 * p + 1], 0.0), wgt);
                double r[1], J[12];
                if (lin && P_.analytic_laser) {
                    laser_point_analytic(P_, fac.a1, fac.a2, fac.c, wgt, xi, xj, r, J);
                } else if (lin) {
                    autodiff<1, 12>(fac, {xi, xi + 3, xj, xj + 3}, {3, 3, 3, 3}, r, J,
                                    [](const laser_point_factor& f, const Jet<12>* const* a, Jet<12>* res) {
                                        f(a[0], a[1], a[2], a[3], res);
                                    });
                } else {
                    fac(xi, xi + 3, xj, xj + 3, r);
                }
                double rho = r[0] * r[0];
                if (P_.huber_delta > 0 && rho > P_.huber_delta * P_.huber_delta) {
                    // ceres::HuberLoss + Corrector (rho'' <= 0: residual and Jacobian scaled by sqrt(rho'))
                    const double sq = std::sqrt(rho);
                    const double scale = std::sqrt(P_.huber_delta / sq);
                    rho = 2.0 * P_.huber_delta * sq - P_.huber_delta * P_.huber_delta;
                    r[0] *= scale;
                    if (lin) for (int c = 0; c < 12; ++c) J[c] *= scale;
                }
                if (lin) accumulate(lin, refs, 4, r, J, 1, 12);
                sumsq += rho;
            }
        }
        // ---- imu (solver.cpp:701-711) and wheel (solver.cpp:714-723)
        for (int i = 1; i < n; ++i) {
            const double* xa = x + 15 * (i - 1);
            const double* xb = x + 15 * i;
            if (W_.B->imu) {
                BlockRef refs[8] = {{i - 1, 0}, {i - 1, 1}, {i - 1, 2}, {i - 1, 3}, {i, 0}, {i, 1}, {i, 2}, {i, 3}};
                if (any_free(refs, 8)) {
                    imu_factor fac(&P_, W_.imu(i));
                    double r[15], J[15 * 30];
                    if (lin) {
                        autodiff<15, 30>(fac, {xa, xa + 3, xa + 6, xa + 9, xb, xb + 3, xb + 6, xb + 9}, {3, 3, 3, 6, 3, 3, 3, 6}, r, J,
                                         [](const imu_factor& f, const Jet<30>* const* a, Jet<30>* res) {
                                             f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], res);
                                         });
                        accumulate(lin, refs, 8, r, J, 15, 30);
                    } else {
                        fac(xa, xa + 3, xa + 6, xa + 9, xb, xb + 3, xb + 6, xb + 9, r);
                    }
                    for (int k = 0; k < 15; ++k) sumsq += r[k] * r[k];
                }
            }
            if (W_.B->wheel) {
                BlockRef refs[4] = {{i - 1, 0}, {i - 1, 1}, {i, 0}, {i, 1}};
                if (any_free(refs, 4)) {
                    wheel_odom_factor fac(&P_, W_.wheel(i));
                    double r[3], J[3 * 12];
                    if (lin) {
                        autodiff<3, 12>(fac, {xa, xa + 3, xb, xb + 3}, {3, 3, 3, 3}, r, J,
                                        [](const wheel_odom_factor& f, const Jet<12>* const* a, Jet<12>* res) {
                                            f(a[0], a[1], a[2], a[3], res);
                                        });
                        accumulate(lin, refs, 4, r, J, 3, 12);
                    } else {
                        fac(xa, xa + 3, xb, xb + 3, r);
                    }
                    for (int k = 0; k < 3; ++k) sumsq += r[k] * r[k];
                }
            }
        }
        // ---- ground: `multiplicity` copies of both factors per frame (solver.cpp:727-743, :546-586)
        for (int rep = 0; rep < W_.B->ground_multiplicity; ++rep)
            for (int j = 0; j < n; ++j) {
                BlockRef refs[2] = {{j, 0}, {j, 1}};
                if (!any_free(refs, 2)) continue;
                const double* xj = x + 15 * j;
                double r[1], J[6];
                ground_factor_p fp(&P_);
                ground_factor_q fq(&P_);
                if (lin) {
                    autodiff<1, 6>(fp, {xj, xj + 3}, {3, 3}, r, J,
                                   [](const ground_factor_p& f, const Jet<6>* const* a, Jet<6>* res) { f(a[0], a[1], res); });
                    accumulate(lin, refs, 2, r, J, 1, 6);
                } else {
                    fp(xj, xj + 3, r);
                }
                sumsq += r[0] * r[0];
                if (lin) {
                    autodiff<1, 6>(fq, {xj, xj + 3}, {3, 3}, r, J,
                                   [](const ground_factor_q& f, const Jet<6>* const* a, Jet<6>* res) { f(a[0], a[1], res); });
                    accumulate(lin, refs, 2, r, J, 1, 6);
                } else {
                    fq(xj, xj + 3, r);
                }
                sumsq += r[0] * r[0];
            }
        // ---- prior (solver.cpp:744-785, clac_prior_J :197-255)
        if (W_.has_prior()) {
            const int pf = W_.B->prior_frame;
            BlockRef refs[4] = {{pf, 0}, {pf, 1}, {pf, 2}, {pf, 3}};
            if (any_free(refs, 4)) {
                const double* xp = x + 15 * pf;
                marginalization_factor fac(W_.prior_X0(), W_.prior_J());
                double r[15], J[15 * 15];
                if (lin) {
                    autodiff<15, 15>(fac, {xp, xp + 3, xp + 6, xp + 9}, {3, 3, 3, 6}, r, J,
                                     [](const marginalization_factor& f, const Jet<15>* const* a, Jet<15>* res) {
                                         f(a[0], a[1], a[2], a[3], res);
                                     });
                    accumulate(lin, refs, 4, r, J, 15, 15);
                } else {
                    fac(xp, xp + 3, xp + 6, xp + 9, r);
                }
                for (int k = 0; k < 15; ++k) sumsq += r[k] * r[k];
            }
        }
        if (lin) lin->cost = 0.5 * sumsq;
        return 0.5 * sumsq;
    }

    bool is_const(int frame, int blk) const {
        if (frame < 0) return true;
        if (mode_ == 1) return false;
        return (W_.cmask(frame) & kBit[blk]) != 0;
    }

private:
    bool any_free(const BlockRef* refs, int nb) const {
        for (int b = 0; b < nb; ++b)
            if (!is_const(refs[b].frame, refs[b].blk)) return true;
        return false;
    }
    // H += J^T J, g += J^T r restricted to non-constant blocks
    void accumulate(Linearization* lin, const BlockRef* refs, int nb, const double* r, const double* J, int nr, int np) const {
        int col[32];
        int c = 0;
        for (int b = 0; b < nb; ++b)
            for (int k = 0; k < kLen[refs[b].blk]; ++k, ++c)
                col[c] = is_const(refs[b].frame, refs[b].blk) ? -1 : 15 * refs[b].frame + kOff[refs[b].blk] + k;
        const int dim = lin->dim;
        for (int a = 0; a < np; ++a) {
            if (col[a] < 0) continue;
            double ga = 0.0;
            for (int i = 0; i < nr; ++i) ga += J[i * np + a] * r[i];
            lin->g[col[a]] += ga;
            for (int b = 0; b < np; ++b) {
                if (col[b] < 0) continue;
                double h = 0.0;
                for (int i = 0; i < nr; ++i) h += J[i * np + a] * J[i * np + b];
                lin->H[(size_t)col[a] * dim + col[b]] += h;
            }
        }
    }
    const Params& P_;
    const Window& W_;
    int mode_;
};

// Plus of the reduced program: additive, so3_parameterization on q
inline void plus_states(const double* x, const double* delta, const std::vector<uint8_t>& free_col, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        const double* xi = x + 15 * i;
        const double* di = delta + 15 * i;
        double* oi = out + 15 * i;
        for (int k = 0; k < 15; ++k) oi[k] = free_col[15 * i + k] ? xi[k] + di[k] : xi[k];
        if (free_col[15 * i + 3]) so3_plus(xi + 3, di + 3, oi + 3);
    }
}

struct LMOptions {
    int max_num_iterations = 50;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
};
inline LMOptions lm_options_from(const lvio2d_params& p) {
    LMOptions o;
    if (p.max_iters > 0) o.max_num_iterations = p.max_iters;
    if (p.function_tolerance > 0) o.function_tolerance = p.function_tolerance;
    if (p.gradient_tolerance > 0) o.gradient_tolerance = p.gradient_tolerance;
    if (p.parameter_tolerance > 0) o.parameter_tolerance = p.parameter_tolerance;
    if (p.initial_trust_region_radius > 0) o.initial_trust_region_radius = p.initial_trust_region_radius;
    return o;
}

// Ceres 1.14 TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy, jacobi_scaling = true,
// monotonic steps, no inner iterations, no bounds, exact (direct) linear solver.
// `Problem` supplies: int dim(); double evaluate(const double* x, Linearization* lin) (cost only when lin == nullptr);
// void plus(const double* x, const double* delta, const std::vector<uint8_t>& free_col, double* out).
// seconds spent in the linear solve (scaled system build, Cholesky, substitutions) by lm_minimize on this thread since the
// last reset: bench.py reports its share of the CPU baseline (the oracle solves the dense reduced system, the reference's
// SPARSE_SCHUR would exploit the band)
inline double& lm_linear_solve_seconds() { static thread_local double t = 0.0; return t; }
template <class Problem>
inline lvio2d_summary lm_minimize(const Problem& ev, const LMOptions& opt, double* x /* [dim] in/out */) {
    const int dim = ev.dim();
    lvio2d_summary S;
    std::memset(&S, 0, sizeof(S));
    Linearization lin;
    double x_cost = ev.evaluate(x, &lin);
    S.initial_cost = x_cost;
    std::vector<int> idx;  // reduced columns
    for (int c = 0; c < dim; ++c) if (lin.free_col[c]) idx.push_back(c);
    const int m = (int)idx.size();
    auto norm_free = [&](const double* a, const double* b) {
        double s = 0.0;
        for (int c : idx) { const double d = a[c] - (b ? b[c] : 0.0); s += d * d; }
        return std::sqrt(s);
    };
    auto gradient_max_norm = [&](const Linearization& L) {
        // |x - Plus(x, -g)|_inf  (TrustRegionMinimizer::EvaluateGradientAndJacobian)
        std::vector<double> neg(dim, 0.0), xp(dim);
        for (int c : idx) neg[c] = -L.g[c];
        ev.plus(x, neg.data(), lin.free_col, xp.data());
        double mx = 0.0;
        for (int c : idx) mx = std::max(mx, std::fabs(x[c] - xp[c]));
        return mx;
    };
    double x_norm = norm_free(x, nullptr);
    // jacobi scaling from the Jacobian at iteration 0
    std::vector<double> scale(m);
    for (int a = 0; a < m; ++a) scale[a] = 1.0 / (1.0 + std::sqrt(lin.H[(size_t)idx[a] * dim + idx[a]]));
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    int num_consecutive_invalid = 0;
    int iteration = 0;
    bool last_successful = true;
    std::vector<double> Hs((size_t)m * m), gs(m), A((size_t)m * m), Lc((size_t)m * m), y(m), step(m), delta(dim), xc(dim);
    S.termination = LVIO2D_TERM_NO_CONVERGENCE;
    if (m == 0) { S.final_cost = x_cost; S.final_radius = radius; S.termination = LVIO2D_TERM_CONVERGENCE_GRADIENT; return S; }
    for (;;) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        if (iteration >= opt.max_num_iterations) { S.termination = LVIO2D_TERM_NO_CONVERGENCE; break; }
        if (last_successful && gradient_max_norm(lin) <= opt.gradient_tolerance) { S.termination = LVIO2D_TERM_CONVERGENCE_GRADIENT; break; }
        if (radius < opt.min_trust_region_radius) { S.termination = LVIO2D_TERM_CONVERGENCE_RADIUS; break; }
        ++iteration;
        // ComputeTrustRegionStep: scaled normal equations + LM diagonal
        const auto t_lin0 = std::chrono::steady_clock::now();
        for (int a = 0; a < m; ++a) {
            gs[a] = lin.g[idx[a]] * scale[a];
            for (int b = 0; b < m; ++b) Hs[(size_t)a * m + b] = lin.H[(size_t)idx[a] * dim + idx[b]] * scale[a] * scale[b];
        }
        A = Hs;
        for (int a = 0; a < m; ++a) {
            double d = std::min(std::max(Hs[(size_t)a * m + a], opt.min_lm_diagonal), opt.max_lm_diagonal);
            A[(size_t)a * m + a] += d / radius;  // lm_diagonal^2 = diagonal / radius
        }
        bool valid = cholesky_lower(A.data(), Lc.data(), m);
        double model_cost_change = 0.0;
        if (valid) {
            // (Js^T Js + D^2) y = Js^T r ; step = -y
            for (int i = 0; i < m; ++i) { double s = gs[i]; for (int j = 0; j < i; ++j) s -= Lc[(size_t)i * m + j] * y[j]; y[i] = s / Lc[(size_t)i * m + i]; }
            for (int i = m - 1; i >= 0; --i) { double s = y[i]; for (int j = i + 1; j < m; ++j) s -= Lc[(size_t)j * m + i] * step[j]; step[i] = s / Lc[(size_t)i * m + i]; }
            for (int i = 0; i < m; ++i) { step[i] = -step[i]; if (!std::isfinite(step[i])) valid = false; }
        }
        lm_linear_solve_seconds() += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_lin0).count();
        if (valid) {
            // model_cost_change = -(J step)^T (r + J step / 2) = -step^T gs - 1/2 step^T Hs step
            double sg = 0.0, shs = 0.0;
            for (int a = 0; a < m; ++a) {
                sg += step[a] * gs[a];
                double t = 0.0;
                for (int b = 0; b < m; ++b) t += Hs[(size_t)a * m + b] * step[b];
                shs += step[a] * t;
            }
            model_cost_change = -sg - 0.5 * shs;
            valid = model_cost_change > 0.0;
        }
        if (!valid) {
            // HandleInvalidStep
            ++num_consecutive_invalid;
            last_successful = false;
            ++S.num_unsuccessful_steps;
            if (num_consecutive_invalid >= opt.max_num_consecutive_invalid_steps) { S.termination = LVIO2D_TERM_FAILURE; break; }
            radius = radius / decrease_factor; decrease_factor *= 2.0;
            continue;
        }
        num_consecutive_invalid = 0;
        std::fill(delta.begin(), delta.end(), 0.0);
        for (int a = 0; a < m; ++a) delta[idx[a]] = step[a] * scale[a];
        ev.plus(x, delta.data(), lin.free_col, xc.data());
        double candidate_cost = ev.evaluate(xc.data(), nullptr);
        if (!std::isfinite(candidate_cost)) candidate_cost = std::numeric_limits<double>::max();
        // ParameterToleranceReached
        const double step_norm = norm_free(x, xc.data());
        if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { S.termination = LVIO2D_TERM_CONVERGENCE_PARAMETER; break; }
        // FunctionToleranceReached
        const double cost_change = x_cost - candidate_cost;
        if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { S.termination = LVIO2D_TERM_CONVERGENCE_FUNCTION; break; }
        const double relative_decrease = cost_change / model_cost_change;
        if (relative_decrease > opt.min_relative_decrease) {
            // HandleSuccessfulStep
            std::memcpy(x, xc.data(), sizeof(double) * dim);
            x_norm = norm_free(x, nullptr);
            x_cost = ev.evaluate(x, &lin);
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
            radius = std::min(opt.max_trust_region_radius, radius);
            decrease_factor = 2.0;
            last_successful = true;
            ++S.num_successful_steps;
        } else {
            // HandleUnsuccessfulStep
            radius = radius / decrease_factor; decrease_factor *= 2.0;
            last_successful = false;
            ++S.num_unsuccessful_steps;
        }
    }
    S.iterations = iteration;
    S.final_cost = x_cost;
    S.final_radius = radius;
    return S;
}

// the sliding-window program (solver::solve / do_init_solve)
struct WindowProblem {
    Evaluator ev;
    int n;
    WindowProblem(const Params& P, const Window& W) : ev(P, W, 0), n(W.n) {}
    int dim() const { return 15 * n; }
    double evaluate(const double* x, Linearization* lin) const { return ev.evaluate(x, lin); }
    void plus(const double* x, const double* delta, const std::vector<uint8_t>& free_col, double* out) const {
        plus_states(x, delta, free_col, n, out);
    }
};
inline lvio2d_summary lm_solve(const Params& P, const Window& W, const LMOptions& opt, double* x /* [n][15] in/out */) {
    WindowProblem prob(P, W);
    return lm_minimize(prob, opt, x);
}

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix; eigenvalues ascending (the order
// Eigen::SelfAdjointEigenSolver returns), V columns = eigenvectors, sign: largest-|.| component > 0
inline void symmetric_eig(const double* Ain, int n, double* evals, double* V) {
    std::vector<double> A(Ain, Ain + n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) { diag += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j]; }
        if (off <= 1e-60 + 1e-32 * diag) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) if (A[order[j] * n + order[j]] < A[order[i] * n + order[i]]) std::swap(order[i], order[j]);
    std::vector<double> Vs(n * n);
    for (int c = 0; c < n; ++c) {
        evals[c] = A[order[c] * n + order[c]];
        int big = 0;
        for (int k = 1; k < n; ++k) if (std::fabs(V[k * n + order[c]]) > std::fabs(V[big * n + order[c]])) big = k;
        const double sgn = V[big * n + order[c]] < 0 ? -1.0 : 1.0;
        for (int k = 0; k < n; ++k) Vs[k * n + c] = sgn * V[k * n + order[c]];
    }
    std::memcpy(V, Vs.data(), sizeof(double) * n * n);
}

// solver::marginalization (solver.cpp:257-442) for camera-off windows: keep the last frame's 15 columns.
// Outputs X0[15], J_lin[15][15], r_lin[15] (+ the intermediate Delta_H, Delta_g when asked).
inline bool marginalize(const Params& P, const Window& W, const double* x, double* X0, double* J_lin, double* r_lin,
                        double* Delta_H_out, double* Delta_g_out) {
    const int n = W.n, dim = 15 * n, r_len = 15, m_len = dim - r_len;
    Evaluator ev(P, W, 1);
    Linearization lin;
    ev.evaluate(x, &lin);
    // marginalization_matrix (solver.cpp:4-40): H = J^T J, g = -J^T R
    std::vector<double> g(dim);
    for (int i = 0; i < dim; ++i) g[i] = -lin.g[i];
    double dH[225], dg[15];
    for (int a = 0; a < 15; ++a) { dg[a] = g[m_len + a]; for (int b = 0; b < 15; ++b) dH[a * 15 + b] = lin.H[(size_t)(m_len + a) * dim + m_len + b]; }
    if (m_len > 0) {
        std::vector<double> Hmm((size_t)m_len * m_len), Hinv((size_t)m_len * m_len), T((size_t)15 * m_len);
        for (int a = 0; a < m_len; ++a) for (int b = 0; b < m_len; ++b) Hmm[(size_t)a * m_len + b] = lin.H[(size_t)a * dim + b];
        if (!inverse_pplu(Hmm.data(), Hinv.data(), m_len)) return false;
        // T = Hrm Hmm^-1
        for (int a = 0; a < 15; ++a)
            for (int b = 0; b < m_len; ++b) {
                double s = 0.0;
                for (int k = 0; k < m_len; ++k) s += lin.H[(size_t)(m_len + a) * dim + k] * Hinv[(size_t)k * m_len + b];
                T[(size_t)a * m_len + b] = s;
            }
        for (int a = 0; a < 15; ++a) {
            double s = 0.0;
            for (int k = 0; k < m_len; ++k) s += T[(size_t)a * m_len + k] * g[k];
            dg[a] -= s;
            for (int b = 0; b < 15; ++b) {
                double h = 0.0;
                for (int k = 0; k < m_len; ++k) h += T[(size_t)a * m_len + k] * lin.H[(size_t)k * dim + m_len + b];
                dH[a * 15 + b] -= h;
            }
        }
    }
    if (Delta_H_out) std::memcpy(Delta_H_out, dH, sizeof(dH));
    if (Delta_g_out) std::memcpy(Delta_g_out, dg, sizeof(dg));
    // symmetrise before the EVD (SelfAdjointEigenSolver reads the lower triangle only)
    double sym[225];
    for (int a = 0; a < 15; ++a) for (int b = 0; b < 15; ++b) sym[a * 15 + b] = dH[(a >= b ? a : b) * 15 + (a >= b ? b : a)];
    double evals[15], V[225];
    symmetric_eig(sym, 15, evals, V);
    const double eps = 1e-8;
    for (int r = 0; r < 15; ++r) {
        const double s = evals[r] > eps ? evals[r] : 0.0;
        const double sinv = evals[r] > eps ? 1.0 / evals[r] : 0.0;
        const double ss = std::sqrt(s), sis = std::sqrt(sinv);
        double acc = 0.0;
        for (int c = 0; c < 15; ++c) { J_lin[r * 15 + c] = ss * V[c * 15 + r]; acc += V[c * 15 + r] * dg[c]; }
        r_lin[r] = -(sis * acc);
    }
    std::memcpy(X0, x + 15 * (n - 1), sizeof(double) * 15);
    return true;
}

}  // namespace oracle
