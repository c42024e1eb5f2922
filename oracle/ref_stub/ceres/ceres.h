// ceres/ceres.h — STUB (test infrastructure, oracle/_ref).  The subset of the Ceres 1.14 API the reference's front-end
// path uses, written from scratch so that src/factor/*.h and src/factor/solver.cpp compile UNMODIFIED without the real
// Ceres (absent from this image):
//   Jet (jet.h), AutoDiffCostFunction, DynamicAutoDiffCostFunction, AutoDiffLocalParameterization,
//   Problem (AddResidualBlock / SetParameterization / SetParameterBlockConstant), Solver::Options / Summary, Solve.
// Solve restates Ceres 1.14's TrustRegionMinimizer + LevenbergMarquardtStrategy with an exact dense Cholesky solve of the
// damped normal equations (what DENSE_SCHUR / SPARSE_SCHUR compute, up to rounding): Jacobi scaling 1/(1+|col|) fixed at
// iteration 0, lm diagonal = sqrt(clamp(|col|^2, 1e-6, 1e32) / radius), step validity model_cost_change > 0,
// parameter / function tolerance before the step-quality test, rho > 1e-3, radius /= max(1/3, 1-(2 rho-1)^3),
// rejection radius /= nu, nu *= 2, gradient tolerance on |x - Plus(x, -g)|_inf after successful steps only, constant
// and unused parameter blocks removed, residual blocks with no variable block counted as fixed cost.
// Cost-only evaluations run the functors on doubles, Jacobian evaluations on Jets — as Ceres does.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ceres/jet.h"

namespace ceres {

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
enum { DYNAMIC = -1 };

class CostFunction {
public:
    CostFunction() : num_residuals_(0) {}
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
    const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
    int num_residuals() const { return num_residuals_; }
protected:
    std::vector<int32_t>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
    void set_num_residuals(int n) { num_residuals_ = n; }
private:
    std::vector<int32_t> parameter_block_sizes_;
    int num_residuals_;
};

class LossFunction {
public:
    virtual ~LossFunction() {}
    virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

namespace stub_detail {
template <int... Ns> struct Sum;
template <> struct Sum<> { enum { value = 0 }; };
template <int N, int... Ns> struct Sum<N, Ns...> { enum { value = N + Sum<Ns...>::value }; };
template <size_t... I> struct Seq {};
template <size_t N, size_t... I> struct MakeSeq : MakeSeq<N - 1, N - 1, I...> {};
template <size_t... I> struct MakeSeq<0, I...> { typedef Seq<I...> type; };
template <class F, class T, size_t... I> inline bool call(const F& f, T* const* p, T* out, Seq<I...>) { return f(p[I]..., out); }
}  // namespace stub_detail

// AutoDiffCostFunction<Functor, kNumResiduals, N0, N1, ...>: one Jet<double, N0 + N1 + ...> pass (Ceres 1.14)
template <class CostFunctor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
public:
    enum { kBlocks = sizeof...(Ns), kParams = stub_detail::Sum<Ns...>::value };
    explicit AutoDiffCostFunction(CostFunctor* functor) : functor_(functor) {
        static_assert(kNumResiduals != DYNAMIC, "stub: fixed residual count only");
        set_num_residuals(kNumResiduals);
        const int sizes[] = {Ns...};
        for (int s : sizes) mutable_parameter_block_sizes()->push_back(s);
    }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
        const int sizes[] = {Ns...};
        if (!jacobians) {
            double* p[kBlocks];
            for (int b = 0; b < kBlocks; ++b) p[b] = const_cast<double*>(parameters[b]);
            return stub_detail::call(*functor_, (double* const*)p, residuals, typename stub_detail::MakeSeq<kBlocks>::type());
        }
        typedef Jet<double, kParams> J;
        std::vector<J> x((size_t)kParams), out((size_t)kNumResiduals);
        J* p[kBlocks];
        int off = 0;
        for (int b = 0; b < kBlocks; ++b) {
            p[b] = x.data() + off;
            for (int k = 0; k < sizes[b]; ++k) x[(size_t)(off + k)] = J(parameters[b][k], off + k);
            off += sizes[b];
        }
        if (!stub_detail::call(*functor_, (J* const*)p, out.data(), typename stub_detail::MakeSeq<kBlocks>::type())) return false;
        for (int r = 0; r < kNumResiduals; ++r) residuals[r] = out[(size_t)r].a;
        off = 0;
        for (int b = 0; b < kBlocks; ++b) {
            if (jacobians[b])
                for (int r = 0; r < kNumResiduals; ++r)
                    for (int k = 0; k < sizes[b]; ++k) jacobians[b][r * sizes[b] + k] = out[(size_t)r].v[off + k];
            off += sizes[b];
        }
        return true;
    }
private:
    std::unique_ptr<CostFunctor> functor_;
};

// DynamicAutoDiffCostFunction<Functor, Stride>: passes of Stride derivative directions (Ceres 1.14)
template <class CostFunctor, int Stride = 4>
class DynamicAutoDiffCostFunction : public CostFunction {
public:
    explicit DynamicAutoDiffCostFunction(CostFunctor* functor) : functor_(functor) {}
    void AddParameterBlock(int size) { mutable_parameter_block_sizes()->push_back(size); }
    void SetNumResiduals(int n) { set_num_residuals(n); }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
        if (!jacobians) return (*functor_)(parameters, residuals);
        typedef Jet<double, Stride> J;
        const std::vector<int32_t>& sizes = parameter_block_sizes();
        const int nb = (int)sizes.size(), nr = num_residuals();
        int total = 0;
        for (int s : sizes) total += s;
        std::vector<J> x((size_t)total), out((size_t)nr);
        std::vector<J*> p((size_t)nb);
        std::vector<int> start((size_t)nb);
        { int off = 0; for (int b = 0; b < nb; ++b) { p[(size_t)b] = x.data() + off; start[(size_t)b] = off; off += sizes[(size_t)b]; } }
        for (int pass = 0; pass * Stride < std::max(total, 1); ++pass) {
            for (int b = 0; b < nb; ++b)
                for (int k = 0; k < sizes[(size_t)b]; ++k) {
                    const int g = start[(size_t)b] + k;
                    J j(parameters[b][k]);
                    if (g >= pass * Stride && g < (pass + 1) * Stride) j.v[g - pass * Stride] = 1.0;
                    x[(size_t)g] = j;
                }
            if (!(*functor_)((J const* const*)p.data(), out.data())) return false;
            if (pass == 0) for (int r = 0; r < nr; ++r) residuals[r] = out[(size_t)r].a;
            for (int b = 0; b < nb; ++b) {
                if (!jacobians[b]) continue;
                for (int k = 0; k < sizes[(size_t)b]; ++k) {
                    const int g = start[(size_t)b] + k;
                    if (g < pass * Stride || g >= (pass + 1) * Stride) continue;
                    for (int r = 0; r < nr; ++r) jacobians[b][r * sizes[(size_t)b] + k] = out[(size_t)r].v[g - pass * Stride];
                }
            }
        }
        return true;
    }
private:
    std::unique_ptr<CostFunctor> functor_;
};

class LocalParameterization {
public:
    virtual ~LocalParameterization() {}
    virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
    virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;   // row-major GlobalSize x LocalSize
    virtual int GlobalSize() const = 0;
    virtual int LocalSize() const = 0;
};

template <class Functor, int kGlobalSize, int kLocalSize>
class AutoDiffLocalParameterization : public LocalParameterization {
public:
    AutoDiffLocalParameterization() : functor_(new Functor()) {}
    explicit AutoDiffLocalParameterization(Functor* f) : functor_(f) {}
    bool Plus(const double* x, const double* delta, double* x_plus_delta) const override { return (*functor_)(x, delta, x_plus_delta); }
    bool ComputeJacobian(const double* x, double* jacobian) const override {
        typedef Jet<double, kLocalSize> J;
        J xj[kGlobalSize], dj[kLocalSize], out[kGlobalSize];
        for (int i = 0; i < kGlobalSize; ++i) xj[i] = J(x[i]);
        for (int i = 0; i < kLocalSize; ++i) dj[i] = J(0.0, i);
        if (!(*functor_)((const J*)xj, (const J*)dj, (J*)out)) return false;
        for (int i = 0; i < kGlobalSize; ++i) for (int j = 0; j < kLocalSize; ++j) jacobian[i * kLocalSize + j] = out[i].v[j];
        return true;
    }
    int GlobalSize() const override { return kGlobalSize; }
    int LocalSize() const override { return kLocalSize; }
private:
    std::unique_ptr<Functor> functor_;
};

typedef void* ResidualBlockId;

class Problem {
public:
    struct Options {
        Ownership cost_function_ownership = TAKE_OWNERSHIP, loss_function_ownership = TAKE_OWNERSHIP, local_parameterization_ownership = TAKE_OWNERSHIP;
    };
    struct ParamBlock { double* ptr; int size; bool constant; LocalParameterization* lp; };
    struct ResBlock { CostFunction* cost; std::vector<double*> params; };
    Problem() {}
    explicit Problem(const Options&) {}
    Problem(const Problem&) = delete;
    ~Problem() {
        std::set<CostFunction*> costs;
        for (auto& r : residual_blocks_) costs.insert(r.cost);
        for (CostFunction* c : costs) delete c;
        std::set<LocalParameterization*> lps;
        for (auto& p : param_blocks_) if (p.lp) lps.insert(p.lp);
        for (LocalParameterization* l : lps) delete l;
    }
    template <class... Ptrs> ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, Ptrs... xs) {
        return AddResidualBlock(cost, loss, std::vector<double*>{x0, xs...});
    }
    ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, const std::vector<double*>& params) {
        (void)loss;   // the reference passes nullptr everywhere (solver.cpp:56, :635)
        const std::vector<int32_t>& sizes = cost->parameter_block_sizes();
        if (sizes.size() != params.size()) throw std::runtime_error("ceres stub: parameter block count mismatch");
        for (size_t i = 0; i < params.size(); ++i) AddParameterBlock(params[i], sizes[i]);
        residual_blocks_.push_back(ResBlock{cost, params});
        return (ResidualBlockId)(residual_blocks_.size());
    }
    void AddParameterBlock(double* values, int size, LocalParameterization* lp = nullptr) {
        auto it = index_.find(values);
        if (it == index_.end()) {
            index_[values] = (int)param_blocks_.size();
            param_blocks_.push_back(ParamBlock{values, size, false, lp});
        } else {
            if (param_blocks_[(size_t)it->second].size != size) throw std::runtime_error("ceres stub: parameter block size mismatch");
            if (lp) param_blocks_[(size_t)it->second].lp = lp;
        }
    }
    void SetParameterization(double* values, LocalParameterization* lp) { block(values).lp = lp; }
    void SetParameterBlockConstant(double* values) { block(values).constant = true; }
    void SetParameterBlockVariable(double* values) { block(values).constant = false; }
    int NumParameterBlocks() const { return (int)param_blocks_.size(); }
    int NumResidualBlocks() const { return (int)residual_blocks_.size(); }
    int NumResiduals() const { int n = 0; for (auto& r : residual_blocks_) n += r.cost->num_residuals(); return n; }
    const std::vector<ParamBlock>& stub_param_blocks() const { return param_blocks_; }
    const std::vector<ResBlock>& stub_residual_blocks() const { return residual_blocks_; }
    int stub_index(double* p) const { auto it = index_.find(p); return it == index_.end() ? -1 : it->second; }
private:
    ParamBlock& block(double* values) {
        auto it = index_.find(values);
        if (it == index_.end()) throw std::runtime_error("ceres stub: unknown parameter block");
        return param_blocks_[(size_t)it->second];
    }
    std::vector<ParamBlock> param_blocks_;
    std::vector<ResBlock> residual_blocks_;
    std::map<double*, int> index_;
};

struct IterationSummary {
    int iteration = 0;
    bool step_is_valid = false, step_is_successful = false;
    double cost = 0, cost_change = 0, gradient_max_norm = 0, step_norm = 0, relative_decrease = 0, trust_region_radius = 0;
};

class Solver {
public:
    struct Options {
        LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
        TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
        bool use_nonmonotonic_steps = false, minimizer_progress_to_stdout = false, jacobi_scaling = true;
        int max_num_iterations = 50, num_threads = 1, num_linear_solver_threads = 1, max_num_consecutive_invalid_steps = 5;
        double max_solver_time_in_seconds = 1e9;
        double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
        double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
        double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    };
    struct Summary {
        TerminationType termination_type = FAILURE;
        std::string message;
        double initial_cost = 0, final_cost = 0, fixed_cost = 0;
        int num_successful_steps = 0, num_unsuccessful_steps = 0;
        std::vector<IterationSummary> iterations;
        double total_time_in_seconds = 0;
        // stub extras (read by oracle/ref_driver.cpp): LM iteration count as Ceres numbers it, final radius, and which
        // test ended the run: 0 max iterations, 1 function tolerance, 2 parameter tolerance, 3 gradient tolerance,
        // 4 minimum radius, 5 failure (the codes of include/lvio2d.h)
        int stub_iterations = 0, stub_termination = 0;
        double stub_final_radius = 0;
        std::string BriefReport() const { return "ceres stub: " + message; }
        std::string FullReport() const { return BriefReport(); }
        bool IsSolutionUsable() const { return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE; }
    };
};

namespace stub_detail {
inline Solver::Summary& last_summary() { static Solver::Summary s; return s; }
// test knob (not Ceres): > 0 caps max_num_iterations below what the caller asked for, so that two implementations can be
// compared in lock-step before the reference's 50-iteration runs enter the rounding-dominated zig-zag regime
inline int& max_iterations_cap() { static int cap = 0; return cap; }

// the reduced program: variable, used parameter blocks in order of first use; dense evaluation
struct Program {
    struct Var { double* user; int size, local, off, loff; LocalParameterization* lp; };
    std::vector<Var> vars;
    std::map<double*, int> var_index;
    std::vector<const Problem::ResBlock*> blocks;   // residual blocks with at least one variable block
    int n_global = 0, n_local = 0, n_res = 0;
    double fixed_cost = 0;

    explicit Program(const Problem& pb) {
        const auto& pbs = pb.stub_param_blocks();
        for (const auto& rb : pb.stub_residual_blocks()) {
            bool any = false;
            for (double* p : rb.params) if (!pbs[(size_t)pb.stub_index(p)].constant) any = true;
            if (!any) {
                std::vector<double> r((size_t)rb.cost->num_residuals());
                std::vector<const double*> pp(rb.params.begin(), rb.params.end());
                rb.cost->Evaluate(pp.data(), r.data(), nullptr);
                for (double v : r) fixed_cost += 0.5 * v * v;
                continue;
            }
            blocks.push_back(&rb);
            n_res += rb.cost->num_residuals();
            for (double* p : rb.params) {
                const auto& b = pbs[(size_t)pb.stub_index(p)];
                if (b.constant || var_index.count(p)) continue;
                var_index[p] = (int)vars.size();
                const int local = b.lp ? b.lp->LocalSize() : b.size;
                vars.push_back(Var{p, b.size, local, n_global, n_local, b.lp});
                n_global += b.size;
                n_local += local;
            }
        }
    }
    void gather(std::vector<double>& x) const { x.resize((size_t)n_global); for (const Var& v : vars) std::memcpy(x.data() + v.off, v.user, sizeof(double) * (size_t)v.size); }
    void scatter(const std::vector<double>& x) const { for (const Var& v : vars) std::memcpy(v.user, x.data() + v.off, sizeof(double) * (size_t)v.size); }
    bool plus(const std::vector<double>& x, const std::vector<double>& delta, std::vector<double>& out) const {
        out.resize((size_t)n_global);
        for (const Var& v : vars) {
            if (v.lp) { if (!v.lp->Plus(x.data() + v.off, delta.data() + v.loff, out.data() + v.off)) return false; }
            else for (int k = 0; k < v.size; ++k) out[(size_t)(v.off + k)] = x[(size_t)(v.off + k)] + delta[(size_t)(v.loff + k)];
        }
        return true;
    }
    // cost = 1/2 |r|^2; with jac: residuals and the dense row-major [n_res x n_local] Jacobian in the tangent space
    bool evaluate(const std::vector<double>& x, double* cost, std::vector<double>* res, std::vector<double>* jac) const {
        if (res) res->assign((size_t)n_res, 0.0);
        if (jac) jac->assign((size_t)n_res * (size_t)n_local, 0.0);
        double c = 0.0;
        int row = 0;
        std::vector<double> r, plusj;
        std::vector<std::vector<double>> jb;
        for (const Problem::ResBlock* rb : blocks) {
            const int nr = rb->cost->num_residuals();
            const size_t np = rb->params.size();
            r.assign((size_t)nr, 0.0);
            std::vector<const double*> pp(np);
            std::vector<double*> jp(np, nullptr);
            std::vector<int> vi(np, -1);
            if (jac) jb.assign(np, std::vector<double>());
            for (size_t i = 0; i < np; ++i) {
                auto it = var_index.find(rb->params[i]);
                if (it == var_index.end()) { pp[i] = rb->params[i]; continue; }
                vi[i] = it->second;
                pp[i] = x.data() + vars[(size_t)it->second].off;
                if (jac) { jb[i].assign((size_t)nr * (size_t)vars[(size_t)it->second].size, 0.0); jp[i] = jb[i].data(); }
            }
            if (!rb->cost->Evaluate(pp.data(), r.data(), jac ? jp.data() : nullptr)) return false;
            for (int k = 0; k < nr; ++k) { if (!std::isfinite(r[(size_t)k])) return false; c += r[(size_t)k] * r[(size_t)k]; if (res) (*res)[(size_t)(row + k)] = r[(size_t)k]; }
            if (jac)
                for (size_t i = 0; i < np; ++i) {
                    if (vi[i] < 0) continue;
                    const Var& v = vars[(size_t)vi[i]];
                    // the same block may appear twice in one residual block: accumulate
                    if (v.lp) {
                        plusj.assign((size_t)v.size * (size_t)v.local, 0.0);
                        if (!v.lp->ComputeJacobian(pp[i], plusj.data())) return false;
                        for (int k = 0; k < nr; ++k)
                            for (int l = 0; l < v.local; ++l) {
                                double s = 0.0;
                                for (int g = 0; g < v.size; ++g) s += jb[i][(size_t)(k * v.size + g)] * plusj[(size_t)(g * v.local + l)];
                                (*jac)[(size_t)(row + k) * (size_t)n_local + (size_t)(v.loff + l)] += s;
                            }
                    } else {
                        for (int k = 0; k < nr; ++k)
                            for (int g = 0; g < v.size; ++g) (*jac)[(size_t)(row + k) * (size_t)n_local + (size_t)(v.loff + g)] += jb[i][(size_t)(k * v.size + g)];
                    }
                }
            row += nr;
        }
        *cost = 0.5 * c;
        return true;
    }
};

inline bool cholesky_solve(std::vector<double>& A, std::vector<double>& b, int n) {   // A (row-major, SPD) x = b, in place
    for (int k = 0; k < n; ++k) {
        double d = A[(size_t)k * n + k];
        for (int j = 0; j < k; ++j) d -= A[(size_t)k * n + j] * A[(size_t)k * n + j];
        if (!(d > 0.0) || !std::isfinite(d)) return false;
        d = std::sqrt(d);
        A[(size_t)k * n + k] = d;
        for (int i = k + 1; i < n; ++i) {
            double s = A[(size_t)i * n + k];
            for (int j = 0; j < k; ++j) s -= A[(size_t)i * n + j] * A[(size_t)k * n + j];
            A[(size_t)i * n + k] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) { double s = b[(size_t)i]; for (int j = 0; j < i; ++j) s -= A[(size_t)i * n + j] * b[(size_t)j]; b[(size_t)i] = s / A[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[(size_t)i]; for (int j = i + 1; j < n; ++j) s -= A[(size_t)j * n + i] * b[(size_t)j]; b[(size_t)i] = s / A[(size_t)i * n + i]; }
    return true;
}
}  // namespace stub_detail

inline void Solve(const Solver::Options& opt, Problem* problem, Solver::Summary* summary) {
    using stub_detail::Program;
    Solver::Summary S;
    Program prog(*problem);
    const int n = prog.n_local, m = prog.n_res;
    std::vector<double> x, xc, res, jac, delta((size_t)n), neg((size_t)n), proj;
    prog.gather(x);
    double x_cost = 0.0;
    S.fixed_cost = prog.fixed_cost;
    auto finish = [&](TerminationType t, int code, const char* msg, int iters, double radius) {
        S.termination_type = t; S.stub_termination = code; S.message = msg; S.stub_iterations = iters; S.stub_final_radius = radius;
        S.final_cost = x_cost + prog.fixed_cost;
        prog.scatter(x);
        stub_detail::last_summary() = S;
        if (summary) *summary = S;
    };
    if (n == 0) { S.initial_cost = prog.fixed_cost; finish(CONVERGENCE, 3, "no variable parameter blocks", 0, opt.initial_trust_region_radius); return; }
    if (!prog.evaluate(x, &x_cost, &res, &jac)) { finish(FAILURE, 5, "initial evaluation failed", 0, opt.initial_trust_region_radius); return; }
    S.initial_cost = x_cost + prog.fixed_cost;
    auto x_norm_of = [&](const std::vector<double>& v) { double s = 0; for (double e : v) s += e * e; return std::sqrt(s); };
    double x_norm = x_norm_of(x);
    // gradient (tangent space) and |x - Plus(x, -g)|_inf from the UNSCALED Jacobian
    std::vector<double> grad((size_t)n);
    auto gradient_max_norm = [&]() {
        for (int c = 0; c < n; ++c) { double s = 0; for (int r = 0; r < m; ++r) s += jac[(size_t)r * n + c] * res[(size_t)r]; grad[(size_t)c] = s; neg[(size_t)c] = -s; }
        if (!prog.plus(x, neg, proj)) return std::numeric_limits<double>::infinity();
        double mx = 0;
        for (size_t i = 0; i < x.size(); ++i) mx = std::max(mx, std::fabs(x[i] - proj[i]));
        return mx;
    };
    double gmax = gradient_max_norm();
    std::vector<double> scale((size_t)n, 1.0);
    if (opt.jacobi_scaling)
        for (int c = 0; c < n; ++c) { double s = 0; for (int r = 0; r < m; ++r) s += jac[(size_t)r * n + c] * jac[(size_t)r * n + c]; scale[(size_t)c] = 1.0 / (1.0 + std::sqrt(s)); }
    auto scale_jac = [&]() { for (int r = 0; r < m; ++r) for (int c = 0; c < n; ++c) jac[(size_t)r * n + c] *= scale[(size_t)c]; };
    scale_jac();
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    int iteration = 0, invalid = 0;
    const int max_iterations = stub_detail::max_iterations_cap() > 0 ? std::min(opt.max_num_iterations, stub_detail::max_iterations_cap()) : opt.max_num_iterations;
    bool last_successful = true;
    std::vector<double> A((size_t)n * n), rhs((size_t)n), step((size_t)n), model((size_t)m), diag((size_t)n);
    for (;;) {
        if (iteration >= max_iterations) { finish(NO_CONVERGENCE, 0, "maximum number of iterations reached", iteration, radius); return; }
        if (last_successful && gmax <= opt.gradient_tolerance) { finish(CONVERGENCE, 3, "gradient tolerance reached", iteration, radius); return; }
        if (radius < opt.min_trust_region_radius) { finish(CONVERGENCE, 4, "minimum trust region radius reached", iteration, radius); return; }
        ++iteration;
        // LevenbergMarquardtStrategy::ComputeStep on the scaled Jacobian
        for (int a = 0; a < n; ++a) {
            for (int b = a; b < n; ++b) { double s = 0; for (int r = 0; r < m; ++r) s += jac[(size_t)r * n + a] * jac[(size_t)r * n + b]; A[(size_t)a * n + b] = A[(size_t)b * n + a] = s; }
            double g = 0; for (int r = 0; r < m; ++r) g += jac[(size_t)r * n + a] * res[(size_t)r];
            rhs[(size_t)a] = g;
        }
        for (int a = 0; a < n; ++a) { diag[(size_t)a] = std::min(std::max(A[(size_t)a * n + a], opt.min_lm_diagonal), opt.max_lm_diagonal); A[(size_t)a * n + a] += diag[(size_t)a] / radius; }
        step = rhs;
        bool valid = stub_detail::cholesky_solve(A, step, n);
        double model_cost_change = 0.0;
        if (valid) for (int a = 0; a < n; ++a) { step[(size_t)a] = -step[(size_t)a]; if (!std::isfinite(step[(size_t)a])) valid = false; }
        if (valid) {
            // model_cost_change = -(J step)^T (r + J step / 2)
            for (int r = 0; r < m; ++r) { double s = 0; for (int c = 0; c < n; ++c) s += jac[(size_t)r * n + c] * step[(size_t)c]; model[(size_t)r] = s; }
            double s = 0; for (int r = 0; r < m; ++r) s += model[(size_t)r] * (res[(size_t)r] + model[(size_t)r] / 2.0);
            model_cost_change = -s;
            valid = model_cost_change > 0.0;
        }
        if (!valid) {
            ++invalid; last_successful = false; ++S.num_unsuccessful_steps;
            if (invalid >= opt.max_num_consecutive_invalid_steps) { finish(FAILURE, 5, "too many consecutive invalid steps", iteration, radius); return; }
            radius = radius / decrease_factor; decrease_factor *= 2.0;
            continue;
        }
        invalid = 0;
        for (int a = 0; a < n; ++a) delta[(size_t)a] = step[(size_t)a] * scale[(size_t)a];
        double candidate_cost = std::numeric_limits<double>::max();
        if (prog.plus(x, delta, xc)) { double c = 0; if (prog.evaluate(xc, &c, nullptr, nullptr) && std::isfinite(c)) candidate_cost = c; }
        double sn = 0; for (size_t i = 0; i < x.size(); ++i) sn += (x[i] - xc[i]) * (x[i] - xc[i]);
        const double step_norm = std::sqrt(sn);
        if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { finish(CONVERGENCE, 2, "parameter tolerance reached", iteration, radius); return; }
        const double cost_change = x_cost - candidate_cost;
        if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { finish(CONVERGENCE, 1, "function tolerance reached", iteration, radius); return; }
        const double rho = cost_change / model_cost_change;
        if (rho > opt.min_relative_decrease) {
            x = xc;
            x_norm = x_norm_of(x);
            if (!prog.evaluate(x, &x_cost, &res, &jac)) { finish(FAILURE, 5, "evaluation failed", iteration, radius); return; }
            gmax = gradient_max_norm();
            scale_jac();
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
            radius = std::min(opt.max_trust_region_radius, radius);
            decrease_factor = 2.0;
            last_successful = true;
            ++S.num_successful_steps;
        } else {
            radius = radius / decrease_factor; decrease_factor *= 2.0;
            last_successful = false;
            ++S.num_unsuccessful_steps;
        }
    }
}

}  // namespace ceres
