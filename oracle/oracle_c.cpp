// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED.
//
// oracle_c.cpp — extern "C" surface of the CPU restatement, loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).  Mirrors include/lvio2d.h
// entry point by entry point so a parity test feeds the same structs to both sides.
#include <cstdio>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>

#include "solver.hpp"
#include "laser_lines.hpp"
#include "scan_points.hpp"
#include "laser_match.hpp"
#include "pose_graph.hpp"

using namespace oracle;

extern "C" {

int oracle_abi_version() { return LVIO2D_ABI_VERSION; }

// ---- primitives (src/utilies/common.h:86-95, :121-163)
void oracle_exp_so3(const double* so3, double* R9) {
    Mat3<double> R = lie::exp_so3<double>(Vec3<double>(so3[0], so3[1], so3[2]));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R.m[i][j];
}
void oracle_log_SO3(const double* R9, double* so3) {
    Mat3<double> R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = R9[i * 3 + j];
    Vec3<double> v = lie::log_SO3<double>(R);
    so3[0] = v.x; so3[1] = v.y; so3[2] = v.z;
}
void oracle_normalize_so3(double* so3) {
    Vec3<double> v(so3[0], so3[1], so3[2]);
    lie::normalize_so3(v);
    so3[0] = v.x; so3[1] = v.y; so3[2] = v.z;
}
double oracle_dis_from_line(const double* p, const double* p1, const double* p2) {
    return e_laser::dis_from_line<double>(Vec3<double>(p[0], p[1], p[2]), Vec3<double>(p1[0], p1[1], p1[2]),
                                          Vec3<double>(p2[0], p2[1], p2[2]));
}
void oracle_so3_plus(const double* theta, const double* delta, double* out) { so3_plus(theta, delta, out); }

// ---- per-factor residual + Jacobian = auto_diff::compute_res_and_jacobi (common.h:201-217)
int oracle_eval_laser_factor(const lvio2d_params* p, const double* l1_p1, const double* l1_p2, const double* l2_p1,
                             const double* l2_p2, const double* pose_i, const double* pose_j, double* res, double* jac) {
    Params P(*p);
    laser_factor fac(&P, Vec3<double>(l1_p1[0], l1_p1[1], l1_p1[2]), Vec3<double>(l1_p2[0], l1_p2[1], l1_p2[2]),
                     Vec3<double>(l2_p1[0], l2_p1[1], l2_p1[2]), Vec3<double>(l2_p2[0], l2_p2[1], l2_p2[2]));
    autodiff<2, 12>(fac, {pose_i, pose_i + 3, pose_j, pose_j + 3}, {3, 3, 3, 3}, res, jac,
                    [](const laser_factor& f, const Jet<12>* const* a, Jet<12>* r) { f(a[0], a[1], a[2], a[3], r); });
    return 0;
}
int oracle_eval_laser_point(const lvio2d_params* p, const double* a1, const double* a2, const double* c, double weight,
                            const double* pose_i, const double* pose_j, double* res, double* jac) {
    Params P(*p);
    laser_point_factor fac(&P, Vec3<double>(a1[0], a1[1], 0.0), Vec3<double>(a2[0], a2[1], 0.0), Vec3<double>(c[0], c[1], 0.0), weight);
    autodiff<1, 12>(fac, {pose_i, pose_i + 3, pose_j, pose_j + 3}, {3, 3, 3, 3}, res, jac,
                    [](const laser_point_factor& f, const Jet<12>* const* a, Jet<12>* r) { f(a[0], a[1], a[2], a[3], r); });
    return 0;
}
int oracle_eval_imu_factor(const lvio2d_params* p, const double* blob, const double* si, const double* sj, double* res, double* jac) {
    Params P(*p);
    imu_factor fac(&P, blob);
    autodiff<15, 30>(fac, {si, si + 3, si + 6, si + 9, sj, sj + 3, sj + 6, sj + 9}, {3, 3, 3, 6, 3, 3, 3, 6}, res, jac,
                     [](const imu_factor& f, const Jet<30>* const* a, Jet<30>* r) { f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], r); });
    return 0;
}
int oracle_eval_wheel_factor(const lvio2d_params* p, const double* blob, const double* pose_i, const double* pose_j, double* res, double* jac) {
    Params P(*p);
    wheel_odom_factor fac(&P, blob);
    autodiff<3, 12>(fac, {pose_i, pose_i + 3, pose_j, pose_j + 3}, {3, 3, 3, 3}, res, jac,
                    [](const wheel_odom_factor& f, const Jet<12>* const* a, Jet<12>* r) { f(a[0], a[1], a[2], a[3], r); });
    return 0;
}
int oracle_eval_ground_factors(const lvio2d_params* p, const double* pose, double* res, double* jac) {
    Params P(*p);
    ground_factor_p fp(&P);
    ground_factor_q fq(&P);
    autodiff<1, 6>(fp, {pose, pose + 3}, {3, 3}, res, jac,
                   [](const ground_factor_p& f, const Jet<6>* const* a, Jet<6>* r) { f(a[0], a[1], r); });
    autodiff<1, 6>(fq, {pose, pose + 3}, {3, 3}, res + 1, jac ? jac + 6 : nullptr,
                   [](const ground_factor_q& f, const Jet<6>* const* a, Jet<6>* r) { f(a[0], a[1], r); });
    return 0;
}
int oracle_eval_prior_factor(const double* X0, const double* J, const double* state, double* res, double* jac) {
    marginalization_factor fac(X0, J);
    autodiff<15, 15>(fac, {state, state + 3, state + 6, state + 9}, {3, 3, 3, 6}, res, jac,
                     [](const marginalization_factor& f, const Jet<15>* const* a, Jet<15>* r) { f(a[0], a[1], a[2], a[3], r); });
    return 0;
}

// ---- preintegration
int oracle_imu_preintegrate(const lvio2d_params* p, int32_t n_intervals, const int64_t* sample_offset, const double* samples,
                            const double* bias0, double* out_blobs) {
    Params P(*p);
    int bad = 0;
    for (int i = 0; i < n_intervals; ++i) {
        imu_preintegraption pre(&P);
        pre.reset(bias0 + 6 * i, bias0 + 6 * i + 3);
        for (int64_t s = sample_offset[i]; s < sample_offset[i + 1]; ++s) pre.update(samples[7 * s], samples + 7 * s + 1, samples + 7 * s + 4);
        if (!pre.result(out_blobs + (size_t)i * LVIO2D_IMU_BLOB)) ++bad;
    }
    return bad ? -1 : 0;
}
int oracle_wheel_preintegrate(const lvio2d_params* p, int32_t n_intervals, const int64_t* step_offset, const double* steps, double* out_blobs) {
    Params P(*p);
    for (int i = 0; i < n_intervals; ++i) {
        wheel_odom_preintegration pre(&P);
        for (int64_t s = step_offset[i]; s < step_offset[i + 1]; ++s) pre.update_by_v(steps[7 * s], steps + 7 * s + 1, steps + 7 * s + 4);
        pre.result(out_blobs + (size_t)i * LVIO2D_WHEEL_BLOB);
    }
    return 0;
}

// ---- window level
// states == NULL: use batch->states
int oracle_linearize(const lvio2d_params* p, const lvio2d_window_batch* batch, const double* states, int32_t mode, double* H,
                     double* g, double* cost) {
    Params P(*p);
    const int n = batch->n_frames, dim = 15 * n;
    const double* X = states ? states : batch->states;
    for (int w = 0; w < batch->n_windows; ++w) {
        Window W(batch, w);
        Evaluator ev(P, W, mode);
        Linearization lin;
        ev.evaluate(X + (size_t)w * dim, &lin);
        if (H) std::memcpy(H + (size_t)w * dim * dim, lin.H.data(), sizeof(double) * dim * dim);
        if (g) std::memcpy(g + (size_t)w * dim, lin.g.data(), sizeof(double) * dim);
        if (cost) cost[w] = lin.cost;
    }
    return 0;
}
int oracle_cost(const lvio2d_params* p, const lvio2d_window_batch* batch, const double* states, double* cost) {
    Params P(*p);
    const int dim = 15 * batch->n_frames;
    const double* X = states ? states : batch->states;
    for (int w = 0; w < batch->n_windows; ++w) {
        Window W(batch, w);
        Evaluator ev(P, W, 0);
        cost[w] = ev.evaluate(X + (size_t)w * dim, nullptr);
    }
    return 0;
}
// solver::solve / do_init_solve.  n_threads > 1 parallelises over windows (std::thread) for the all-cores
// baseline; the reference itself is single threaded (solver.cpp:798).
static int solve_impl(const lvio2d_params* p, const lvio2d_window_batch* batch, double* states_out, lvio2d_summary* summaries, int32_t n_threads,
                      bool analytic_laser);
int oracle_solve(const lvio2d_params* p, const lvio2d_window_batch* batch, double* states_out, lvio2d_summary* summaries, int32_t n_threads) {
    return solve_impl(p, batch, states_out, summaries, n_threads, false);
}
// seconds the calling thread's single-threaded solves spent in the linear solve since the last reset (reset != 0 clears it)
double oracle_linear_solve_seconds(int32_t reset) {
    const double t = lm_linear_solve_seconds();
    if (reset) lm_linear_solve_seconds() = 0.0;
    return t;
}
// the "analytic" CPU-baseline flavour: closed-form Jacobian for the scan points, Jets for the rest
int oracle_solve_analytic(const lvio2d_params* p, const lvio2d_window_batch* batch, double* states_out, lvio2d_summary* summaries,
                          int32_t n_threads) {
    return solve_impl(p, batch, states_out, summaries, n_threads, true);
}
int oracle_eval_laser_point_analytic(const lvio2d_params* p, const double* a1, const double* a2, const double* c, double weight,
                                     const double* pose_i, const double* pose_j, double* res, double* jac) {
    Params P(*p);
    laser_point_analytic(P, Vec3<double>(a1[0], a1[1], 0.0), Vec3<double>(a2[0], a2[1], 0.0), Vec3<double>(c[0], c[1], 0.0), weight, pose_i,
                         pose_j, res, jac);
    return 0;
}
static int solve_impl(const lvio2d_params* p, const lvio2d_window_batch* batch, double* states_out, lvio2d_summary* summaries, int32_t n_threads,
                      bool analytic_laser) {
    Params P(*p);
    P.analytic_laser = analytic_laser;
    LMOptions opt = lm_options_from(*p);
    const int dim = 15 * batch->n_frames;
    std::memcpy(states_out, batch->states, sizeof(double) * (size_t)dim * batch->n_windows);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > batch->n_windows) n_threads = batch->n_windows;
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            const int w = next.fetch_add(1);
            if (w >= batch->n_windows) break;
            Window W(batch, w);
            lvio2d_summary S = lm_solve(P, W, opt, states_out + (size_t)w * dim);
            if (summaries) summaries[w] = S;
        }
    };
    if (n_threads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    return 0;
}
int oracle_marginalize(const lvio2d_params* p, const lvio2d_window_batch* batch, const double* states, double* X0, double* J_lin,
                       double* r_lin, double* Delta_H, double* Delta_g) {
    Params P(*p);
    const int dim = 15 * batch->n_frames;
    const double* X = states ? states : batch->states;
    int bad = 0;
    for (int w = 0; w < batch->n_windows; ++w) {
        Window W(batch, w);
        if (!marginalize(P, W, X + (size_t)w * dim, X0 + 15 * w, J_lin + 225 * w, r_lin + 15 * w, Delta_H ? Delta_H + 225 * w : nullptr,
                         Delta_g ? Delta_g + 15 * w : nullptr))
            ++bad;
    }
    return bad ? -1 : 0;
}
// ---- scan -> line segments (laser_manager::spawn_scan), same argument layout as lvio2d_extract_lines (host buffers)
static void normalise_sign(double* v) {
    int k = 0;
    for (int i = 1; i < 3; ++i) if (std::fabs(v[i]) > std::fabs(v[k])) k = i;
    if (v[k] < 0) for (int i = 0; i < 3; ++i) v[i] = -v[i];
}
int oracle_extract_lines(const lvio2d_line_params* lp, int32_t n_scans, const int64_t* point_offset, const int32_t* point_count,
                         const double* points, const double* point_z, int32_t max_lines, int32_t* n_lines, double* out_lines,
                         double* abc, int32_t* index_range) {
    lines::LineParams P;
    P.continuous_threshold = lp->line_continuous_threshold;
    P.max_tolerance_angle = lp->line_max_tolerance_angle_deg / 180.0 * M_PI;
    P.max_dis = lp->line_max_dis; P.min_len = lp->line_min_len; P.resolution = lp->laser_resolution;
    P.w = (int)(lp->w_laser_each_scan / lp->laser_resolution + 1);
    P.h = (int)(lp->h_laser_each_scan / lp->laser_resolution + 1);
    for (int s = 0; s < n_scans; ++s) {
        std::vector<lines::P3> pts;
        const int64_t end = point_count ? point_offset[s] + point_count[s] : point_offset[s + 1];
        for (int64_t i = point_offset[s]; i < end; ++i) pts.emplace_back(points[2 * i], points[2 * i + 1], point_z ? point_z[i] : 0.0);
        const std::vector<lines::Line> L = lines::spawn_scan(P, pts);
        n_lines[s] = (int32_t)L.size();
        for (int k = 0; k < (int)L.size() && k < max_lines; ++k) {
            const size_t o = (size_t)s * max_lines + k;
            out_lines[4 * o] = L[k].p1.x; out_lines[4 * o + 1] = L[k].p1.y; out_lines[4 * o + 2] = L[k].p2.x; out_lines[4 * o + 3] = L[k].p2.y;
            abc[3 * o] = L[k].abc.x; abc[3 * o + 1] = L[k].abc.y; abc[3 * o + 2] = L[k].abc.z;
            normalise_sign(abc + 3 * o);
            index_range[2 * o] = L[k].index1; index_range[2 * o + 1] = L[k].index2;
        }
    }
    return 0;
}
// ---- LaserScan ranges -> (de-skewed) points, same argument layout as lvio2d_scan_to_points (host buffers)
int oracle_scan_to_points(int32_t n_scans, int32_t n_beams, const float* ranges, const lvio2d_scan_header* headers, int32_t deskew,
                          int32_t* point_count, double* points, double* point_z, double* point_time) {
    for (int s = 0; s < n_scans; ++s) {
        scanpts::Points P = scanpts::laser_to_point_times(headers[s], ranges + (size_t)s * n_beams, n_beams);
        if (deskew) scanpts::correct(headers[s], &P);
        point_count[s] = (int32_t)P.p.size();
        for (size_t i = 0; i < P.p.size(); ++i) {
            const size_t o = (size_t)s * n_beams + i;
            points[2 * o] = P.p[i].x; points[2 * o + 1] = P.p[i].y; point_z[o] = P.p[i].z;
            if (point_time) point_time[o] = P.t[i];
        }
    }
    return 0;
}
// ---- laser_manager::do_match for a batch of scan pairs, same argument layout as lvio2d_match_lines (host buffers)
int oracle_match_lines(const lvio2d_params* prm, const lvio2d_line_params* lp, int32_t n_pairs, int32_t kk, const int64_t* point_offset1,
                       const int32_t* point_count1, const double* points1, int32_t max_lines1, const int32_t* n_lines1, const double* lines1,
                       const int32_t* index_range1, int32_t max_lines2, const int32_t* n_lines2, const double* lines2, const double* pose1,
                       const double* pose2, int32_t* n_match, int32_t* out) {
    Params PR(*prm);
    lines::LineParams P;
    P.continuous_threshold = lp->line_continuous_threshold;
    P.max_tolerance_angle = lp->line_max_tolerance_angle_deg / 180.0 * M_PI;
    P.max_dis = lp->line_max_dis; P.min_len = lp->line_min_len; P.resolution = lp->laser_resolution;
    P.w = (int)(lp->w_laser_each_scan / lp->laser_resolution + 1);
    P.h = (int)(lp->h_laser_each_scan / lp->laser_resolution + 1);
    for (int p = 0; p < n_pairs; ++p) {
        std::vector<match::Seg> L1, L2;
        for (int j = 0; j < std::min(n_lines1[p], max_lines1); ++j) {
            const double* l = lines1 + ((size_t)p * max_lines1 + j) * 4;
            match::Seg s{lines::P3(l[0], l[1], 0.0), lines::P3(l[2], l[3], 0.0), 0, -1};
            if (index_range1) { s.index1 = index_range1[((size_t)p * max_lines1 + j) * 2]; s.index2 = index_range1[((size_t)p * max_lines1 + j) * 2 + 1]; }
            L1.push_back(s);
        }
        for (int i = 0; i < std::min(n_lines2[p], max_lines2); ++i) {
            const double* l = lines2 + ((size_t)p * max_lines2 + i) * 4;
            L2.push_back(match::Seg{lines::P3(l[0], l[1], 0.0), lines::P3(l[2], l[3], 0.0), 0, -1});
        }
        std::vector<lines::P3> pts;
        if (points1) {
            const int64_t b = point_offset1[p], e = point_count1 ? b + point_count1[p] : point_offset1[p + 1];
            for (int64_t i = b; i < e; ++i) pts.emplace_back(points1[2 * i], points1[2 * i + 1], 0.0);
        }
        const match::Grid g = match::build_grid(P, L1, points1 ? &pts : nullptr);
        const auto pairs = match::do_match(P, PR.T_imu_to_laser, L1, g, L2, pose1 + 6 * p, pose2 + 6 * p, kk);
        n_match[p] = (int32_t)pairs.size();
        for (size_t k = 0; k < pairs.size() && (int)k < max_lines2; ++k) {
            out[((size_t)p * max_lines2 + k) * 2] = pairs[k].first;
            out[((size_t)p * max_lines2 + k) * 2 + 1] = pairs[k].second;
        }
    }
    return 0;
}
// ---- back-end pose graph (keyframe_manager::solve; SURVEY section 8f rank 4 — oracle only, no device path yet)
int oracle_eval_edge_factor(const double* tf12 /*3x4 row-major*/, double weight, const double* sqrt_info /*6x6*/, const double* pose_i,
                            const double* pose_j, double* res, double* jac) {
    posegraph::edge_factor fac{iso_from_rowmajor_3x4(tf12), weight, sqrt_info};
    autodiff<6, 12>(fac, {pose_i, pose_i + 3, pose_j, pose_j + 3}, {3, 3, 3, 3}, res, jac,
                    [](const posegraph::edge_factor& f, const Jet<12>* const* a, Jet<12>* r) { f(a[0], a[1], a[2], a[3], r); });
    return 0;
}
int oracle_pose_graph_solve(const lvio2d_params* p, int32_t n_poses, double* poses /*[K][6] in/out*/, int32_t n_edges, const int32_t* edge_index,
                            const double* edge_tf /*[E][12]*/, const double* edge_weight, const double* sqrt_info, int32_t ground_p,
                            int32_t ground_q, int32_t fixed_pose, lvio2d_summary* summary) {
    Params P(*p);
    posegraph::Graph G;
    G.P = &P; G.K = n_poses; G.Jn = sqrt_info; G.ground_p = ground_p != 0; G.ground_q = ground_q != 0; G.fixed = fixed_pose;
    for (int e = 0; e < n_edges; ++e)
        G.edges.push_back(posegraph::Edge{edge_index[2 * e], edge_index[2 * e + 1], iso_from_rowmajor_3x4(edge_tf + 12 * (size_t)e), edge_weight[e]});
    const lvio2d_summary S = posegraph::solve(G, lm_options_from(*p), poses);
    if (summary) *summary = S;
    return 0;
}
// the SVD restatement alone, for the numpy pin: smallest right singular vector of [x y 1]
void oracle_fit_line(const double* points, int32_t n, double* abc) {
    std::vector<lines::P3> pts;
    for (int i = 0; i < n; ++i) pts.emplace_back(points[2 * i], points[2 * i + 1], 0.0);
    const lines::P3 v = lines::fit_line_by_least_square(pts, 0, n - 1);
    abc[0] = v.x; abc[1] = v.y; abc[2] = v.z;
    normalise_sign(abc);
}

int oracle_max_threads() {
    const unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

}  // extern "C"
