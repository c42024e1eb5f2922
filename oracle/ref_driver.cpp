// ref_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the REFERENCE'S OWN, UNMODIFIED sources, compiled where
// they lie under /root/reference/src against the stub include tree oracle/ref_stub/ (Eigen / Ceres / ROS are absent
// from this image).  Built by `make -C oracle ref` into oracle/_ref/libref.so (git-ignored, travels to the GPU box).
//
// What executes reference text here:
//   src/utilies/common.h            e_laser::dis_from_line, lie::*, auto_diff::compute_res_and_jacobi
//   src/utilies/common.cpp          convert::laser_to_point_times
//   src/trajectory/sensor.h         sensor::laser::correct
//   src/factor/{laser,imu,wheel,ground,marginalization,edge}_factor.h, factor_common.h (so3_parameterization)
//   src/factor/imu_preintegraption.h, wheel_odom_preintegration.h
//   src/factor/solver.cpp           solver::solve / init_solve / marginalization, marginalization_matrix
//   src/trajectory/laser_manager.cpp  spawn_scan, scan::add_line, do_match, add_scan, match_with_ref
// What does NOT: param::manager (params.cpp needs a ROS parameter server — the values arrive through ref_set_params),
// feature_manger (camera path, `enable_camera: false` in every shipped config — an empty feature set), and the three
// third-party libraries, restated in oracle/ref_stub (see the header of each stub).
//
// Nothing in the product path may include, link or load this file (tests/test_abi_cpu.py checks).
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include <Eigen/Dense>
#include <ceres/ceres.h>

// the reference keeps the linearised prior, the sub-maps and the preintegrators' state private; the tests read them
#define private public
#include "factor/edge_factor.h"
#include "factor/solver.h"
#include "trajectory/laser_manager.h"
#undef private

using namespace lvio_2d;

// defined at file scope in the reference (src/factor/solver.cpp:4-40), not declared in any header
std::tuple<Eigen::MatrixXd, Eigen::VectorXd> marginalization_matrix(const int& r_len, const Eigen::MatrixXd& J, const Eigen::VectorXd& R);

// ---------------------------------------------------------------------------------------------------------------
// param::manager without ROS (reference src/utilies/params.cpp:86-190 reads the same fields from the parameter server)
namespace lvio_2d {
namespace param {
manager::manager(const ros::NodeHandle&) {}
bool manager::check_param() { return true; }
manager::ptr manager::get_param_manager() {
    static manager::ptr m(new manager(ros::NodeHandle()));
    return m;
}
}  // namespace param

// camera path: an empty feature set (camera_manager.cpp is not compiled)
feature_manger::feature_manger() { feature_map_ptr = std::make_shared<feature_map>(); }
std::map<long long, feature_info::ptr>& feature_manger::get_feature_infos() { return feature_infos; }
std::vector<feature_info::ptr>& feature_manger::get_lastest_frame_features() { return lastest_frame_features; }
}  // namespace lvio_2d

namespace {
Eigen::Isometry3d iso_from_3x4(const double* m) {   // row-major [R | t]
    Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) T.matrix()(r, c) = m[r * 4 + c];
    return T;
}
Eigen::Vector3d v3(const double* p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
void put3(double* o, const Eigen::Vector3d& v) { o[0] = v(0); o[1] = v(1); o[2] = v(2); }
bool g_params_set = false;

imu_preint_result::ptr imu_from_blob(const double* b) {
    Eigen::Matrix<double, 15, 1> X;
    Eigen::Matrix<double, 15, 15> J, S;
    for (int i = 0; i < 15; ++i) X(i) = b[i];
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) { J(i, j) = b[15 + i * 15 + j]; S(i, j) = b[240 + i * 15 + j]; }
    return imu_preint_result::create(X, J, S, b[465]);
}
wheel_odom_preint_result::ptr wheel_from_blob(const double* b) {
    Eigen::Matrix3d S = Eigen::Matrix3d::Zero();
    for (int i = 0; i < 3; ++i) S(i, i) = b[12 + i];
    return wheel_odom_preint_result::create(iso_from_3x4(b), S, 0.0);
}
}  // namespace

extern "C" {

// ---- parameters: v[] in the order documented in tests/ref_lib.py (PARAM_LAYOUT)
int ref_set_params(const double* v, int n) {
    if (n < 70) return -1;
    auto P = param::manager::get_param_manager();
    int k = 0;
    P->T_imu_to_laser = iso_from_3x4(v + k); k += 12;
    P->T_imu_to_wheel = iso_from_3x4(v + k); k += 12;
    P->T_imu_to_camera = Eigen::Isometry3d::Identity();
    P->g = v[k++];
    P->line_to_line_sigma = v[k++];
    P->manifold_p_sigma = v[k++];
    P->manifold_q_sigma = v[k++];
    P->imu_noise_acc_sigma = v3(v + k); k += 3;
    P->imu_bias_acc_sigma = v3(v + k); k += 3;
    P->imu_noise_gyro_sigma = v3(v + k); k += 3;
    P->imu_bias_gyro_sigma = v3(v + k); k += 3;
    P->wheel_sigma = v3(v + k); k += 3;
    P->loop_sigma_p = v3(v + k); k += 3;
    P->loop_sigma_q = v3(v + k); k += 3;
    P->fast_mode = v[k++] != 0.0;
    P->w_laser_each_scan = v[k++];
    P->h_laser_each_scan = v[k++];
    P->laser_resolution = v[k++];
    P->line_continuous_threshold = v[k++];
    P->line_max_tolerance_angle = v[k++];
    P->line_min_len = v[k++];
    P->line_max_dis = v[k++];
    P->ref_motion_filter_p = v[k++];
    P->ref_motion_filter_q = v[k++];
    P->ref_n_accumulation = (int)v[k++];
    P->loop_edge_k = v[k++];
    P->enable_laser_vis = false;
    P->enable_camera = false;
    P->enable_laser = true;
    P->camera_sigma = Eigen::Vector2d(1.0, 1.0);
    P->camera_K = Eigen::Matrix3d::Identity();
    P->slide_window_size = 10;
    g_params_set = true;
    return k;
}
// fast_mode is read on every solver call (solver.cpp:259, :744, :791, :800), so it may change between calls; every
// other value is cached by the reference's noise singletons on first use
// test knob of the Ceres stub: cap every ceres::Solve at `cap` iterations (0 = what the reference asks for)
int ref_set_iteration_cap(int cap) { ceres::stub_detail::max_iterations_cap() = cap; return 0; }
int ref_set_fast_mode(int on) { param::manager::get_param_manager()->fast_mode = on != 0; return 0; }

// ---- primitives (src/utilies/common.h)
double ref_dis_from_line(const double* p, const double* p1, const double* p2) { return e_laser::dis_from_line<double>(v3(p), v3(p1), v3(p2)); }
void ref_exp_so3(const double* so3, double* R9) {
    const Eigen::Matrix3d R = lie::exp_so3<double>(v3(so3));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[i * 3 + j] = R(i, j);
}
void ref_log_SO3(const double* R9, double* so3) {
    Eigen::Matrix3d R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R9[i * 3 + j];
    put3(so3, lie::log_SO3<double>(R));
}
void ref_normalize_so3(double* so3) { Eigen::Vector3d v = v3(so3); lie::normalize_so3<double>(v); put3(so3, v); }
// so3_parameterization (factor_common.h:40-53) through the Ceres interface the solver uses
void ref_so3_plus(const double* theta, const double* delta, double* out, double* jac9) {
    std::unique_ptr<ceres::LocalParameterization> lp(factor::so3_parameterization::Create());
    lp->Plus(theta, delta, out);
    if (jac9) lp->ComputeJacobian(theta, jac9);
}

// ---- factors through the reference's own hook auto_diff::compute_res_and_jacobi (common.h:201-217)
// jac: row-major [n_res][sum of block sizes], blocks side by side in argument order
int ref_eval_laser_factor(const double* l1_p1, const double* l1_p2, const double* l2_p1, const double* l2_p2, const double* pose_i,
                          const double* pose_j, double* res, double* jac) {
    double x[12];
    std::memcpy(x, pose_i, 6 * sizeof(double));
    std::memcpy(x + 6, pose_j, 6 * sizeof(double));
    std::vector<double*> params = {x, x + 3, x + 6, x + 9};
    Eigen::Matrix<double, 2, 1> r;
    std::vector<auto_diff::rMatrix> J;
    auto_diff::compute_res_and_jacobi<laser_factor, 2, 3, 3, 3, 3>(new laser_factor(v3(l1_p1), v3(l1_p2), v3(l2_p1), v3(l2_p2)), params, r, J);
    for (int k = 0; k < 2; ++k) { res[k] = r(k); for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) jac[k * 12 + b * 3 + c] = J[b](k, c); }
    return 0;
}
double ref_laser_pair_weight(const double* l1_p1, const double* l1_p2, const double* l2_p1, const double* l2_p2) {
    return laser_factor(v3(l1_p1), v3(l1_p2), v3(l2_p1), v3(l2_p2)).sum;
}
// states are [p q v bs] (15); the functor's blocks are (p, q, v, bs) x 2
int ref_eval_imu_factor(const double* blob, const double* state_i, const double* state_j, double* res, double* jac) {
    double x[30];
    std::memcpy(x, state_i, 15 * sizeof(double));
    std::memcpy(x + 15, state_j, 15 * sizeof(double));
    std::vector<double*> params = {x, x + 3, x + 6, x + 9, x + 15, x + 18, x + 21, x + 24};
    Eigen::Matrix<double, 15, 1> r;
    std::vector<auto_diff::rMatrix> J;
    auto_diff::compute_res_and_jacobi<imu_factor, 15, 3, 3, 3, 6, 3, 3, 3, 6>(new imu_factor(imu_from_blob(blob)), params, r, J);
    const int off[8] = {0, 3, 6, 9, 15, 18, 21, 24}, len[8] = {3, 3, 3, 6, 3, 3, 3, 6};
    for (int k = 0; k < 15; ++k) { res[k] = r(k); for (int b = 0; b < 8; ++b) for (int c = 0; c < len[b]; ++c) jac[k * 30 + off[b] + c] = J[b](k, c); }
    return 0;
}
int ref_eval_wheel_factor(const double* blob, const double* pose_i, const double* pose_j, double* res, double* jac) {
    double x[12];
    std::memcpy(x, pose_i, 6 * sizeof(double));
    std::memcpy(x + 6, pose_j, 6 * sizeof(double));
    std::vector<double*> params = {x, x + 3, x + 6, x + 9};
    Eigen::Matrix<double, 3, 1> r;
    std::vector<auto_diff::rMatrix> J;
    auto_diff::compute_res_and_jacobi<wheel_odom_factor, 3, 3, 3, 3, 3>(new wheel_odom_factor(wheel_from_blob(blob)), params, r, J);
    for (int k = 0; k < 3; ++k) { res[k] = r(k); for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) jac[k * 12 + b * 3 + c] = J[b](k, c); }
    return 0;
}
int ref_eval_ground_factors(const double* pose, double* res, double* jac) {
    double x[6];
    std::memcpy(x, pose, sizeof(x));
    std::vector<double*> params = {x, x + 3};
    Eigen::Matrix<double, 1, 1> r;
    std::vector<auto_diff::rMatrix> J;
    auto_diff::compute_res_and_jacobi<ground_factor_p, 1, 3, 3>(new ground_factor_p, params, r, J);
    res[0] = r(0);
    for (int b = 0; b < 2; ++b) for (int c = 0; c < 3; ++c) jac[b * 3 + c] = J[b](0, c);
    auto_diff::compute_res_and_jacobi<ground_factor_q, 1, 3, 3>(new ground_factor_q, params, r, J);
    res[1] = r(0);
    for (int b = 0; b < 2; ++b) for (int c = 0; c < 3; ++c) jac[6 + b * 3 + c] = J[b](0, c);
    return 0;
}
// marginalization_factor (marginalization_factor.h:22-53) through compute_res_and_jacobi_dynamic, as solver::clac_prior_J does
int ref_eval_prior_factor(const double* X0, const double* J_lin, const double* state, double* res, double* jac) {
    Eigen::VectorXd X(15), r0 = Eigen::VectorXd::Zero(15);
    Eigen::MatrixXd Jm(15, 15);
    for (int i = 0; i < 15; ++i) { X(i) = X0[i]; for (int j = 0; j < 15; ++j) Jm(i, j) = J_lin[i * 15 + j]; }
    double x[15];
    std::memcpy(x, state, sizeof(x));
    std::vector<double*> params = {x, x + 3, x + 6, x + 9};
    std::vector<int> n_block = {3, 3, 3, 6};
    Eigen::VectorXd r(15);
    std::vector<auto_diff::rMatrix> J;
    auto_diff::compute_res_and_jacobi_dynamic(new marginalization_factor(X, Jm, r0, 0), params, n_block, r, 15, J);
    const int off[4] = {0, 3, 6, 9};
    for (int k = 0; k < 15; ++k) { res[k] = r(k); for (int b = 0; b < 4; ++b) for (int c = 0; c < n_block[b]; ++c) jac[k * 15 + off[b] + c] = J[b](k, c); }
    return 0;
}
// edge_factor (edge_factor.h:79-126) with the reference's edge_noise (including its J(1,2) assignment, edge_factor.h:21)
int ref_eval_edge_factor(const double* tf12, double weight, const double* pose_i, const double* pose_j, double* res, double* jac, double* noise_J36) {
    double x[12];
    std::memcpy(x, pose_i, 6 * sizeof(double));
    std::memcpy(x + 6, pose_j, 6 * sizeof(double));
    std::vector<double*> params = {x, x + 3, x + 6, x + 9};
    Eigen::Matrix<double, 6, 1> r;
    std::vector<auto_diff::rMatrix> J;
    auto_diff::compute_res_and_jacobi<edge_factor, 6, 3, 3, 3, 3>(new edge_factor(iso_from_3x4(tf12), weight), params, r, J);
    for (int k = 0; k < 6; ++k) { res[k] = r(k); for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c) jac[k * 12 + b * 3 + c] = J[b](k, c); }
    if (noise_J36) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) noise_J36[i * 6 + j] = edge_noise::get_edge_noise()->J(i, j);
    return 0;
}

// ---- preintegration.  samples: [dt, acc(3), gyro(3)] per update, the pair (acc, gyro) being the reading that is
// integrated over dt (last_info in imu_preintegraption::update, imu_preintegraption.h:170-181)
int ref_imu_preintegrate(int n_intervals, const int64_t* sample_offset, const double* samples, const double* bias0, double* out_blobs) {
    for (int i = 0; i < n_intervals; ++i) {
        imu_preintegraption pre;
        pre.reset_imu_measure(0.0, v3(bias0 + 6 * i), v3(bias0 + 6 * i + 3));
        for (int64_t s = sample_offset[i]; s < sample_offset[i + 1]; ++s) {
            pre.last_info.acc = v3(samples + 7 * s + 1);
            pre.last_info.gyro = v3(samples + 7 * s + 4);
            pre.update(samples[7 * s]);
        }
        const imu_preint_result::ptr r = pre.get_preintegraption_result();
        double* b = out_blobs + (size_t)i * 466;
        for (int k = 0; k < 15; ++k) b[k] = r->X(k);
        for (int a = 0; a < 15; ++a) for (int c = 0; c < 15; ++c) { b[15 + a * 15 + c] = r->J(a, c); b[240 + a * 15 + c] = r->sqrt_inverse_P(a, c); }
        b[465] = r->Dt;
    }
    return 0;
}
// the same interval through the public interface with time stamps (add_imu_measure / update_only_t), as trajectory.cpp drives it
int ref_imu_preintegrate_stamped(int n_samples, const double* stamps /* [n+1] */, const double* acc_gyro /* [n][6] */, const double* bias, double* blob) {
    imu_preintegraption pre;
    pre.reset_imu_measure(-1, v3(bias), v3(bias + 3));
    for (int s = 0; s < n_samples; ++s) {
        sensor::imu::u_ptr d(new sensor::imu());
        d->acc = v3(acc_gyro + 6 * s);
        d->gyro = v3(acc_gyro + 6 * s + 3);
        d->time_stamp = stamps[s];
        pre.add_imu_measure(d);
    }
    pre.update_only_t(stamps[n_samples]);
    const imu_preint_result::ptr r = pre.get_preintegraption_result();
    for (int k = 0; k < 15; ++k) blob[k] = r->X(k);
    for (int a = 0; a < 15; ++a) for (int c = 0; c < 15; ++c) { blob[15 + a * 15 + c] = r->J(a, c); blob[240 + a * 15 + c] = r->sqrt_inverse_P(a, c); }
    blob[465] = r->Dt;
    return 0;
}
// steps: [dt, v(3), omega(3)] per update_by_v (wheel_odom_preintegration.h:141-152)
int ref_wheel_preintegrate(int n_intervals, const int64_t* step_offset, const double* steps, double* out_blobs) {
    for (int i = 0; i < n_intervals; ++i) {
        wheel_odom_preintegration pre;
        pre.reset_wheel_odom_measure(0.0);
        for (int64_t s = step_offset[i]; s < step_offset[i + 1]; ++s) {
            pre.v = v3(steps + 7 * s + 1);
            pre.omega = v3(steps + 7 * s + 4);
            pre.update_by_v(steps[7 * s]);
        }
        const wheel_odom_preint_result::ptr r = pre.get_preintegraption_result();
        double* b = out_blobs + (size_t)i * 15;
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 4; ++c) b[a * 4 + c] = r->delta_Tij.matrix()(a, c);
        for (int a = 0; a < 3; ++a) b[12 + a] = r->sqrt_inverse_P(a, a);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// lvio_2d::solver on a window given frame by frame.
//   states [n][15] in/out;  pair_offset [n+1], pairs [.][12] = l1_p1 l1_p2 l2_p1 l2_p2 (xyz each) of frame i's laser_match
//   (no match object when has_match[i] == 0);  match_pose [n][12] = p1 q1 p2 q2 in/out;  imu [n][466], wheel [n][15]
//   (row 0 unused);  which: 0 solve, 1 init_solve, 2 marginalization
//   summary[8]: iterations, termination code, successful, unsuccessful steps, initial cost, final cost, final radius, fixed cost
struct RefSolver {
    solver s;
    feature_manger features;
};
void* ref_solver_new() { return new RefSolver(); }
void ref_solver_free(void* h) { delete static_cast<RefSolver*>(h); }
int ref_solver_run(void* h, int which, int n, double* states, const int32_t* has_match, const int64_t* pair_offset, const double* pairs,
                   double* match_pose, const double* imu, const double* wheel, double* sqrt_H_last, double* summary) {
    RefSolver* R = static_cast<RefSolver*>(h);
    std::deque<frame_info::ptr> frames;
    for (int i = 0; i < n; ++i) {
        const double* x = states + 15 * i;
        Eigen::Matrix<double, 6, 1> bs;
        for (int k = 0; k < 6; ++k) bs(k) = x[9 + k];
        frame_info::ptr f = frame_info::create((double)i, v3(x), v3(x + 3), v3(x + 6), bs, i > 0 && imu ? imu_from_blob(imu + (size_t)i * 466) : nullptr,
                                               i > 0 && wheel ? wheel_from_blob(wheel + (size_t)i * 15) : nullptr);
        if (has_match && has_match[i]) {
            laser_match::ptr m(new laser_match);
            for (int64_t j = pair_offset[i]; j < pair_offset[i + 1]; ++j) {
                const double* q = pairs + 12 * j;
                m->lines1.push_back(std::make_shared<line>(v3(q), v3(q + 3), Eigen::Vector3d(0, 0, 0)));
                m->lines2.push_back(std::make_shared<line>(v3(q + 6), v3(q + 9), Eigen::Vector3d(0, 0, 0)));
            }
            m->p1 = v3(match_pose + 12 * i); m->q1 = v3(match_pose + 12 * i + 3);
            m->p2 = v3(match_pose + 12 * i + 6); m->q2 = v3(match_pose + 12 * i + 9);
            f->add_laser_match(m);
        } else {
            f->type = frame_info::laser;
        }
        frames.push_back(f);
    }
    ceres::stub_detail::last_summary() = ceres::Solver::Summary();
    if (which == 0) R->s.solve(frames, R->features);
    else if (which == 1) R->s.init_solve(frames, R->features);
    else R->s.marginalization(frames, R->features);
    for (int i = 0; i < n; ++i) {
        double* x = states + 15 * i;
        put3(x, frames[(size_t)i]->p); put3(x + 3, frames[(size_t)i]->q); put3(x + 6, frames[(size_t)i]->v);
        for (int k = 0; k < 6; ++k) x[9 + k] = frames[(size_t)i]->bs(k);
        if (frames[(size_t)i]->laser_match_ptr) {
            const laser_match::ptr& m = frames[(size_t)i]->laser_match_ptr;
            put3(match_pose + 12 * i, m->p1); put3(match_pose + 12 * i + 3, m->q1); put3(match_pose + 12 * i + 6, m->p2); put3(match_pose + 12 * i + 9, m->q2);
        }
    }
    if (sqrt_H_last) for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) sqrt_H_last[a * 6 + b] = frames.back()->sqrt_H(a, b);
    if (summary) {
        const ceres::Solver::Summary& S = ceres::stub_detail::last_summary();
        summary[0] = S.stub_iterations; summary[1] = S.stub_termination; summary[2] = S.num_successful_steps; summary[3] = S.num_unsuccessful_steps;
        summary[4] = S.initial_cost; summary[5] = S.final_cost; summary[6] = S.stub_final_radius; summary[7] = S.fixed_cost;
    }
    return 0;
}
// the linearised prior kept by solver::marginalization (solver.cpp:399-441): returns has_linearized_block
int ref_solver_get_prior(void* h, double* X0, double* J, double* r) {
    RefSolver* R = static_cast<RefSolver*>(h);
    if (!R->s.has_linearized_block) return 0;
    for (int i = 0; i < 15; ++i) { X0[i] = R->s.linearized_X(i); r[i] = R->s.linearized_residuals(i); for (int j = 0; j < 15; ++j) J[i * 15 + j] = R->s.linearized_jacobians(i, j); }
    return 1;
}
int ref_solver_set_prior(void* h, const double* X0, const double* J, const double* r) {
    RefSolver* R = static_cast<RefSolver*>(h);
    R->s.linearized_X = Eigen::VectorXd::Zero(15, 1);
    R->s.linearized_residuals = Eigen::VectorXd::Zero(15, 1);
    R->s.linearized_jacobians = Eigen::MatrixXd::Zero(15, 15);
    for (int i = 0; i < 15; ++i) { R->s.linearized_X(i) = X0[i]; R->s.linearized_residuals(i) = r[i]; for (int j = 0; j < 15; ++j) R->s.linearized_jacobians(i, j) = J[i * 15 + j]; }
    R->s.has_linearized_block = true;
    return 0;
}
// J and R assembled by the last solver::marginalization call (solver.cpp:364-381) and marginalization_matrix of them
int ref_solver_marg_system(void* h, int r_len, double* Delta_H, double* Delta_g, int* rows, int* cols) {
    RefSolver* R = static_cast<RefSolver*>(h);
    *rows = (int)R->s.J.rows(); *cols = (int)R->s.J.cols();
    if (R->s.J.rows() == 0) return -1;
    Eigen::VectorXd Rv = R->s.R;
    auto [dH, dg] = marginalization_matrix(r_len, R->s.J, Rv);
    for (int i = 0; i < r_len; ++i) { Delta_g[i] = dg(i); for (int j = 0; j < r_len; ++j) Delta_H[i * r_len + j] = dH(i, j); }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// laser front-end (src/utilies/common.cpp:4-40, src/trajectory/sensor.h:51-94, src/trajectory/laser_manager.cpp)
// header: angle_min, angle_increment, time_increment, range_min, range_max (float32 in the message), stamp, linear(3), angular(3)
int ref_scan_to_points(int n_beams, const float* ranges, const double* header12, int deskew, double* points /* [n][3] */, double* times /* [n] */) {
    sensor_msgs::LaserScan::Ptr msg(new sensor_msgs::LaserScan);
    msg->angle_min = (float)header12[0]; msg->angle_increment = (float)header12[1]; msg->time_increment = (float)header12[2];
    msg->range_min = (float)header12[3]; msg->range_max = (float)header12[4];
    msg->header.stamp = ros::Time(header12[5]);
    msg->ranges.assign(ranges, ranges + n_beams);
    sensor::laser L(msg);
    if (deskew) L.correct(v3(header12 + 6), v3(header12 + 9));
    const int n = (int)L.points_ptr->size();
    for (int i = 0; i < n; ++i) { put3(points + 3 * i, (*L.points_ptr)[(size_t)i]); times[i] = (*L.times_ptr)[(size_t)i]; }
    return n;
}

struct RefScan { scan::ptr s; };
static sensor::laser::u_ptr laser_from_points(int n, const double* pts3) {
    sensor_msgs::LaserScan::Ptr msg(new sensor_msgs::LaserScan);
    msg->angle_increment = 1.0f;
    sensor::laser::u_ptr L(new sensor::laser(msg));
    for (int i = 0; i < n; ++i) { L->points_ptr->push_back(v3(pts3 + 3 * i)); L->times_ptr->push_back(0.0); }
    return L;
}
// laser_manager::spawn_scan on a point list (z kept as given)
void* ref_scan_from_points(int n, const double* pts3) {
    laser_manager lm;
    sensor::laser::u_ptr L = laser_from_points(n, pts3);
    if (n == 0) return nullptr;
    return new RefScan{lm.spawn_scan(L)};
}
// the sub-map flavour: an empty scan filled by scan::add_line(p1, p2, false) (laser_manager.cpp:232-241, :438-440)
void* ref_scan_from_lines(int n, const double* lines6) {
    laser_manager lm;
    scan::ptr s = std::make_shared<scan>(lm.w, lm.h, lm.resolution, 0);
    for (int i = 0; i < n; ++i) s->add_line(v3(lines6 + 6 * i), v3(lines6 + 6 * i + 3), false);
    return new RefScan{s};
}
void ref_scan_free(void* h) { delete static_cast<RefScan*>(h); }
int ref_scan_num_lines(void* h) { return (int)static_cast<RefScan*>(h)->s->lines.size(); }
int ref_scan_num_corners(void* h) { return (int)static_cast<RefScan*>(h)->s->concers.size(); }
// lines: [n][9] = p1 p2 abc;  corners [m][3]
void ref_scan_get(void* h, double* lines9, double* corners3) {
    const scan::ptr& s = static_cast<RefScan*>(h)->s;
    for (size_t i = 0; i < s->lines.size(); ++i) { put3(lines9 + 9 * i, s->lines[i]->p1); put3(lines9 + 9 * i + 3, s->lines[i]->p2); put3(lines9 + 9 * i + 6, s->lines[i]->abc); }
    if (corners3) for (size_t i = 0; i < s->concers.size(); ++i) put3(corners3 + 3 * i, s->concers[i]);
}
static int index_of(const std::vector<line::ptr>& v, const line::ptr& l) {
    for (size_t i = 0; i < v.size(); ++i) if (v[i] == l) return (int)i;
    return -1;
}
// laser_manager::do_match: pairs (index into scan1->lines, index into scan2->lines); returns the pair count
int ref_do_match(void* h1, void* h2, const double* pose1, const double* pose2, int kk, int32_t* pairs) {
    const scan::ptr &s1 = static_cast<RefScan*>(h1)->s, &s2 = static_cast<RefScan*>(h2)->s;
    laser_match::ptr m = laser_manager::do_match(s1, s2, v3(pose1), v3(pose1 + 3), v3(pose2), v3(pose2 + 3), kk);
    for (size_t j = 0; j < m->lines1.size(); ++j) { pairs[2 * j] = index_of(s1->lines, m->lines1[j]); pairs[2 * j + 1] = index_of(s2->lines, m->lines2[j]); }
    return (int)m->lines1.size();
}

// laser_manager with its reference sub-map bookkeeping (add_scan / match_with_ref, laser_manager.cpp:424-496, :527-543)
struct RefLaserManager { laser_manager lm; };
void* ref_lm_new() { return new RefLaserManager(); }
void ref_lm_free(void* h) { delete static_cast<RefLaserManager*>(h); }
int ref_lm_add_scan(void* h, void* scan_h, const double* pose) {
    static_cast<RefLaserManager*>(h)->lm.add_scan(static_cast<RefScan*>(scan_h)->s, v3(pose), v3(pose + 3));
    return 0;
}
// which: 0 ref_submap_ptr, 1 spawnning_ref_submap_ptr.  Returns the line count (-1: no such sub-map); fills pose[6],
// lines [.][6] when non-null
int ref_lm_submap(void* h, int which, double* pose, double* lines6, int* current_count) {
    laser_manager& lm = static_cast<RefLaserManager*>(h)->lm;
    const laser_submap::ptr& sm = which == 0 ? lm.ref_submap_ptr : lm.spawnning_ref_submap_ptr;
    if (current_count) *current_count = lm.current_count;
    if (!sm) return -1;
    if (pose) { put3(pose, sm->current_p); put3(pose + 3, sm->current_q); }
    const auto& L = sm->scan_ptr->lines;
    if (lines6) for (size_t i = 0; i < L.size(); ++i) { put3(lines6 + 6 * i, L[i]->p1); put3(lines6 + 6 * i + 3, L[i]->p2); }
    return (int)L.size();
}
// match_with_ref: pairs as (index into the reference sub-map's lines, index into the scan's lines); ref pose out
int ref_lm_match_with_ref(void* h, void* scan_h, const double* pose, int32_t* pairs, double* ref_pose) {
    laser_manager& lm = static_cast<RefLaserManager*>(h)->lm;
    const scan::ptr& s2 = static_cast<RefScan*>(scan_h)->s;
    laser_match::ptr m = lm.match_with_ref(s2, v3(pose), v3(pose + 3));
    put3(ref_pose, m->p1); put3(ref_pose + 3, m->q1);
    if (!lm.ref_submap_ptr) return 0;
    const auto& L1 = lm.ref_submap_ptr->scan_ptr->lines;
    for (size_t j = 0; j < m->lines1.size(); ++j) { pairs[2 * j] = index_of(L1, m->lines1[j]); pairs[2 * j + 1] = index_of(s2->lines, m->lines2[j]); }
    return (int)m->lines1.size();
}

}  // extern "C"
