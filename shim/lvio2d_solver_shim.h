// lvio2d_solver_shim.h — the reference-side binding of the C ABI (include/lvio2d.h).
//
// Drop-in replacement for the reference's `lvio_2d::solver` (reference src/factor/solver.h:28-80): same class name,
// same public signatures (solver(), solve, init_solve, marginalization), same in-place write-back into the
// `std::deque<frame_info::ptr>` the trajectory manager owns (src/trajectory/trajectory.cpp:534, :544, :446, :479).
// A maintainer replaces `#include "factor/solver.h"` in src/trajectory/trajectory.h:5 by this header and links
// liblvio2d.so; nothing else in lvio_2d_node changes.
//
// NOT compiled in this repository: it needs the reference's own headers (Eigen, ROS params, frame_info), which are
// absent from the build image.  The Python class `lvio2d_b200.solver.Solver` is the same logic line for line and is
// what tests/test_gpu_parity.py::test_solver_class_mirrors_reference_flow exercises.
#pragma once
#include <cmath>
#include <deque>
#include <stdexcept>
#include <vector>

#include "lvio2d.h"
#include "trajectory/camera_manager.h"   // feature_manger (unused: camera path is off in every shipped config)
#include "trajectory/trajectory_type.h"  // frame_info, laser_match, line
#include "utilies/params.h"              // PARAM()

namespace lvio_2d
{
    class solver
    {
    private:
        lvio2d_ctx *ctx = nullptr;
        bool has_linearized_block = false;
        // the sqrt-information prior kept between frames (reference solver.h:35-37)
        Eigen::Matrix<double, 15, 1> linearized_X;
        Eigen::Matrix<double, 15, 15, Eigen::RowMajor> linearized_jacobians;
        Eigen::Matrix<double, 15, 1> linearized_residuals;

        static void copy_tf(const Eigen::Isometry3d &T, double *out12)
        {
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 4; c++)
                    out12[r * 4 + c] = T.matrix()(r, c);
        }

        // flatten frame_infos into one lvio2d_window_batch (B = 1)
        struct flat
        {
            std::vector<double> states, points, weights, lines, ref_pose, imu, wheel;
            std::vector<uint8_t> cmask;
            std::vector<int32_t> point_line, ref_frame;
            std::vector<int64_t> poff, loff;
            lvio2d_window_batch b;
        };

        enum topology { TRACKING, INIT, MARG };

        void build(std::deque<frame_info::ptr> &frames, topology topo, flat &f)
        {
            const int n = frames.size();
            f.states.resize(n * 15);
            f.cmask.assign(n, 0);
            f.ref_frame.assign(n, -1);
            f.ref_pose.assign(n * 6, 0.0);
            f.poff.assign(1, 0);
            f.loff.assign(1, 0);
            for (int i = 0; i < n; i++)
            {
                auto &fr = frames[i];
                Eigen::Map<Eigen::Matrix<double, 15, 1>> s(f.states.data() + 15 * i);
                s << fr->p, fr->q, fr->v, fr->bs;
                // which frames carry laser factors: solver.cpp:669 (newest only), :87-113 (all but frame 0), :448-478 (all)
                const bool use_laser = fr->type == frame_info::laser && fr->laser_match_ptr &&
                                       (topo == MARG || (topo == TRACKING && i == n - 1) || (topo == INIT && i > 0));
                if (use_laser)
                {
                    auto &m = fr->laser_match_ptr;
                    for (size_t j = 0; j < m->lines1.size(); j++)
                    {
                        // laser_factor::sum (laser_factor.h:38-42)
                        const double len1 = (m->lines1[j]->p1 - m->lines1[j]->p2).norm(), len2 = (m->lines2[j]->p1 - m->lines2[j]->p2).norm();
                        const double w = std::sqrt(std::min(len1, len2) / 2.0 / 0.02);
                        const Eigen::Vector3d *c[2] = {&m->lines2[j]->p1, &m->lines2[j]->p2};
                        for (int e = 0; e < 2; e++)
                        {
                            f.points.push_back((*c[e])(0));
                            f.points.push_back((*c[e])(1));
                            f.point_line.push_back(j);
                            f.weights.push_back(w);
                        }
                        f.lines.insert(f.lines.end(), {m->lines1[j]->p1(0), m->lines1[j]->p1(1), m->lines1[j]->p2(0), m->lines1[j]->p2(1)});
                    }
                    if (topo == INIT)
                        f.ref_frame[i] = 0;
                    for (int k = 0; k < 3; k++)
                    {
                        f.ref_pose[6 * i + k] = m->p1(k);
                        f.ref_pose[6 * i + 3 + k] = m->q1(k);
                    }
                }
                f.poff.push_back(f.point_line.size());
                f.loff.push_back(f.lines.size() / 4);
                if (topo == TRACKING && i < n - 1) // solver.cpp:787-794
                    f.cmask[i] = LVIO2D_CONST_P | LVIO2D_CONST_Q | (PARAM(fast_mode) ? LVIO2D_CONST_BS : 0);
                if (i > 0)
                {
                    auto &im = fr->imu_observation_reslut;
                    const size_t o = f.imu.size();
                    f.imu.resize(o + LVIO2D_IMU_BLOB);
                    double *d = f.imu.data() + o;
                    Eigen::Map<Eigen::Matrix<double, 15, 1>>(d) = im->X;
                    Eigen::Map<Eigen::Matrix<double, 15, 15, Eigen::RowMajor>>(d + 15) = im->J;
                    Eigen::Map<Eigen::Matrix<double, 15, 15, Eigen::RowMajor>>(d + 240) = im->sqrt_inverse_P;
                    d[465] = im->Dt;
                    auto &wh = fr->wheel_observation_reslut;
                    double T[12];
                    copy_tf(wh->delta_Tij, T);
                    f.wheel.insert(f.wheel.end(), T, T + 12);
                    for (int k = 0; k < 3; k++)
                        f.wheel.push_back(wh->sqrt_inverse_P(k, k));
                }
            }
            lvio2d_window_batch &b = f.b;
            b.n_windows = 1;
            b.n_frames = n;
            b.states = f.states.data();
            b.const_mask = f.cmask.data();
            b.point_offset = f.poff.data();
            b.points = f.points.data();
            b.point_line = f.point_line.data();
            b.point_weight = f.weights.data();
            b.line_offset = f.loff.data();
            b.lines = f.lines.data();
            b.ref_frame = f.ref_frame.data();
            b.ref_pose = f.ref_pose.data();
            b.imu = n > 1 ? f.imu.data() : nullptr;
            b.wheel = n > 1 ? f.wheel.data() : nullptr;
            b.ground_multiplicity = n; // solver.cpp:727-743
            const bool prior = topo != INIT && !PARAM(fast_mode) && has_linearized_block && n >= 2;
            b.prior_frame = prior ? n - 2 : -1; // solver.cpp:750
            b.prior_X0 = prior ? linearized_X.data() : nullptr;
            b.prior_J = prior ? linearized_jacobians.data() : nullptr;
        }

        void write_back(std::deque<frame_info::ptr> &frames)
        {
            std::vector<double> st(frames.size() * 15);
            check(lvio2d_get_states(ctx, st.data()), "lvio2d_get_states");
            for (size_t i = 0; i < frames.size(); i++)
            {
                Eigen::Map<Eigen::Matrix<double, 15, 1>> s(st.data() + 15 * i);
                frames[i]->p = s.segment<3>(0);
                frames[i]->q = s.segment<3>(3);
                frames[i]->v = s.segment<3>(6);
                frames[i]->bs = s.segment<6>(9);
            }
        }

        void check(int rc, const char *what)
        {
            // the reference ignores ceres::Solver::Summary (solver.cpp:799-802); errors of the device path are fatal
            if (rc != LVIO2D_OK)
                throw std::runtime_error(std::string(what) + ": " + lvio2d_last_error(ctx));
        }

    public:
        solver()
        {
            lvio2d_params p = {};
            p.abi_version = LVIO2D_ABI_VERSION;
            p.device = 0;
            copy_tf(PARAM(T_imu_to_laser), p.T_imu_to_laser);
            copy_tf(PARAM(T_imu_to_wheel), p.T_imu_to_wheel);
            p.g = PARAM(g);
            p.line_to_line_sigma = PARAM(line_to_line_sigma);
            p.manifold_p_sigma = PARAM(manifold_p_sigma);
            p.manifold_q_sigma = PARAM(manifold_q_sigma);
            for (int i = 0; i < 3; i++)
            {
                p.imu_noise_acc_sigma[i] = PARAM(imu_noise_acc_sigma)(i);
                p.imu_bias_acc_sigma[i] = PARAM(imu_bias_acc_sigma)(i);
                p.imu_noise_gyro_sigma[i] = PARAM(imu_noise_gyro_sigma)(i);
                p.imu_bias_gyro_sigma[i] = PARAM(imu_bias_gyro_sigma)(i);
                p.wheel_sigma[i] = PARAM(wheel_sigma)(i);
            }
            p.max_iters = 50; // Ceres' default; lowered per call like the reference does (solver.cpp:800-801)
            if (lvio2d_create(&ctx, &p) != LVIO2D_OK)
                throw std::runtime_error("lvio2d_create failed (no sm_100 device?)");
        }
        ~solver() { lvio2d_destroy(ctx); }
        solver(const solver &) = delete;

        // solver.cpp:631-820
        void solve(std::deque<frame_info::ptr> &frame_infos, feature_manger &)
        {
            flat f;
            build(frame_infos, TRACKING, f);
            check(lvio2d_set_windows(ctx, &f.b), "lvio2d_set_windows");
            check(lvio2d_set_max_iterations(ctx, PARAM(fast_mode) ? 10 : 50), "lvio2d_set_max_iterations"); // solver.cpp:800-801
            check(lvio2d_solve(ctx, nullptr), "lvio2d_solve");
            write_back(frame_infos);
            auto &last = frame_infos.back();
            if (last->type == frame_info::laser && last->laser_match_ptr) // solver.cpp:804-814
            {
                last->laser_match_ptr->p2 = last->p;
                last->laser_match_ptr->q2 = last->q;
            }
        }

        // solver.cpp:171-195
        void init_solve(std::deque<frame_info::ptr> &frame_infos, feature_manger &)
        {
            flat f;
            build(frame_infos, INIT, f);
            check(lvio2d_set_windows(ctx, &f.b), "lvio2d_set_windows");
            // do_init_solve never lowers max_num_iterations: 50 even in fast_mode (solver.cpp:161-168)
            check(lvio2d_set_max_iterations(ctx, 50), "lvio2d_set_max_iterations");
            check(lvio2d_solve(ctx, nullptr), "lvio2d_solve");
            write_back(frame_infos);
            for (auto &fr : frame_infos)
                if (fr->type == frame_info::laser && fr->laser_match_ptr)
                {
                    fr->laser_match_ptr->p1 = frame_infos[0]->p;
                    fr->laser_match_ptr->q1 = frame_infos[0]->q;
                    fr->laser_match_ptr->p2 = fr->p;
                    fr->laser_match_ptr->q2 = fr->q;
                }
        }

        // solver.cpp:257-442
        void marginalization(std::deque<frame_info::ptr> &frame_infos, feature_manger &)
        {
            if (PARAM(fast_mode))
                return;
            flat f;
            build(frame_infos, MARG, f);
            check(lvio2d_set_windows(ctx, &f.b), "lvio2d_set_windows");
            check(lvio2d_marginalize(ctx, linearized_X.data(), linearized_jacobians.data(), linearized_residuals.data()),
                  "lvio2d_marginalize");
            frame_infos.back()->sqrt_H = linearized_jacobians.block<6, 6>(0, 0); // solver.cpp:401
            has_linearized_block = true;
        }
    };
} // namespace lvio_2d
