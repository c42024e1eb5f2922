// lvio2d_laser_shim.h — reference-side binding of the laser front-end entry points of the C ABI (include/lvio2d.h).
//
// Three free functions with the bodies a maintainer would put behind the reference's own member functions:
//   lvio2d_shim::laser_to_points   -> body of sensor::laser::laser + sensor::laser::correct
//                                     (reference src/trajectory/sensor.h:43-46, :51-94; src/utilies/common.cpp:4-40)
//   lvio2d_shim::spawn_scan        -> body of laser_manager::spawn_scan (src/trajectory/laser_manager.cpp:350-422)
//   lvio2d_shim::do_match          -> body of laser_manager::do_match   (src/trajectory/laser_manager.cpp:244-348)
//   lvio2d_shim::device_submap     -> the state and bodies of laser_manager::add_scan (:424-496) and match_with_ref
//                                     (:531-546): ref_submap_ptr / spawnning_ref_submap_ptr / last_add_tf / current_count
//                                     live in device memory (lvio2d_submap_*), one kernel per call
// match_with_front / match_with_back / pop_scan (the key_frame deque) are unchanged host code and keep calling do_match.
//
// NOT compiled in this repository: it needs the reference's own headers (Eigen, ROS messages, laser_type.h), which are
// absent from the build image.  `lvio2d_b200.frontend.{Laser, LaserManager}` is the same logic in Python and is what
// tests/test_frontend.py exercises against the CPU oracle.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <vector>

#include <sensor_msgs/LaserScan.h>

#include "lvio2d.h"
#include "trajectory/laser_type.h"  // line, scan, laser_match
#include "utilies/params.h"         // PARAM()

namespace lvio2d_shim
{
    inline void check(int rc, const char *what)
    {
        if (rc != LVIO2D_OK)
            throw std::runtime_error(std::string(what) + ": " + lvio2d_strerror(rc));
    }

    inline lvio2d_line_params line_params()
    {
        lvio2d_line_params lp;
        lp.line_continuous_threshold = PARAM(line_continuous_threshold);
        lp.line_max_tolerance_angle_deg = PARAM(line_max_tolerance_angle);
        lp.line_max_dis = PARAM(line_max_dis);
        lp.line_min_len = PARAM(line_min_len);
        lp.laser_resolution = PARAM(laser_resolution);
        lp.w_laser_each_scan = PARAM(w_laser_each_scan);
        lp.h_laser_each_scan = PARAM(h_laser_each_scan);
        return lp;
    }

    // sensor::laser::laser(msg) followed (when `deskew`) by correct(linear, angular): fills points_ptr / times_ptr
    inline void laser_to_points(lvio2d_ctx *ctx, const sensor_msgs::LaserScan::ConstPtr &msg, bool deskew,
                                const Eigen::Vector3d &linear, const Eigen::Vector3d &angular,
                                std::vector<Eigen::Vector3d> &points, std::vector<double> &times)
    {
        const int nb = (int)msg->ranges.size();
        lvio2d_scan_header h;
        h.angle_min = msg->angle_min;
        h.angle_increment = msg->angle_increment;
        h.time_increment = msg->time_increment;
        h.reserved = 0.f;
        h.stamp = msg->header.stamp.toSec();
        for (int k = 0; k < 3; k++)
        {
            h.linear[k] = linear(k);
            h.angular[k] = angular(k);
        }
        int32_t count = 0;
        std::vector<double> xy(2 * (size_t)nb), z(nb), t(nb);
        check(lvio2d_scan_to_points(ctx, 1, nb, msg->ranges.data(), &h, deskew ? 1 : 0, &count, xy.data(), z.data(), t.data(), 0),
              "lvio2d_scan_to_points");
        points.resize(count);
        times.assign(t.begin(), t.begin() + count);
        for (int i = 0; i < count; i++)
            points[i] = Eigen::Vector3d(xy[2 * i], xy[2 * i + 1], z[i]);
    }

    // laser_manager::spawn_scan: scan::lines from the device; line_map / concers (used by the back-end only) are rebuilt on
    // the host for the accepted lines by scan::add_line(points, index1, index2)
    inline lvio_2d::scan::ptr spawn_scan(lvio2d_ctx *ctx, const std::vector<Eigen::Vector3d> &pts, double time, int w, int h, double resolution)
    {
        const lvio2d_line_params lp = line_params();
        std::vector<double> xy(2 * pts.size()), z(pts.size());
        for (size_t i = 0; i < pts.size(); i++)
        {
            xy[2 * i] = pts[i](0);
            xy[2 * i + 1] = pts[i](1);
            z[i] = pts[i](2);
        }
        const int64_t off[2] = {0, (int64_t)pts.size()};
        const int cap = 512;
        int32_t n = 0;
        std::vector<double> seg(4 * cap), abc(3 * cap);
        std::vector<int32_t> rng(2 * cap);
        check(lvio2d_extract_lines(ctx, &lp, 1, off, nullptr, xy.data(), z.data(), cap, &n, seg.data(), abc.data(), rng.data(), 0),
              "lvio2d_extract_lines");
        lvio_2d::scan::ptr s = std::make_shared<lvio_2d::scan>(w, h, resolution, time);
        for (int k = 0; k < std::min<int>(n, cap); k++)
            s->add_line(pts, rng[2 * k], rng[2 * k + 1]);
        return s;
    }

    // laser_manager::do_match.  `pts1` = the points scan1 was spawned from (empty for a sub-map built by
    // add_line(p1, p2, false): the device then rasterises the segments every 0.05 m like the reference does),
    // `rng1` = (index1, index2) of every line of scan1 into pts1.
    inline lvio_2d::laser_match::ptr do_match(lvio2d_ctx *ctx, const lvio_2d::scan::ptr &scan1, const std::vector<Eigen::Vector3d> &pts1,
                                              const std::vector<int32_t> &rng1, const lvio_2d::scan::ptr &scan2,
                                              const Eigen::Vector3d &p1, const Eigen::Vector3d &q1, const Eigen::Vector3d &p2,
                                              const Eigen::Vector3d &q2, int kk = 0)
    {
        const lvio2d_line_params lp = line_params();
        const int m1 = std::max<int>(1, scan1->lines.size()), m2 = std::max<int>(1, scan2->lines.size());
        std::vector<double> l1(4 * (size_t)m1), l2(4 * (size_t)m2), xy(2 * pts1.size());
        for (size_t j = 0; j < scan1->lines.size(); j++)
        {
            l1[4 * j] = scan1->lines[j]->p1(0); l1[4 * j + 1] = scan1->lines[j]->p1(1);
            l1[4 * j + 2] = scan1->lines[j]->p2(0); l1[4 * j + 3] = scan1->lines[j]->p2(1);
        }
        for (size_t i = 0; i < scan2->lines.size(); i++)
        {
            l2[4 * i] = scan2->lines[i]->p1(0); l2[4 * i + 1] = scan2->lines[i]->p1(1);
            l2[4 * i + 2] = scan2->lines[i]->p2(0); l2[4 * i + 3] = scan2->lines[i]->p2(1);
        }
        for (size_t i = 0; i < pts1.size(); i++)
        {
            xy[2 * i] = pts1[i](0);
            xy[2 * i + 1] = pts1[i](1);
        }
        const int32_t n1 = (int32_t)scan1->lines.size(), n2 = (int32_t)scan2->lines.size();
        const int64_t off[2] = {0, (int64_t)pts1.size()};
        const double pose1[6] = {p1(0), p1(1), p1(2), q1(0), q1(1), q1(2)}, pose2[6] = {p2(0), p2(1), p2(2), q2(0), q2(1), q2(2)};
        int32_t n_match = 0;
        std::vector<int32_t> pairs(2 * (size_t)m2);
        const bool own_points = !pts1.empty();
        check(lvio2d_match_lines(ctx, &lp, 1, kk, own_points ? off : nullptr, nullptr, own_points ? xy.data() : nullptr, m1, &n1, l1.data(),
                                 own_points ? rng1.data() : nullptr, m2, &n2, l2.data(), pose1, pose2, &n_match, pairs.data(), 0),
              "lvio2d_match_lines");
        lvio_2d::laser_match::ptr ret(new lvio_2d::laser_match);
        ret->p1 = p1; ret->q1 = q1; ret->p2 = p2; ret->q2 = q2;
        ret->scan2 = scan2;
        for (int k = 0; k < n_match; k++)
        {
            ret->lines1.push_back(scan1->lines[pairs[2 * k]]);
            ret->lines2.push_back(scan2->lines[pairs[2 * k + 1]]);
        }
        return ret;
    }

    // laser_manager's reference sub-map on the device.  In laser_manager: a member `lvio2d_shim::device_submap ref_;`,
    // add_scan(scan_ptr, p, q) { key_frame.push_back(...); ref_.add_scan(scan_ptr, p, q); } and
    // match_with_ref(scan_ptr, p, q) { return ref_.match_with_ref(scan_ptr, p, q); }.
    class device_submap
    {
    public:
        device_submap(lvio2d_ctx *ctx, int line_cap = 16384) : ctx_(ctx)
        {
            const lvio2d_line_params lp = line_params();
            check(lvio2d_submap_create(ctx, 1, line_cap, &lp, PARAM(ref_motion_filter_p), PARAM(ref_motion_filter_q),
                                       PARAM(ref_n_accumulation), &sm_),
                  "lvio2d_submap_create");
        }
        ~device_submap() { lvio2d_submap_destroy(sm_); }
        device_submap(const device_submap &) = delete;
        device_submap &operator=(const device_submap &) = delete;

        void add_scan(const lvio_2d::scan::ptr &scan_ptr, const Eigen::Vector3d &p, const Eigen::Vector3d &q)
        {
            std::vector<double> l;
            const int32_t n = pack(scan_ptr, l);
            const double pose[6] = {p(0), p(1), p(2), q(0), q(1), q(2)};
            check(lvio2d_submap_add_scan(sm_, std::max<int32_t>(1, n), &n, l.data(), pose, 0), "lvio2d_submap_add_scan");
        }

        lvio_2d::laser_match::ptr match_with_ref(const lvio_2d::scan::ptr &scan_ptr, const Eigen::Vector3d &p, const Eigen::Vector3d &q)
        {
            lvio_2d::laser_match::ptr ret(new lvio_2d::laser_match);
            ret->p1 = p; ret->p2 = p; ret->q1 = q; ret->q2 = q;
            ret->scan2 = scan_ptr;
            int32_t meta[4] = {0, 0, 0, 0};
            check(lvio2d_submap_get(sm_, 0, meta, nullptr, nullptr, nullptr), "lvio2d_submap_get");
            if (!meta[0])
                return ret; // ref_submap_ptr == nullptr
            std::vector<double> l2;
            const int32_t n2 = pack(scan_ptr, l2), m2 = std::max<int32_t>(1, n2);
            const double pose2[6] = {p(0), p(1), p(2), q(0), q(1), q(2)};
            int32_t n_match = 0;
            std::vector<int32_t> pairs(2 * (size_t)m2);
            std::vector<double> l1(4 * (size_t)m2);
            double pose1[6];
            check(lvio2d_submap_match(sm_, 0, m2, &n2, l2.data(), pose2, &n_match, pairs.data(), l1.data(), pose1, 0), "lvio2d_submap_match");
            ret->p1 = Eigen::Vector3d(pose1[0], pose1[1], pose1[2]);
            ret->q1 = Eigen::Vector3d(pose1[3], pose1[4], pose1[5]);
            for (int k = 0; k < n_match; k++)
            {
                // laser_match only reads p1 / p2 of lines1 (laser_factor's constructor)
                const Eigen::Vector3d a(l1[4 * k], l1[4 * k + 1], 0), b(l1[4 * k + 2], l1[4 * k + 3], 0);
                const Eigen::Vector2d nrm = Eigen::Vector2d(a(1) - b(1), b(0) - a(0)).normalized();
                ret->lines1.push_back(std::make_shared<lvio_2d::line>(a, b, Eigen::Vector3d(nrm(0), nrm(1), -nrm.dot(a.head<2>()))));
                ret->lines2.push_back(scan_ptr->lines[pairs[2 * k + 1]]);
            }
            return ret;
        }

    private:
        static int32_t pack(const lvio_2d::scan::ptr &s, std::vector<double> &l)
        {
            l.assign(4 * std::max<size_t>(1, s->lines.size()), 0.0);
            for (size_t j = 0; j < s->lines.size(); j++)
            {
                l[4 * j] = s->lines[j]->p1(0); l[4 * j + 1] = s->lines[j]->p1(1);
                l[4 * j + 2] = s->lines[j]->p2(0); l[4 * j + 3] = s->lines[j]->p2(1);
            }
            return (int32_t)s->lines.size();
        }
        lvio2d_ctx *ctx_;
        lvio2d_submap *sm_ = nullptr;
    };
} // namespace lvio2d_shim
