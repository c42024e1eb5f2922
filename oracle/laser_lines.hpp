// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED BY THE REFERENCE (it ships no tests);
// the one third-party routine on this path (Eigen::JacobiSVD) is pinned against numpy.linalg.svd in
// tests/test_oracle_lines.py.
//
// laser_lines.hpp — sequential CPU restatement of the scan -> line segments step that feeds the solver
// (SURVEY.md §8f rank 1):
//   laser_manager::spawn_scan            src/trajectory/laser_manager.cpp:350-422
//   scan::add_line (fit + filters)       src/trajectory/laser_manager.cpp:137-154, :155-213 (a line enters `lines` only
//                                        when one of its points falls on a valid grid cell)
//   fit_line_by_least_square             src/trajectory/laser_manager.cpp:19-36   (JacobiSVD, V.col(2))
//   create_line / project_to_line        src/trajectory/laser_manager.cpp:62-94, :8-18
//   is_continuous / clac_cos / clac_angle src/trajectory/laser_manager.cpp:96-120
//   scan::xy_to_index / is_index_valid   src/trajectory/laser_type.h:34-41
// Eigen::JacobiSVD is restated as a one-sided (Hestenes) Jacobi SVD of the n x 3 matrix [x y 1]: the same quantity
// (right singular vector of the smallest singular value) through an equally accurate route.  Its sign is arbitrary
// in Eigen too; everything downstream only uses ratios of (a, b, c).
#pragma once
#include <cmath>
#include <vector>

#include "lie.hpp"

namespace oracle {
namespace lines {

constexpr double kEpsilo = 0.0008;  // laser_manager.cpp:3

struct LineParams {
    double continuous_threshold;  // line_continuous_threshold   config/corridor.yaml:84
    double max_tolerance_angle;   // radians (line_max_tolerance_angle: 175 deg, :89; angle_to_rad common.h:48)
    double max_dis;               // line_max_dis                :86
    double min_len;               // line_min_len                :85
    double resolution;            // laser_resolution            :81
    int w, h;                     // w_laser_each_scan / resolution + 1 (laser_manager.cpp:231-236)
};

struct Line {
    Vec3<double> p1, p2, abc;
    int index1, index2;
};

using P3 = Vec3<double>;

inline P3 project_to_line(const P3& p, const P3& s, const P3& e) {
    if (norm(e - s) < kEpsilo) return p;
    const P3 u = normalized(e - s);
    const double t = dot(p - s, u);
    return s + t * u;
}

// smallest right singular vector of [x y 1] rows index1..index2
inline P3 fit_line_by_least_square(const std::vector<P3>& pts, int index1, int index2) {
    const int n = index2 - index1 + 1;
    std::vector<double> A(3 * (size_t)n);
    for (int i = 0; i < n; ++i) { A[3 * i] = pts[index1 + i].x; A[3 * i + 1] = pts[index1 + i].y; A[3 * i + 2] = 1.0; }
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < n; ++i) { alpha += A[3 * i + p] * A[3 * i + p]; beta += A[3 * i + q] * A[3 * i + q]; gamma += A[3 * i + p] * A[3 * i + q]; }
                if (std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta) || gamma == 0.0) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < n; ++i) {
                    const double ap = A[3 * i + p], aq = A[3 * i + q];
                    A[3 * i + p] = c * ap - s * aq;
                    A[3 * i + q] = s * ap + c * aq;
                }
                for (int i = 0; i < 3; ++i) {
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq;
                    V[i][q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double best = -1;
    int k = 0;
    for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += A[3 * i + j] * A[3 * i + j];
        if (best < 0 || s < best) { best = s; k = j; }
    }
    return P3(V[0][k], V[1][k], V[2][k]);
}

inline void create_line(const std::vector<P3>& pts, const P3& abc, int index1, int index2, P3* p1, P3* p2, double* max_dis) {
    P3 a(0, 0, 0), b(0, 0, 0);
    if (std::fabs(abc.y) < 0.5) {
        a.y = 0; a.x = -abc.z / abc.x;
        b.y = 1; b.x = (-abc.z - abc.y) / abc.x;
    } else {
        a.x = 0; b.x = 1;
        a.y = -abc.z / abc.y;
        b.y = (-abc.z - abc.x) / abc.y;
    }
    double m = 0;
    for (int i = index1; i <= index2; ++i) {
        const double e = e_laser::dis_from_line<double>(pts[i], a, b);
        if (e > m) m = e;
    }
    *p1 = project_to_line(pts[index1], a, b);
    *p2 = project_to_line(pts[index2], a, b);
    *max_dis = m;
}

inline double clac_cos(const P3& pj, const P3& pi, const P3& pk) {
    if (norm(pi - pj) < kEpsilo) return -1;
    if (norm(pj - pk) < kEpsilo) return -1;
    return dot(normalized(pi - pj), normalized(pk - pj));
}

// scan::add_line up to the decision whether the line enters scan::lines
inline void add_line(const LineParams& P, const std::vector<P3>& pts, int index1, int index2, std::vector<Line>* out) {
    if (index2 - index1 < 2) return;
    Line l;
    l.abc = fit_line_by_least_square(pts, index1, index2);
    double err;
    create_line(pts, l.abc, index1, index2, &l.p1, &l.p2, &err);
    l.index1 = index1; l.index2 = index2;
    const double len = norm(l.p1 - l.p2);
    if (err > P.max_dis) return;
    if (len < P.min_len) return;
    for (int i = index1; i <= index2; ++i) {
        const int c = (int)(pts[i].x / P.resolution + P.w / 2), r = (int)(pts[i].y / P.resolution + P.h / 2);
        if (r >= 0 && r < P.h && c >= 0 && c < P.w) { out->push_back(l); return; }
    }
}

// laser_manager::spawn_scan: the ordered list scan::lines
inline std::vector<Line> spawn_scan(const LineParams& P, const std::vector<P3>& pts) {
    std::vector<Line> out;
    const int n = (int)pts.size();
    if (n == 0) return out;
    std::vector<std::pair<int, int>> segs;
    {
        int start = 0;
        for (int i = 1; i < n; ++i)
            if (!(norm(pts[i - 1] - pts[i]) <= P.continuous_threshold)) { segs.emplace_back(start, i - 1); start = i; }
        segs.emplace_back(start, n - 1);
    }
    const int step = 3;
    std::vector<double> resp((size_t)n, -1.0);
    for (const auto& se : segs) {
        const int s = se.first, e = se.second;
        for (int i = s + 1; i <= e - 1; ++i) resp[i] = clac_cos(pts[i], pts[std::max(i - step, s)], pts[std::min(i + step, e)]);
        std::vector<int> maybe;
        maybe.push_back(s);
        for (int i = s + 1; i <= e - 1; ++i) {
            bool is_max = true;
            const int bj = std::max(i - step, s + 1), ej = std::min(i + step, e - 1);
            for (int j = bj; j <= ej; ++j)
                if (resp[j] >= resp[i] && j != i) { is_max = false; break; }
            if (is_max) { maybe.push_back(i); i += step; }
        }
        maybe.push_back(e);
        int last = 0;
        for (int i = 1; i + 1 < (int)maybe.size(); ++i) {
            const double angle = std::acos(clac_cos(pts[maybe[i]], pts[maybe[last]], pts[maybe[i + 1]]));
            if (std::fabs(angle) < P.max_tolerance_angle) { add_line(P, pts, maybe[last], maybe[i], &out); last = i; }
        }
        add_line(P, pts, maybe[last], maybe.back(), &out);
    }
    return out;
}

}  // namespace lines
}  // namespace oracle
