// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED BY THE REFERENCE (it ships no tests).
//
// laser_match.hpp — CPU restatement of the segment association step:
//   laser_manager::do_match            src/trajectory/laser_manager.cpp:244-348
//   scan::line_map as filled by scan::add_line   :155-213  (spawn_scan flavour: the cells of a line's own points, a line
//                                      appended to a cell once; sub-map flavour add_line(p1, p2, false): cells sampled
//                                      every 0.05 m along the segment)
// The grid is kept literally (a map cell -> ordered list of line ids) so that candidate order, duplicates and the
// strict `angle < best_angle` tie rule are the reference's.
#pragma once
#include <cmath>
#include <map>
#include <utility>
#include <vector>

#include "laser_lines.hpp"

namespace oracle {
namespace match {

using lines::P3;

struct Seg { P3 p1, p2; int index1, index2; };
using Grid = std::map<std::pair<int, int>, std::vector<int>>;  // (r, c) -> line ids in insertion order

inline std::pair<int, int> xy_to_index(const lines::LineParams& P, double x, double y) {  // returns (c, r) like the reference
    return {(int)(x / P.resolution + P.w / 2), (int)(y / P.resolution + P.h / 2)};
}
inline bool valid(const lines::LineParams& P, int r, int c) { return r >= 0 && r < P.h && c >= 0 && c < P.w; }

// scan::add_line's rasterisation for every line of a scan, in line order
inline Grid build_grid(const lines::LineParams& P, const std::vector<Seg>& L, const std::vector<P3>* pts) {
    Grid g;
    for (int j = 0; j < (int)L.size(); ++j) {
        auto push = [&](double x, double y) {
            const auto cr = xy_to_index(P, x, y);
            const int c = cr.first, r = cr.second;
            if (!valid(P, r, c)) return;
            auto& cell = g[{r, c}];
            if (cell.empty() || cell.back() != j) cell.push_back(j);
        };
        if (pts) {
            for (int i = L[j].index1; i <= L[j].index2; ++i) push((*pts)[i].x, (*pts)[i].y);
        } else {
            const double len = norm(L[j].p1 - L[j].p2);
            const P3 unit = normalized(L[j].p2 - L[j].p1);
            for (double tr = 0; tr <= len; tr += 0.05) { const P3 t = L[j].p1 + unit * tr; push(t.x, t.y); }
        }
    }
    return g;
}

// returns the (line of scan1, line of scan2) pairs of ret2
inline std::vector<std::pair<int, int>> do_match(const lines::LineParams& P, const Iso3<double>& T_il, const std::vector<Seg>& L1,
                                                 const Grid& g1, const std::vector<Seg>& L2, const double* pose1, const double* pose2, int kk) {
    const Iso3<double> T1 = lie::make_tf<double>(P3(pose1[0], pose1[1], pose1[2]), P3(pose1[3], pose1[4], pose1[5])) * T_il;
    const Iso3<double> T2 = lie::make_tf<double>(P3(pose2[0], pose2[1], pose2[2]), P3(pose2[3], pose2[4], pose2[5])) * T_il;
    const Iso3<double> T12 = inverse(T1) * T2;
    auto tf = [&](const P3& p) { return T12.R * p + T12.t; };
    std::vector<std::pair<int, int>> ret;
    for (int i = 0; i < (int)L2.size(); ++i) {
        const P3 mid = (L2[i].p1 + L2[i].p2) / 2.0;
        const P3 tm = tf(mid);
        const auto cr = xy_to_index(P, tm.x, tm.y);
        const int c = cr.first, r = cr.second;
        std::vector<int> tmp;
        const int a = 1 + kk;
        for (int dr = -a; dr <= a; ++dr)
            for (int dc = -a; dc <= a; ++dc) {
                const int rr = r + dr, cc = c + dc;
                if (!valid(P, rr, cc)) continue;
                const auto it = g1.find({rr, cc});
                if (it != g1.end()) tmp.insert(tmp.end(), it->second.begin(), it->second.end());
            }
        if (tmp.empty()) continue;
        int best = -1;
        double best_angle = M_PI * 2;
        const P3 v2 = tf(L2[i].p2) - tf(L2[i].p1);
        for (int j : tmp) {
            const P3 v1 = L1[j].p2 - L1[j].p1;
            const double angle = std::acos(std::fabs(dot(normalized(v1), normalized(v2))));
            if (angle < best_angle) { best = j; best_angle = angle; }
        }
        if (best_angle / M_PI * 180.0 > 10) continue;
        ret.emplace_back(best, i);
    }
    double aver = 0;
    std::vector<double> diss(ret.size(), 0.0);
    for (size_t k = 0; k < ret.size(); ++k) {
        const Seg& l1 = L1[ret[k].first];
        const Seg& l2 = L2[ret[k].second];
        const double d = 0.5 * (e_laser::dis_from_line<double>(tf(l2.p1), l1.p1, l1.p2) + e_laser::dis_from_line<double>(tf(l2.p2), l1.p1, l1.p2));
        aver += d;
        diss[k] = d;
    }
    aver /= (double)ret.size();
    std::vector<std::pair<int, int>> ret2;
    for (size_t k = 0; k < ret.size(); ++k)
        if (diss[k] < aver * 1.2) ret2.push_back(ret[k]);
    return ret2;
}

}  // namespace match
}  // namespace oracle
