// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).  PARITY UNPINNED BY THE REFERENCE (it ships no tests).
//
// scan_points.hpp — CPU restatement of the wire-format step in front of spawn_scan:
//   convert::laser_to_point_times   src/utilies/common.cpp:4-40
//   sensor::laser::correct          src/trajectory/sensor.h:51-94   (called from trajectory.cpp:147)
// Arithmetic types follow the reference's declarations: angle_start, angle_increment, time_increment and the ranges
// are `float`; `angle_start + i * angle_increment` and `i * time_increment` are float expressions (size_t -> float);
// `cos(float)` is taken as the C function ::cos(double) (the call is unqualified inside namespace convert; whether a
// float overload is visible depends on the reference's include order, which cannot be compiled here — a float
// cosine would change the points at the 1e-7 relative level), the product with the float range is double.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "lie.hpp"
#include "../include/lvio2d.h"

namespace oracle {
namespace scanpts {

struct Points { std::vector<Vec3<double>> p; std::vector<double> t; };

inline Points laser_to_point_times(const lvio2d_scan_header& h, const float* ranges, int n_beams) {
    Points out;
    const float angle_start = h.angle_min, angle_increment = h.angle_increment, time_increment = h.time_increment;
    const double time = h.stamp;
    for (size_t i = 0; i < (size_t)n_beams; ++i) {
        const float r = ranges[i];
        if (!std::isnan(r) && !std::isinf(r) && r > 0.1) {
            volatile float step = i * angle_increment;   // keep the float product and sum separate roundings (the
            const float ang = angle_start + step;        // reference's Release build has no FMA contraction)
            const Vec3<double> point(std::cos((double)ang) * r, std::sin((double)ang) * r, 0.0);
            if (!out.p.empty() && norm(point - out.p.back()) < 0.01) continue;
            out.p.push_back(point);
            out.t.push_back(time + i * time_increment);
        }
    }
    return out;
}

inline void correct(const lvio2d_scan_header& h, Points* pts) {
    const Vec3<double> lin(h.linear[0], h.linear[1], h.linear[2]), ang(h.angular[0], h.angular[1], h.angular[2]);
    for (size_t i = 0; i < pts->p.size(); ++i) {
        const double dt = pts->t[i] - h.stamp;
        const Iso3<double> T = lie::make_tf<double>(lin * dt, ang * dt);
        pts->p[i] = T.R * pts->p[i] + T.t;
    }
}

}  // namespace scanpts
}  // namespace oracle
