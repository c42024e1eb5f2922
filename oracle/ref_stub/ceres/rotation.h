// ceres/rotation.h — STUB (test infrastructure, oracle/_ref): the two conversions the reference calls
// (src/utilies/common.h:142, :155), restated from Ceres 1.14's documented behaviour, including the first-order
// Taylor branches at the origin that keep Jet derivatives finite.
#pragma once
#include "ceres/jet.h"

namespace ceres {

template <typename T>
inline void AngleAxisToQuaternion(const T* angle_axis, T* quaternion) {
    const T& a0 = angle_axis[0];
    const T& a1 = angle_axis[1];
    const T& a2 = angle_axis[2];
    const T theta_squared = a0 * a0 + a1 * a1 + a2 * a2;
    if (theta_squared > T(0.0)) {
        const T theta = sqrt(theta_squared);
        const T half_theta = theta * T(0.5);
        const T k = sin(half_theta) / theta;
        quaternion[0] = cos(half_theta);
        quaternion[1] = a0 * k;
        quaternion[2] = a1 * k;
        quaternion[3] = a2 * k;
    } else {
        const T k(0.5);
        quaternion[0] = T(1.0);
        quaternion[1] = a0 * k;
        quaternion[2] = a1 * k;
        quaternion[3] = a2 * k;
    }
}

template <typename T>
inline void QuaternionToAngleAxis(const T* quaternion, T* angle_axis) {
    const T& q1 = quaternion[1];
    const T& q2 = quaternion[2];
    const T& q3 = quaternion[3];
    const T sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3;
    if (sin_squared_theta > T(0.0)) {
        const T sin_theta = sqrt(sin_squared_theta);
        const T& cos_theta = quaternion[0];
        // cos_theta < 0: theta > pi/2, so 2 theta > pi; use theta - pi = atan2(-sin, -cos) to stay in (-pi, pi]
        const T two_theta = T(2.0) * ((cos_theta < T(0.0)) ? atan2(-sin_theta, -cos_theta) : atan2(sin_theta, cos_theta));
        const T k = two_theta / sin_theta;
        angle_axis[0] = q1 * k;
        angle_axis[1] = q2 * k;
        angle_axis[2] = q3 * k;
    } else {
        const T k(2.0);
        angle_axis[0] = q1 * k;
        angle_axis[1] = q2 * k;
        angle_axis[2] = q3 * k;
    }
}

}  // namespace ceres
