"""Hot-path parameter values = what `param::manager` hands to the solver (reference
src/utilies/params.cpp:92-175), with the shipped `config/corridor.yaml` numbers as defaults.

`normalize_tf` mirrors lie::normalize_tf (src/utilies/common.h:183-189), applied to every extrinsic at
load time by load_transform (src/utilies/params.cpp:42-53): R -> Eigen::Quaternion(R) -> toRotationMatrix.
"""
import math

import numpy as np

from . import abi

# config/corridor.yaml:13-16, :28-31 (row-major 4x4, last row dropped)
CORRIDOR_T_IMU_TO_WHEEL = [
    0.0040697, -0.9998940, -0.0139789, -0.061,
    0.0099712, 0.0140189, -0.9998520, 0.919,
    0.9999420, 0.0039297, 0.0100272, -0.224,
]
CORRIDOR_T_IMU_TO_LASER = [
    0.0019070, -0.9999900, 0.0040438, 0.024,
    0.0459794, -0.0039519, -0.9989346, -0.078,
    0.9989406, 0.0020909, 0.0459714, -0.071,
]


def _mat_to_quat(m):
    """Eigen::Quaternion(Matrix3): trace branch, else largest-diagonal branch. Returns (w, x, y, z)."""
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = math.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def _quat_to_mat(q):
    """Eigen::QuaternionBase::toRotationMatrix (no renormalisation)."""
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([
        [1 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1 - (txx + tyy)],
    ])


def normalize_tf(T34):
    T = np.asarray(T34, dtype=np.float64).reshape(3, 4).copy()
    T[:, :3] = _quat_to_mat(_mat_to_quat(T[:, :3]))
    return T


def corridor_params(fast_mode=False, max_iters=None, device=0):
    """lvio2d_params filled from config/corridor.yaml (:13-31 extrinsics, :39 g, :44-51 noise, :62-64
    manifold sigmas, :90 line_to_line_sigma, :123 fast_mode)."""
    p = abi.Params()
    p.abi_version = abi.ABI_VERSION
    p.device = device
    p.T_imu_to_laser[:] = normalize_tf(CORRIDOR_T_IMU_TO_LASER).ravel().tolist()
    p.T_imu_to_wheel[:] = normalize_tf(CORRIDOR_T_IMU_TO_WHEEL).ravel().tolist()
    p.g = 9.8
    p.line_to_line_sigma = 0.01
    p.manifold_p_sigma = 0.01
    p.manifold_q_sigma = 0.0001
    p.imu_noise_acc_sigma[:] = [0.0163] * 3
    p.imu_bias_acc_sigma[:] = [0.00499] * 3
    p.imu_noise_gyro_sigma[:] = [0.003208] * 3
    p.imu_bias_gyro_sigma[:] = [0.000499] * 3
    p.wheel_sigma[:] = [0.02, 99999.0, 999.99]
    # ceres default 50; solver.cpp:800-801 lowers it to 10 in fast_mode
    p.max_iters = max_iters if max_iters is not None else (10 if fast_mode else 50)
    p.assoc_mode = abi.ASSOC_FIXED
    p.huber_delta = 0.0
    p.assoc_gate = 0.0
    p.assoc_max_dist = 0.0
    p.function_tolerance = 0.0
    p.gradient_tolerance = 0.0
    p.parameter_tolerance = 0.0
    p.initial_trust_region_radius = 0.0
    return p


def params_T(p, name):
    return np.array(list(getattr(p, name)), dtype=np.float64).reshape(3, 4)


def corridor_line_params(**overrides):
    """lvio2d_line_params with the values of config/corridor.yaml:79-89."""
    lp = abi.LineParams(line_continuous_threshold=0.5, line_max_tolerance_angle_deg=175.0, line_max_dis=0.1,
                        line_min_len=0.05, laser_resolution=0.1, w_laser_each_scan=100.0, h_laser_each_scan=100.0)
    for k, v in overrides.items():
        setattr(lp, k, v)
    return lp
