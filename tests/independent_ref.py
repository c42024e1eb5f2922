"""Independent float64 re-derivation of the reference residuals in torch (autograd Jacobians).

Written from the reference's MATH, deliberately through a different route than the
oracle: Rodrigues' formula instead of angle-axis -> quaternion -> matrix, the trace/vee log map instead of
matrix -> quaternion -> angle-axis, signed point-line distances instead of vector rejections.  It pins the
oracle's values and Jacobians (tests/test_oracle_factors.py); it is test infrastructure only.
"""
import math

import torch

torch.set_default_dtype(torch.float64)


def hat(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def exp_so3(v):
    th2 = (v * v).sum()
    K = hat(v)
    eye = torch.eye(3, dtype=v.dtype)
    if float(th2.detach()) < 1e-16:
        return eye + K + 0.5 * K @ K
    th = torch.sqrt(th2)
    return eye + (torch.sin(th) / th) * K + ((1.0 - torch.cos(th)) / th2) * (K @ K)


def log_so3(R):
    w = 0.5 * torch.stack([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    c = 0.5 * (R[0, 0] + R[1, 1] + R[2, 2] - 1.0)
    s = torch.sqrt((w * w).sum())
    if float(s.detach()) < 1e-12:
        return w
    th = torch.atan2(s, c)
    return w * (th / s)


def T34(a):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(3, 4)
    return a[:, :3], a[:, 3]


def laser_point_residual(pose_i, pose_j, a1, a2, c, weight, T_il, sigma):
    """w/sigma * distance of the flattened point to the flattened line (laser_factor.h:67-86)."""
    R_il, t_il = T34(T_il)
    Ri, Rj = exp_so3(pose_i[3:6]), exp_so3(pose_j[3:6])

    def world_xy(R, p, pt):
        x = R_il @ torch.cat([pt, torch.zeros(1)]) + t_il
        return (R @ x + p)[:2]

    A1, A2 = world_xy(Ri, pose_i[0:3], a1), world_xy(Ri, pose_i[0:3], a2)
    Cw = world_xy(Rj, pose_j[0:3], c)
    u = (A2 - A1) / torch.sqrt(((A2 - A1) ** 2).sum())
    nrm = torch.stack([-u[1], u[0]])
    d = (nrm * (Cw - A2)).sum()
    return (weight / sigma) * torch.abs(d)


def laser_pair_residual(pose_i, pose_j, l1_p1, l1_p2, l2_p1, l2_p2, T_il, sigma):
    len1 = float(torch.linalg.norm(l1_p1 - l1_p2))
    len2 = float(torch.linalg.norm(l2_p1 - l2_p2))
    w = math.sqrt(min(len1, len2) / 2.0 / 0.02)
    return torch.stack([laser_point_residual(pose_i, pose_j, l1_p1[:2], l1_p2[:2], l2_p1[:2], w, T_il, sigma),
                        laser_point_residual(pose_i, pose_j, l1_p1[:2], l1_p2[:2], l2_p2[:2], w, T_il, sigma)])


def imu_residual(si, sj, blob, g):
    """imu_factor.h:52-86 with state = [p q v ba bw] and blob = X15 | J225 | sqrtP225 | Dt."""
    blob = torch.as_tensor(blob, dtype=torch.float64)
    X, J, S, Dt = blob[:15], blob[15:240].reshape(15, 15), blob[240:465].reshape(15, 15), blob[465]
    pi, qi, vi, bai, bwi = si[0:3], si[3:6], si[6:9], si[9:12], si[12:15]
    pj, qj, vj, baj, bwj = sj[0:3], sj[3:6], sj[6:9], sj[9:12], sj[12:15]
    dba, dbw = bai - X[9:12], bwi - X[12:15]
    alpha = X[0:3] + J[0:3, 9:12] @ dba + J[0:3, 12:15] @ dbw
    beta = X[3:6] + J[3:6, 9:12] @ dba + J[3:6, 12:15] @ dbw
    gamma = X[6:9] + J[6:9, 12:15] @ dbw
    Ri = exp_so3(qi)
    gz = torch.tensor([0.0, 0.0, 1.0]) * g
    r_a = alpha - Ri.T @ (pj - pi + 0.5 * gz * Dt * Dt - vi * Dt)
    r_b = beta - Ri.T @ (vj + gz * Dt - vi)
    r_g = log_so3(exp_so3(gamma).T @ Ri.T @ exp_so3(qj))
    r = torch.cat([r_a, r_b, r_g, baj - bai, bwj - bwi])
    return S @ r


def wheel_residual(pose_i, pose_j, blob, T_io):
    """wheel_factor.h:20-71 (branches chosen on the values, like the Jet comparisons)."""
    blob = torch.as_tensor(blob, dtype=torch.float64)
    R_io, t_io = T34(T_io)
    dR, dt = T34(blob[:12])
    s = blob[12:15]
    Ri, Rj = exp_so3(pose_i[3:6]), exp_so3(pose_j[3:6])
    Roi, toi = Ri @ R_io, Ri @ t_io + pose_i[0:3]
    Roj, toj = Rj @ R_io, Rj @ t_io + pose_j[0:3]
    p = Roi.T @ (toj - toi)
    q = log_so3(Roi.T @ Roj)
    op, oq = dt, log_so3(dR)
    o_len = torch.sqrt(op[0] ** 2 + op[1] ** 2)
    ln = torch.sqrt(p[0] ** 2 + p[1] ** 2)
    if float(o_len.detach()) > 1e-4 and float(ln.detach()) > 1e-4:
        cr = (op[0] * p[1] - op[1] * p[0]) / (o_len * ln)
        angle = torch.asin(torch.abs(cr))
    else:
        angle = ln
    r0 = s[0] * ln if (float(ln.detach()) < 1e-4 or float(o_len.detach()) < 1e-4) else s[0] * (o_len - ln)
    qn, oqn = torch.sqrt((q * q).sum()), torch.sqrt((oq * oq).sum())
    r2 = s[2] * qn if (float(qn.detach()) < 1e-3 or float(oqn.detach()) < 1e-3) else s[2] * (oqn - qn)
    return torch.stack([r0, s[1] * angle, r2])


def ground_residuals(pose, T_io, sigma_p, sigma_q):
    """ground_factor.h:27-48, :59-82."""
    R_io, t_io = T34(T_io)
    R = exp_so3(pose[3:6])
    z = (R @ t_io + pose[0:3])[2]
    zax = R @ R_io[:, 2]
    sinn = torch.sqrt(zax[0] ** 2 + zax[1] ** 2)  # |zax x e3|
    return torch.stack([z / sigma_p, torch.asin(sinn) / sigma_q])


def jac(fn, *xs):
    """Jacobian of fn w.r.t. the concatenation of xs (row-major [nr][sum len])."""
    xs = [torch.as_tensor(x, dtype=torch.float64) for x in xs]
    J = torch.autograd.functional.jacobian(fn, tuple(xs))
    return torch.cat([j.reshape(j.shape[0], -1) for j in J], dim=1).numpy()
