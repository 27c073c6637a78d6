"""obs <-> model-state conversions and the PLUS-frame mixer -- TEST INFRASTRUCTURE ONLY.

numpy fp64 restatement of the reference's utils/model_conversions.py; pinned
against the imported reference by tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import math

import numpy as np

MIN_RPM = 9440.3  # utils/model_conversions.py:98-99


def quat_to_rot(q):
    """scipy ``Rotation.from_quat(q).as_matrix()`` (xyzw; normalises first).
    Used at utils/model_conversions.py:110 and control/lqr/lqr_omega_controller.py:81."""
    x, y, z, w = np.asarray(q, dtype=float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def euler_xyz_to_rot(rpy):
    """scipy ``Rotation.from_euler('xyz', rpy).as_matrix()`` -- extrinsic xyz,
    i.e. Rz(yaw) @ Ry(pitch) @ Rx(roll) (control/lqr/lqr_omega_controller.py:97-98)."""
    a, b, c = rpy
    ca, sa, cb, sb, cc, sc = math.cos(a), math.sin(a), math.cos(b), math.sin(b), math.cos(c), math.sin(c)
    return np.array([[cb * cc, sa * sb * cc - ca * sc, ca * sb * cc + sa * sc],
                     [cb * sc, sa * sb * sc + ca * cc, ca * sb * sc - sa * cc],
                     [-sb, sa * cb, ca * cb]])


def rot_to_euler_xyz(Rm):
    """scipy ``Rotation.from_matrix(Rm).as_euler('xyz')`` away from gimbal lock
    (control/lqr/lqr_omega_controller.py:101)."""
    sb = min(1.0, max(-1.0, -Rm[2, 0]))
    return np.array([math.atan2(Rm[2, 1], Rm[2, 2]), math.asin(sb), math.atan2(Rm[1, 0], Rm[0, 0])])


def mixer_matrix(env):
    """PLUS-frame [f, tx, ty, tz] = C @ motor_thrusts (utils/model_conversions.py:74-77,90-93)."""
    r = env.KM / env.KF
    L = env.L
    if getattr(env, "x_frame_mixer", False) and getattr(getattr(env, "DRONE_MODEL", None), "value", None) == "cf2x":
        # builder extension: the X-frame allocation of the CF2X dynamics (oracle/aviary.py _substep_drone)
        l2, s = L / np.sqrt(2.0), float(getattr(env, "cf2x_torque_sign", -1.0))
        return np.array([[1.0, 1.0, 1.0, 1.0], [s * l2, s * l2, -s * l2, -s * l2], [-l2, l2, l2, -l2], [-r, r, -r, r]])
    return np.array([[1.0, 1.0, 1.0, 1.0],
                     [0.0, L, 0.0, -L],
                     [-L, 0.0, L, 0.0],
                     [-r, r, -r, r]])


def calc_z_thrust(env, obs):
    """utils/model_conversions.py:137-143."""
    return float(np.sum(env.KF * np.asarray(obs[-4:], dtype=float) ** 2))


def obs_to_lin_model(obs, dim=12, env=None):
    """utils/model_conversions.py:20-58 (state layouts of SURVEY.md App. D)."""
    obs = np.asarray(obs, dtype=float)
    rpy, vel, pos = obs[7:10], obs[10:13], obs[0:3]
    if dim == 12:
        return np.concatenate([rpy, obs[13:16], vel, pos])
    if dim == 9:
        return np.concatenate([rpy, vel, pos])
    if dim == 10:
        if env is None:
            raise AssertionError("env must be provided for 10 dim model to calculate the thrust")
        return np.concatenate([rpy, [calc_z_thrust(env, obs)], vel, pos])
    raise ValueError("Invalid dim for linear model")


def obs_to_geo_model(obs):
    """utils/model_conversions.py:105-114: [p3, R9 row-major, v3 world, w3 (obs[13:16] verbatim)]."""
    obs = np.asarray(obs, dtype=float)
    return np.concatenate([obs[0:3], quat_to_rot(obs[3:7]).reshape(-1), obs[10:13], obs[13:16]])


def action_to_input(env, action, cap_rpm=True):
    """utils/model_conversions.py:69-83: RPM -> [f, tx, ty, tz]."""
    rpm = np.asarray(action, dtype=float)
    if cap_rpm:
        rpm = np.clip(rpm, 0, env.MAX_RPM)
    return mixer_matrix(env) @ (env.KF * rpm ** 2)


def input_to_action(env, u):
    """utils/model_conversions.py:85-103: [f, tx, ty, tz] -> RPM.  Mutates ``u[0]``
    (clamped at 0) like the reference; per-motor thrust clipped to
    [MIN_RPM^2 KF, MAX_THRUST] (quirk B6: MAX_THRUST is the 4-motor total)."""
    u[0] = max(u[0], 0.0)
    thrusts = np.linalg.inv(mixer_matrix(env)) @ u
    thrusts = np.clip(thrusts, MIN_RPM ** 2 * env.KF, env.MAX_THRUST)
    return np.sqrt(thrusts / env.KF)


def geo_x_dot_to_linear(g):
    """utils/model_conversions.py:124-135: reorder (v, w, vdot, wdot) -> (w, wdot, vdot, v)."""
    g = np.asarray(g, dtype=float)
    return np.concatenate([g[3:6], g[9:12], g[6:9], g[0:3]])
