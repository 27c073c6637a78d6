"""Tracking controllers and inner-loop PIDs -- TEST INFRASTRUCTURE ONLY.

numpy fp64 restatement of the reference's control/geometric.py,
control/lqr/*.py and control/low_level/*.py (App. B quirks reproduced on
purpose); pinned against the imported reference by
tests/test_oracle_vs_reference.py and tests/golden/controllers.npz.
``DslPid`` restates the upstream DSLPIDControl (PARITY UNPINNED).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

from oracle import conversions as cv
from oracle.constants import DroneModel
from oracle.models import linear_model_matrices

PWM2RPM_SCALE, PWM2RPM_CONST, MIN_PWM, MAX_PWM = 0.2685, 4070.3, 20000.0, 65535.0
MIXER = {"cf2x": np.array([[-.5, -.5, -1.], [-.5, .5, 1.], [.5, .5, -1.], [.5, -.5, 1.]]),
         "cf2p": np.array([[0., -1., -1.], [1., 0., 1.], [0., 1., -1.], [-1., 0., 1.]])}


def _hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def _vee(M):
    """control/geometric.py:36-44 (note the reference's sign/ordering)."""
    return np.array([-M[1, 2], M[0, 2], -M[0, 1]])


def _unit(v):
    return v / np.linalg.norm(v)


# --------------------------------------------------------------------------
# geometric SE(3) controller
# --------------------------------------------------------------------------
class Geometric:
    """control/geometric.py:7-115."""

    KP, KV, KR, KW = 2.25, 3.5, 125.0, 10.0   # :14-17
    G_CTRL = 9.81                             # :20  (quirk B2; the env uses 9.8)
    MAX_TILT = 40 * math.pi / 180             # :23

    def __init__(self, env):
        self.env = env
        self.ref = None

    def set_desired_trajectory(self, robot_idx, desired_pos, desired_vel, desired_acc, desired_yaw, desired_omega):
        self.ref = (np.asarray(desired_pos, float), np.asarray(desired_vel, float),
                    np.asarray(desired_acc, float), float(desired_yaw), float(desired_omega))

    def compute_input(self, obs):
        """u = [f, tau] before the mixer (control/geometric.py:59-114)."""
        m, J = self.env.M, self.env.J
        p_d, v_d, a_d, yaw, yaw_rate = self.ref
        s = cv.obs_to_geo_model(obs)
        p, Rm, v_w, w = s[0:3], s[3:12].reshape(3, 3), s[12:15], s[15:18]   # w: WORLD rates used as body (B3)
        RT = Rm.T
        ev = RT @ v_w - RT @ v_d
        e3 = np.array([0.0, 0.0, 1.0])
        f_b = (RT @ (m * self.G_CTRL * e3) - m * (RT @ (self.KP * (p - p_d))) - m * self.KV * ev
               + m * (RT @ a_d - _hat(w) @ (RT @ v_d)))
        f_w = Rm @ f_b
        tilt = math.acos(f_w[2] / np.linalg.norm(f_w))
        if tilt > self.MAX_TILT:                                      # :79-84 (B5)
            f_w[:2] = f_w[:2] * (f_w[2] * math.tan(self.MAX_TILT) / np.linalg.norm(f_w[:2]))
        f_b = RT @ f_w
        fn = np.linalg.norm(f_w)
        b1c = np.array([math.cos(yaw), math.sin(yaw), 0.0])
        b3 = f_w / fn
        b2 = _unit(np.cross(b3, b1c))
        b1 = _unit(np.cross(b2, b3))
        Rd = np.column_stack([b1, b2, b3])
        b1c_dot = np.array([-math.sin(yaw) * yaw_rate, math.cos(yaw) * yaw_rate, 0.0])
        f_dot = m * (Rm @ (self.KP * ev)) / fn                         # :96 (B4: Kp, not Kv)
        b3_dot = np.cross(np.cross(b3, f_dot), b3)
        b2_dot = np.cross(np.cross(b2, (np.cross(b1c_dot, b3) + np.cross(b1c, b3_dot))
                                   / np.linalg.norm(np.cross(b1c, b3))), b2)
        b1_dot = np.cross(b3_dot, b2) + np.cross(b3, b2_dot)
        Rd_dot = np.column_stack([b1_dot, b2_dot, b3_dot])
        W = Rd @ Rd_dot                                               # :102 (B1: no transpose)
        w_d = np.array([W[2, 1], W[0, 2], W[1, 0]])
        eR = 0.5 * self.KR * _vee(Rd.T @ Rm - RT @ Rd)
        tau = J @ (-eR - self.KW * (w - RT @ (Rd @ w_d))) - np.cross(w, J @ w)
        return np.array([max(0.0, f_b[2]), tau[0], tau[1], tau[2]])

    def compute(self, obs):
        """-> RPM (4,) only (quirk B7)."""
        return cv.input_to_action(self.env, self.compute_input(obs))


# --------------------------------------------------------------------------
# inner loops
# --------------------------------------------------------------------------
class ThrustOmegaPid:
    """control/low_level/thrust_omega_ctrl.py:9-132 (quirk B11)."""

    P, I, D = 17500.0, 10.0, 0.0

    def __init__(self, env):
        self.env = env
        self.mixer = MIXER[env.DRONE_MODEL.value]
        self.reset()

    def reset(self):
        self.control_counter = 0
        self.last_omega = np.zeros(3)
        self.integral = np.zeros(3)

    def compute(self, u, dt, omega_body):
        """``computeControlFromInput`` :81-98 + ``omega_PID`` :103-132; mutates u[0]."""
        self.control_counter += 1
        u[0] = max(u[0], 0.0)
        pwm_thrust = min(max((math.sqrt(u[0] / (self.env.KF * 4)) - PWM2RPM_CONST) / PWM2RPM_SCALE, MIN_PWM), MAX_PWM)
        omega_body = np.asarray(omega_body, float)
        rate_e = -(omega_body - self.last_omega) / dt
        e = np.asarray(u[1:4], float) - omega_body
        self.last_omega = omega_body
        self.integral = np.clip(self.integral - e * dt, -1500.0, 1500.0)
        self.integral[0:2] = np.clip(self.integral[0:2], -1.0, 1.0)
        tq = np.clip(self.P * e + self.I * self.integral + self.D * rate_e, -3200, 3200)
        pwm = np.clip(pwm_thrust + self.mixer @ tq, MIN_PWM, MAX_PWM)
        return PWM2RPM_SCALE * pwm + PWM2RPM_CONST


class YankOmegaPid:
    """control/low_level/yank_omega_ctrl.py:9-55: thrust = f_cur + yank * dt, then ThrustOmegaPid."""

    def __init__(self, env):
        self.env = env
        self.inner = ThrustOmegaPid(env)

    def reset(self):
        self.inner.reset()

    def compute(self, u, dt, omega_body, cur_thrust):
        ut = np.array(u, dtype=float)
        ut[0] = cur_thrust + u[0] * dt
        return self.inner.compute(ut, dt, omega_body)


# --------------------------------------------------------------------------
# LQR family
# --------------------------------------------------------------------------
def bryson_weights(env, kind):
    """(Q, R): control/lqr/lqr_controller.py:17-37, lqr_omega_controller.py:15-31,
    lqr_YO_controller.py:18-39."""
    mt = env.MAX_THRUST
    if kind == "torque12":
        r = [1 / mt ** 2, 1 / 0.001 ** 2, 1 / 0.001 ** 2, 1 / 0.001 ** 2]
        q = [1 / (math.pi / 40) ** 2] * 3 + [1 / .25 ** 2] * 3 + [1 / .15 ** 2] * 3 + [1 / .05 ** 2] * 3
    elif kind == "omega9":
        r = [1 / mt ** 2, 100.0, 100.0, 100.0]
        q = [1 / (math.pi / 20) ** 2] * 2 + [1 / (math.pi / 40) ** 2] + [1 / .15 ** 2] * 3 + [1 / .05 ** 2] * 3
    elif kind == "yank10":
        max_yank = (mt / env.CTRL_TIMESTEP) / 200
        r = [1 / max_yank ** 2, 100.0, 100.0, 100.0]
        q = ([1 / (math.pi / 20) ** 2] * 2 + [1 / (math.pi / 40) ** 2] + [1 / (mt - env.M * env.G) ** 2]
             + [1 / .15 ** 2] * 3 + [1 / .05 ** 2] * 3)
    else:
        raise ValueError(kind)
    return np.diag(q), np.diag(r)


def lqr_gain(env, kind, use_noisy_model=False):
    """K = R^-1 B^T P with P from the CARE (lqr_controller.py:54-59)."""
    A, B, Ah, Bh = linear_model_matrices(env, kind)
    if use_noisy_model:
        A, B = Ah, Bh
    Q, Rw = bryson_weights(env, kind)
    P = sla.solve_continuous_are(A, B, Q, Rw, e=None, s=None, balanced=True)
    return sla.solve(Rw, B.T @ P)


class Lqr:
    """LQRController / LQROmegaController / LQRYankOmegaController restated as one
    class with ``kind`` in {'torque12','omega9','yank10'} (quirks B8-B10)."""

    DIM = {"torque12": 12, "omega9": 9, "yank10": 10}

    def __init__(self, env, kind, low_level=None, use_noisy_model=False, K=None):
        self.env, self.kind = env, kind
        self.K = lqr_gain(env, kind, use_noisy_model) if K is None else np.asarray(K, float)
        self.low = low_level
        self.ref = None

    def set_desired_trajectory(self, robot_idx, desired_pos, desired_vel, desired_acc, desired_yaw, desired_omega):
        self.ref = (np.asarray(desired_pos, float), np.asarray(desired_vel, float), float(desired_yaw), float(desired_omega))

    def error_state(self, obs):
        """Error state of the LQR variants (control/lqr/lqr_controller.py:90-106, lqr_omega_controller.py:97-106, lqr_YO_controller.py:106-116): Euler error via scipy from_euler/as_euler("xyz") of R_eq^T R, position / velocity errors rotated by R_eq^T (yaw only); quirk B8."""
        p_d, v_d, yaw_d, om_d = self.ref
        dim = self.DIM[self.kind]
        x = cv.obs_to_lin_model(obs, dim, self.env)
        e = x.copy()
        Req = cv.euler_xyz_to_rot([0.0, 0.0, yaw_d])
        e[0:3] = cv.rot_to_euler_xyz(Req.T @ cv.euler_xyz_to_rot(x[0:3]))
        e[-3:] = Req.T @ (x[-3:] - p_d)
        e[-6:-3] = Req.T @ (x[-6:-3] - v_d)
        if self.kind == "torque12":
            e[3:6] = Req.T @ (x[3:6] - np.array([0.0, 0.0, om_d]))
        elif self.kind == "yank10":
            e[3] = x[3] - self.env.M * self.env.G
        return e

    def cap_u(self, u):
        """lqr_omega_controller.py:116-119 (omega9 only)."""
        u[0] = min(max(u[0], 4 * (cv.MIN_RPM ** 2 * self.env.KF)), self.env.MAX_THRUST)
        return u

    def body_rates(self, obs):
        """World -> body angular velocity with scipy's normalised R (lqr_omega_controller.py:80-84)."""
        return cv.quat_to_rot(obs[3:7]).T @ np.asarray(obs[13:16], float)

    def compute_low_level(self, u, obs, idx=0):
        """``compute_low_level`` (lqr_omega_controller.py:78-88, lqr_YO_controller.py:87-98): body rates + the inner PID -> RPM."""
        dt = self.env.CTRL_TIMESTEP
        if self.kind == "omega9":
            return self.low.compute(u, dt, self.body_rates(obs))
        if self.kind == "yank10":
            return self.low.compute(u, dt, self.body_rates(obs), cv.calc_z_thrust(self.env, obs))
        raise ValueError("torque12 has no inner loop")

    def compute(self, obs, skip_low_level=False):
        u = -self.K @ self.error_state(obs)
        if self.kind == "torque12":
            u[0] += self.env.M * self.env.G
            action = cv.input_to_action(self.env, u)        # clamps u[0] >= 0 in place
            return action, u
        if self.kind == "omega9":
            u[0] += self.env.M * self.env.G
            if skip_low_level:
                return None, self.cap_u(u)
            action = self.compute_low_level(u, obs)          # sees the un-capped u (B9)
            return action, self.cap_u(u)
        if skip_low_level:
            return None, u
        return self.compute_low_level(u, obs), u


# --------------------------------------------------------------------------
# upstream DSLPIDControl (config 1) -- PARITY UNPINNED
# --------------------------------------------------------------------------
class DslPid:
    """Upstream gym-pybullet-drones ``DSLPIDControl`` restated from SURVEY.md A.5;
    call site MultiDroneExample.py:111-114, halved gains :85-92."""

    def __init__(self, env, gain_scale=1.0):
        self.env = env
        self.GRAVITY = env.G * env.M
        self.KF = env.KF
        s = gain_scale
        self.P_FOR = s * np.array([.4, .4, 1.25])
        self.I_FOR = s * np.array([.05, .05, .05])
        self.D_FOR = s * np.array([.2, .2, .5])
        self.P_TOR = s * np.array([70000., 70000., 60000.])
        self.I_TOR = s * np.array([.0, .0, 500.])
        self.D_TOR = s * np.array([20000., 20000., 12000.])
        self.mixer = MIXER[env.DRONE_MODEL.value]
        self.reset()

    def reset(self):
        self.control_counter = 0
        self.last_rpy = np.zeros(3)
        self.integral_pos_e = np.zeros(3)
        self.integral_rpy_e = np.zeros(3)

    def compute_from_state(self, dt, state, target_pos, target_rpy=np.zeros(3),
                           target_vel=np.zeros(3), target_rpy_rates=np.zeros(3)):
        """Upstream ``DSLPIDControl.computeControlFromState`` -> ``_dslPIDPositionControl`` + ``_dslPIDAttitudeControl`` (call site MultiDroneExample.py:111-114).  Returns (rpm, pos_e)."""
        self.control_counter += 1
        pos, quat, vel = state[0:3], state[3:7], state[10:13]
        Rm = cv.quat_to_rot(quat)
        # position loop
        pe = np.asarray(target_pos, float) - pos
        ve = np.asarray(target_vel, float) - vel
        self.integral_pos_e = np.clip(self.integral_pos_e + pe * dt, -2., 2.)
        self.integral_pos_e[2] = np.clip(self.integral_pos_e[2], -0.15, .15)
        tt = self.P_FOR * pe + self.I_FOR * self.integral_pos_e + self.D_FOR * ve + np.array([0, 0, self.GRAVITY])
        scalar_thrust = max(0., float(tt @ Rm[:, 2]))
        thrust = (math.sqrt(scalar_thrust / (4 * self.KF)) - PWM2RPM_CONST) / PWM2RPM_SCALE
        zax = tt / np.linalg.norm(tt)
        xc = np.array([math.cos(target_rpy[2]), math.sin(target_rpy[2]), 0.0])
        yax = np.cross(zax, xc) / np.linalg.norm(np.cross(zax, xc))
        xax = np.cross(yax, zax)
        Rd = np.column_stack([xax, yax, zax])
        # attitude loop
        from oracle.aviary import quat_to_rpy
        rpy = quat_to_rpy(quat)
        M = Rd.T @ Rm - Rm.T @ Rd
        rot_e = np.array([M[2, 1], M[0, 2], M[1, 0]])
        rate_e = np.asarray(target_rpy_rates, float) - (rpy - self.last_rpy) / dt
        self.last_rpy = rpy
        self.integral_rpy_e = np.clip(self.integral_rpy_e - rot_e * dt, -1500., 1500.)
        self.integral_rpy_e[0:2] = np.clip(self.integral_rpy_e[0:2], -1., 1.)
        tq = np.clip(-self.P_TOR * rot_e + self.D_TOR * rate_e + self.I_TOR * self.integral_rpy_e, -3200, 3200)
        pwm = np.clip(thrust + self.mixer @ tq, MIN_PWM, MAX_PWM)
        return PWM2RPM_SCALE * pwm + PWM2RPM_CONST, pe
