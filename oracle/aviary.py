"""CPU restatement of the upstream env step -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the upstream simulator (gym-pybullet-drones ``BaseAviary`` /
``CtrlAviary``; un-vendored, un-pinned, see reference README.md:11-18) is not in
/root/reference and not installed; this file restates its published explicit
``Physics.DYN`` update from SURVEY.md App. A.2/A.3/A.6 and defines the new
composite ``DYN_GND_DRAG_DW`` of App. A.4.  The reference constrains it only via

* the obs layout it consumes: utils/model_conversions.py:34-45,109-113
  (pos3, quat xyzw 4, rpy3, vel3 world, ang_vel3 WORLD, last clipped rpm 4),
* the call sites simulations/EnvGeometric.py:431,469, simulations/CBFTest.py:299,350,
  MultiDroneExample.py:106,121 (``step(action) -> obs, reward, terminated,
  truncated, info``),
* the env attributes it reads (M, J, G, KF, KM, L, MAX_RPM, MAX_THRUST, ...).

One env, N drones, plain per-drone Python loop in numpy fp64 exactly like the
upstream code structure (this is also the timed CPU baseline).
Explicit switches for the recalled-only choices: ``cf2x_torque_sign`` and
``renormalize_quat``.
"""
from __future__ import annotations

import math

import numpy as np

from oracle.constants import DroneModel, Physics, drone_params


def quat_to_matrix(q):
    """Bullet ``getMatrixFromQuaternion`` (xyzw, self-normalising: s = 2/|q|^2)."""
    x, y, z, w = q
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return np.array([[1.0 - (yy + zz), xy - wz, xz + wy],
                     [xy + wz, 1.0 - (xx + zz), yz - wx],
                     [xz - wy, yz + wx, 1.0 - (xx + yy)]])


def quat_to_rpy(q):
    """Bullet ``getEulerFromQuaternion`` (xyzw -> roll, pitch, yaw; no normalisation)."""
    x, y, z, w = q
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        return np.array([0.0, -0.5 * math.pi, 2.0 * math.atan2(x, -y)])
    if sarg >= 0.99999:
        return np.array([0.0, 0.5 * math.pi, 2.0 * math.atan2(-x, y)])
    return np.array([math.atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz),
                     math.asin(sarg),
                     math.atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz)])


def rpy_to_quat(rpy):
    """Bullet ``getQuaternionFromEuler`` (roll, pitch, yaw -> xyzw)."""
    r, p, y = (0.5 * a for a in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


def integrate_q(quat, omega, dt):
    """Upstream ``_integrateQ``: exponential-map update, body rates."""
    n = float(np.linalg.norm(omega))
    if abs(n) <= 1e-8:  # np.isclose(n, 0) with default atol
        return quat.copy()
    p, q, r = omega
    lam = 0.5 * np.array([[0.0, r, -q, p],
                          [-r, 0.0, p, q],
                          [q, -p, 0.0, r],
                          [-p, -q, -r, 0.0]])
    th = n * dt / 2.0
    return (np.eye(4) * math.cos(th) + (2.0 / n) * lam * math.sin(th)) @ quat


class OracleCtrlAviary:
    """One environment of N drones, upstream ``CtrlAviary`` semantics under
    ``Physics.DYN`` and the composite ``Physics.DYN_GND_DRAG_DW``."""

    def __init__(self, drone_model=DroneModel.CF2P, num_drones=1, initial_xyzs=None,
                 initial_rpys=None, physics=Physics.DYN, pyb_freq=240, ctrl_freq=240,
                 gui=False, record=False, user_debug_gui=False, output_folder="results",
                 cf2x_torque_sign=-1.0, renormalize_quat=False, ground_clamp=None, dw_dz_clip=None, x_frame_mixer=False):
        prm = drone_params(drone_model, pyb_freq, ctrl_freq)
        self.__dict__.update(prm.__dict__)
        self.NUM_DRONES = int(num_drones)
        self.PHYSICS = Physics(physics.value if hasattr(physics, "value") else physics)
        if self.PHYSICS not in (Physics.DYN, Physics.DYN_GND_DRAG_DW):
            raise ValueError("oracle supports Physics.DYN and Physics.DYN_GND_DRAG_DW only")
        self.cf2x_torque_sign = float(cf2x_torque_sign)
        self.renormalize_quat = bool(renormalize_quat)
        self.x_frame_mixer = bool(x_frame_mixer)
        if dw_dz_clip is not None:
            self.DW_DZ_CLIP = float(dw_dz_clip)
        # ground-plane clamp belongs to the composite mode only (DYN stays upstream-exact)
        self.ground_clamp = (self.PHYSICS == Physics.DYN_GND_DRAG_DW) if ground_clamp is None else bool(ground_clamp)
        N = self.NUM_DRONES
        if initial_xyzs is None:
            initial_xyzs = np.vstack([np.array([x * 4 * self.L for x in range(N)]),
                                      np.array([y * 4 * self.L for y in range(N)]),
                                      np.ones(N) * (self.COLLISION_H / 2 - self.COLLISION_Z_OFFSET + .1)]).T
        if initial_rpys is None:
            initial_rpys = np.zeros((N, 3))
        self.INIT_XYZS = np.array(initial_xyzs, dtype=float).reshape(N, 3)
        self.INIT_RPYS = np.array(initial_rpys, dtype=float).reshape(N, 3)
        self.ext_force = np.zeros((N, 3))  # SURVEY 8(f).1 wind stand-in (world frame, N)
        self.reset()

    # ---- housekeeping -------------------------------------------------
    def reset(self):
        """Upstream ``BaseAviary.reset`` / ``_housekeeping``: state back to INIT_XYZS / INIT_RPYS, zero velocities and last RPM; returns (obs, info)."""
        N = self.NUM_DRONES
        self.pos = self.INIT_XYZS.copy()
        self.quat = np.array([rpy_to_quat(self.INIT_RPYS[i]) for i in range(N)])
        self.rpy = np.array([quat_to_rpy(self.quat[i]) for i in range(N)])
        self.vel = np.zeros((N, 3))
        self.ang_v = np.zeros((N, 3))       # WORLD frame (what PyBullet stores)
        self.rpy_rates = np.zeros((N, 3))   # BODY frame (DYN's private state)
        self.last_clipped_action = np.zeros((N, 4))
        self.step_counter = 0
        return self._compute_obs(), {"answer": 42}

    def set_state(self, pos, quat, vel, rpy_rates, last_rpm=None):
        """Test hook (no upstream counterpart): overwrite pos / quat / vel / body rates / last clipped RPM of every drone."""
        N = self.NUM_DRONES
        self.pos = np.array(pos, dtype=float).reshape(N, 3)
        self.quat = np.array(quat, dtype=float).reshape(N, 4)
        self.vel = np.array(vel, dtype=float).reshape(N, 3)
        self.rpy_rates = np.array(rpy_rates, dtype=float).reshape(N, 3)
        self.rpy = np.array([quat_to_rpy(self.quat[i]) for i in range(N)])
        self.ang_v = np.array([quat_to_matrix(self.quat[i]) @ self.rpy_rates[i] for i in range(N)])
        if last_rpm is not None:
            self.last_clipped_action = np.array(last_rpm, dtype=float).reshape(N, 4)

    def _compute_obs(self):
        return np.hstack([self.pos, self.quat, self.rpy, self.vel, self.ang_v,
                          self.last_clipped_action]).reshape(self.NUM_DRONES, 20)

    # ---- one physics sub-step for one drone ------------------------------
    def _substep_drone(self, rpm, i, snap_pos, dt):
        pos, quat, vel, w = self.pos[i], self.quat[i], self.vel[i], self.rpy_rates[i]
        R = quat_to_matrix(quat)
        forces = self.KF * rpm ** 2
        extra_world = np.zeros(3)
        dw = 0.0
        if self.PHYSICS == Physics.DYN_GND_DRAG_DW:
            # ground effect (A.3): per-prop height, clipped, gated on |roll|,|pitch| < pi/2
            rpy = quat_to_rpy(quat)
            if abs(rpy[0]) < math.pi / 2 and abs(rpy[1]) < math.pi / 2:
                h = pos[2] + R[2, 0] * self.PROP_XY[:, 0] + R[2, 1] * self.PROP_XY[:, 1]
                h = np.maximum(h, self.GND_EFF_H_CLIP)
                forces = forces + forces * self.GND_EFF_COEFF * (self.PROP_RADIUS / (4 * h)) ** 2
            # drag (A.3): previous clipped RPM, net world-frame vector
            extra_world = -self.DRAG_COEFF * np.sum(2 * math.pi * self.last_clipped_action[i] / 60) * vel
            # downwash (A.3): every drone above, snapshot positions
            for j in range(self.NUM_DRONES):
                if j == i:
                    continue
                dz = snap_pos[j, 2] - pos[2]
                dxy = math.hypot(snap_pos[j, 0] - pos[0], snap_pos[j, 1] - pos[1])
                if dz > 0 and dxy < 10:
                    alpha = self.DW_COEFF_1 * (self.PROP_RADIUS / (4 * max(dz, self.DW_DZ_CLIP))) ** 2
                    beta = self.DW_COEFF_2 * dz + self.DW_COEFF_3
                    dw += alpha * math.exp(-0.5 * (dxy / beta) ** 2)
        thrust_body = np.array([0.0, 0.0, np.sum(forces) - dw])
        force_world = R @ thrust_body + extra_world + self.ext_force[i] - np.array([0.0, 0.0, self.GRAVITY])
        zt = self.KM * rpm ** 2
        z_torque = -zt[0] + zt[1] - zt[2] + zt[3]
        if self.DRONE_MODEL == DroneModel.CF2X:
            x_torque = self.cf2x_torque_sign * (forces[0] + forces[1] - forces[2] - forces[3]) * (self.L / math.sqrt(2))
            y_torque = (-forces[0] + forces[1] + forces[2] - forces[3]) * (self.L / math.sqrt(2))
        else:
            x_torque = (forces[1] - forces[3]) * self.L
            y_torque = (-forces[0] + forces[2]) * self.L
        torques = np.array([x_torque, y_torque, z_torque]) - np.cross(w, self.J @ w)
        w_dot = self.J_INV @ torques
        acc = force_world / self.M
        vel_n = vel + dt * acc
        w_n = w + dt * w_dot
        pos_n = pos + dt * vel_n
        quat_n = integrate_q(quat, w_n, dt)
        if self.renormalize_quat:
            quat_n = quat_n / np.linalg.norm(quat_n)
        if self.ground_clamp and pos_n[2] < self.Z_FLOOR:
            pos_n[2] = self.Z_FLOOR
            vel_n[2] = max(vel_n[2], 0.0)
        # PyBullet is handed R(old quat) @ w_new as the world angular velocity
        return pos_n, quat_n, vel_n, w_n, R @ w_n

    # ---- step ----------------------------------------------------------
    def step(self, action):
        """Upstream ``BaseAviary.step`` under Physics.DYN as the reference calls it (simulations/EnvGeometric.py:431,469, CBFTest.py:299,350, MultiDroneExample.py:106,121): clip RPM to [0, MAX_RPM], PYB_STEPS_PER_CTRL explicit updates from a Jacobi snapshot of the positions, then obs (N,20), reward -1, False, False, {"answer": 42}."""
        N = self.NUM_DRONES
        clipped = np.clip(np.array(action, dtype=float).reshape(N, 4), 0, self.MAX_RPM)
        dt = self.PYB_TIMESTEP
        for _ in range(self.PYB_STEPS_PER_CTRL):
            snap = self.pos.copy()
            new = [self._substep_drone(clipped[i], i, snap, dt) for i in range(N)]
            for i, (p_, q_, v_, w_, av_) in enumerate(new):
                self.pos[i], self.quat[i], self.vel[i], self.rpy_rates[i], self.ang_v[i] = p_, q_, v_, w_, av_
            self.last_clipped_action = clipped.copy()
        self.rpy = np.array([quat_to_rpy(self.quat[i]) for i in range(N)])
        self.step_counter += self.PYB_STEPS_PER_CTRL
        return self._compute_obs(), -1, False, False, {"answer": 42}

    # no-ops kept so reference-style loops run unchanged
    def close(self):
        """Upstream ``BaseAviary.close`` (PyBullet disconnect): nothing to release here."""
        pass

    def render(self):
        """Upstream ``BaseAviary.render`` (MultiDroneExample.py:123): no-op."""
        pass
