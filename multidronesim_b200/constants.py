"""Drone constants (upstream URDF values) and the C parameter block.

Corroborated by the reference: G, M, cf2p inertia (utils/graph_fedce.py:9,44-49),
KF, KM (model/dynamics.py:38-39).  The remaining coefficients are recalled from the
upstream cf2x/cf2p URDF files (parity unpinned)."""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from .enums import DroneModel, Physics

G = 9.8

URDF = {
    DroneModel.CF2X: dict(m=0.027, arm=0.0397, thrust2weight=2.25, ixx=1.4e-5, iyy=1.4e-5, izz=2.17e-5,
                          kf=3.16e-10, km=7.94e-12, collision_h=0.025, collision_r=0.06, collision_z_offset=0.0,
                          gnd_eff_coeff=11.36859, prop_radius=2.31348e-2, drag_xy=9.1785e-7, drag_z=10.311e-7,
                          dw1=2267.18, dw2=0.16, dw3=-0.11,
                          prop_xy=((0.028, -0.028), (-0.028, -0.028), (-0.028, 0.028), (0.028, 0.028))),
    DroneModel.CF2P: dict(m=0.027, arm=0.0397, thrust2weight=2.25, ixx=2.3951e-5, iyy=2.3951e-5, izz=3.2347e-5,
                          kf=3.16e-10, km=7.94e-12, collision_h=0.025, collision_r=0.06, collision_z_offset=0.0,
                          gnd_eff_coeff=11.36859, prop_radius=2.31348e-2, drag_xy=9.1785e-7, drag_z=10.311e-7,
                          dw1=2267.18, dw2=0.16, dw3=-0.11,
                          prop_xy=((0.0397, 0.0), (0.0, 0.0397), (-0.0397, 0.0), (0.0, -0.0397))),
}


class DroneConstants:
    """Attributes named as the reference reads them from ``env``."""

    def __init__(self, drone_model=DroneModel.CF2P, physics=Physics.DYN, pyb_freq=240, ctrl_freq=240,
                 cf2x_torque_sign=-1, renormalize_quat=False, ground_clamp=None, dw_dz_clip=None, x_frame_mixer=False):
        self.DRONE_MODEL = DroneModel(drone_model)
        self.PHYSICS = Physics(physics)
        u = URDF[self.DRONE_MODEL]
        self.G = G
        self.M = u["m"]
        self.L = u["arm"]
        self.THRUST2WEIGHT_RATIO = u["thrust2weight"]
        self.J = np.diag([u["ixx"], u["iyy"], u["izz"]])
        self.J_INV = np.linalg.inv(self.J)
        self.KF, self.KM = u["kf"], u["km"]
        self.COLLISION_H, self.COLLISION_R, self.COLLISION_Z_OFFSET = u["collision_h"], u["collision_r"], u["collision_z_offset"]
        self.GND_EFF_COEFF, self.PROP_RADIUS = u["gnd_eff_coeff"], u["prop_radius"]
        self.DRAG_COEFF = np.array([u["drag_xy"], u["drag_xy"], u["drag_z"]])
        self.DW_COEFF_1, self.DW_COEFF_2, self.DW_COEFF_3 = u["dw1"], u["dw2"], u["dw3"]
        self.PROP_XY = np.array(u["prop_xy"])
        self.GRAVITY = self.G * self.M
        self.HOVER_RPM = math.sqrt(self.GRAVITY / (4 * self.KF))
        self.MAX_RPM = math.sqrt((self.THRUST2WEIGHT_RATIO * self.GRAVITY) / (4 * self.KF))
        self.MAX_THRUST = 4 * self.KF * self.MAX_RPM ** 2
        if self.DRONE_MODEL == DroneModel.CF2X:
            self.MAX_XY_TORQUE = (2 * self.L * self.KF * self.MAX_RPM ** 2) / math.sqrt(2)
        else:
            self.MAX_XY_TORQUE = self.L * self.KF * self.MAX_RPM ** 2
        self.MAX_Z_TORQUE = 2 * self.KM * self.MAX_RPM ** 2
        self.GND_EFF_H_CLIP = 0.25 * self.PROP_RADIUS * math.sqrt(
            (15 * self.MAX_RPM ** 2 * self.KF * self.GND_EFF_COEFF) / self.MAX_THRUST)
        self.PYB_FREQ, self.CTRL_FREQ = int(pyb_freq), int(ctrl_freq)
        if self.PYB_FREQ % self.CTRL_FREQ != 0:
            raise ValueError("[ERROR] in BaseAviary.__init__(), pyb_freq is not divisible by env_freq.")
        self.PYB_STEPS_PER_CTRL = self.PYB_FREQ // self.CTRL_FREQ
        self.CTRL_TIMESTEP = 1.0 / self.CTRL_FREQ
        self.PYB_TIMESTEP = 1.0 / self.PYB_FREQ
        self.Z_FLOOR = self.COLLISION_H / 2 - self.COLLISION_Z_OFFSET
        self.cf2x_torque_sign = int(cf2x_torque_sign)
        self.renormalize_quat = bool(renormalize_quat)
        # False: the reference's PLUS-frame mixer whatever the model (utils/model_conversions.py:74-77); True: a CF2X gets the
        # X-frame allocation its dynamics apply, so that torque-level controllers (geometric, 12-dim LQR) fly it
        self.x_frame_mixer = bool(x_frame_mixer)
        self.ground_clamp = (self.PHYSICS == Physics.DYN_GND_DRAG_DW) if ground_clamp is None else bool(ground_clamp)
        # Downwash magnitude alpha = DW1 (PROP_RADIUS / (4 dz))^2 is singular as dz -> 0+ (upstream only ever flies one
        # drone well above another).  The composite mode clips dz from below where alpha would exceed the drone's
        # weight -- the same device upstream uses for ground effect (GND_EFF_H_CLIP).  0 restores the upstream formula.
        self.DW_DZ_CLIP = 0.25 * self.PROP_RADIUS * math.sqrt(self.DW_COEFF_1 / self.GRAVITY) if dw_dz_clip is None else float(dw_dz_clip)

    def c_params(self) -> _lib.DroneParams:
        p = _lib.DroneParams()
        p.m, p.g, p.kf, p.km, p.arm_l = self.M, self.G, self.KF, self.KM, self.L
        p.ixx, p.iyy, p.izz = self.J[0, 0], self.J[1, 1], self.J[2, 2]
        p.max_rpm, p.max_thrust = self.MAX_RPM, self.MAX_THRUST
        p.gnd_eff_coeff, p.prop_radius, p.gnd_eff_h_clip = self.GND_EFF_COEFF, self.PROP_RADIUS, self.GND_EFF_H_CLIP
        p.drag_xy, p.drag_z = self.DRAG_COEFF[0], self.DRAG_COEFF[2]
        p.dw1, p.dw2, p.dw3 = self.DW_COEFF_1, self.DW_COEFF_2, self.DW_COEFF_3
        p.dw_dz_clip = self.DW_DZ_CLIP
        for i in range(4):
            p.prop_x[i], p.prop_y[i] = self.PROP_XY[i]
        p.z_floor, p.dt_phys, p.dt_ctrl = self.Z_FLOOR, self.PYB_TIMESTEP, self.CTRL_TIMESTEP
        p.substeps = self.PYB_STEPS_PER_CTRL
        p.drone_model = _lib.DRONE_CF2X if self.DRONE_MODEL == DroneModel.CF2X else _lib.DRONE_CF2P
        p.physics = _lib.PHYSICS_DYN if self.PHYSICS == Physics.DYN else _lib.PHYSICS_DYN_GND_DRAG_DW
        p.cf2x_torque_sign = self.cf2x_torque_sign
        p.renormalize_quat = int(self.renormalize_quat)
        p.ground_clamp = int(self.ground_clamp)
        p.x_frame_mixer = int(self.x_frame_mixer)
        return p
