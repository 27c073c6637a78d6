"""Drone constants for the oracle -- TEST INFRASTRUCTURE ONLY.

Restates the URDF-derived constants of upstream gym-pybullet-drones'
``BaseAviary.__init__`` / ``_parseURDFParameters`` (NOT in /root/reference;
SURVEY.md App. A.1).  Items corroborated by the reference itself:

* G = 9.8, M = 0.027, cf2p J = diag(2.3951e-5, 2.3951e-5, 3.2347e-5)
  -- utils/graph_fedce.py:9,44-49
* KF = 3.16e-10, KM = 7.94e-12 -- model/dynamics.py:38-39
* min RPM 9440.3 -- utils/model_conversions.py:98-99
* PWM2RPM 0.2685 / 4070.3, PWM range 20000..65535, mixers
  -- control/low_level/thrust_omega_ctrl.py:43-60

Everything else (cf2x inertia, arm length, ground-effect / drag / downwash
coefficients, collision cylinder) is recalled from the upstream URDFs and is
"parity unpinned".
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

from oracle.ref_import import DroneModel, Physics  # enum stand-ins (string-valued)

G = 9.8

_URDF = {
    "cf2x": dict(M=0.027, L=0.0397, T2W=2.25,
                 IXX=1.4e-5, IYY=1.4e-5, IZZ=2.17e-5,
                 KF=3.16e-10, KM=7.94e-12,
                 COLLISION_H=0.025, COLLISION_R=0.06, COLLISION_Z_OFFSET=0.0,
                 MAX_SPEED_KMH=30.0, GND_EFF_COEFF=11.36859, PROP_RADIUS=2.31348e-2,
                 DRAG_XY=9.1785e-7, DRAG_Z=10.311e-7,
                 DW1=2267.18, DW2=0.16, DW3=-0.11,
                 # prop link offsets in the body frame (x, y) -- cf2x.urdf
                 PROP_XY=((0.028, -0.028), (-0.028, -0.028), (-0.028, 0.028), (0.028, 0.028))),
    "cf2p": dict(M=0.027, L=0.0397, T2W=2.25,
                 IXX=2.3951e-5, IYY=2.3951e-5, IZZ=3.2347e-5,
                 KF=3.16e-10, KM=7.94e-12,
                 COLLISION_H=0.025, COLLISION_R=0.06, COLLISION_Z_OFFSET=0.0,
                 MAX_SPEED_KMH=30.0, GND_EFF_COEFF=11.36859, PROP_RADIUS=2.31348e-2,
                 DRAG_XY=9.1785e-7, DRAG_Z=10.311e-7,
                 DW1=2267.18, DW2=0.16, DW3=-0.11,
                 PROP_XY=((0.0397, 0.0), (0.0, 0.0397), (-0.0397, 0.0), (0.0, -0.0397))),
}


def drone_params(drone_model="cf2p", pyb_freq=240, ctrl_freq=240):
    """Namespace with the attribute names the reference reads from ``env``
     plus the force-model coefficients."""
    key = drone_model.value if hasattr(drone_model, "value") else str(drone_model)
    u = _URDF[key]
    p = SimpleNamespace()
    p.DRONE_MODEL = DroneModel(key)
    p.G = G
    p.M = u["M"]
    p.L = u["L"]
    p.THRUST2WEIGHT_RATIO = u["T2W"]
    p.J = np.diag([u["IXX"], u["IYY"], u["IZZ"]])
    p.J_INV = np.linalg.inv(p.J)
    p.KF = u["KF"]
    p.KM = u["KM"]
    p.COLLISION_H = u["COLLISION_H"]
    p.COLLISION_R = u["COLLISION_R"]
    p.COLLISION_Z_OFFSET = u["COLLISION_Z_OFFSET"]
    p.GND_EFF_COEFF = u["GND_EFF_COEFF"]
    p.PROP_RADIUS = u["PROP_RADIUS"]
    p.DRAG_COEFF = np.array([u["DRAG_XY"], u["DRAG_XY"], u["DRAG_Z"]])
    p.DW_COEFF_1, p.DW_COEFF_2, p.DW_COEFF_3 = u["DW1"], u["DW2"], u["DW3"]
    p.PROP_XY = np.array(u["PROP_XY"])
    p.GRAVITY = p.G * p.M
    p.HOVER_RPM = math.sqrt(p.GRAVITY / (4 * p.KF))
    p.MAX_RPM = math.sqrt((p.THRUST2WEIGHT_RATIO * p.GRAVITY) / (4 * p.KF))
    p.MAX_THRUST = 4 * p.KF * p.MAX_RPM ** 2
    if key == "cf2x":
        p.MAX_XY_TORQUE = (2 * p.L * p.KF * p.MAX_RPM ** 2) / math.sqrt(2)
    else:
        p.MAX_XY_TORQUE = p.L * p.KF * p.MAX_RPM ** 2
    p.MAX_Z_TORQUE = 2 * p.KM * p.MAX_RPM ** 2
    p.GND_EFF_H_CLIP = 0.25 * p.PROP_RADIUS * math.sqrt(
        (15 * p.MAX_RPM ** 2 * p.KF * p.GND_EFF_COEFF) / p.MAX_THRUST)
    p.PYB_FREQ = int(pyb_freq)
    p.CTRL_FREQ = int(ctrl_freq)
    if p.PYB_FREQ % p.CTRL_FREQ != 0:
        raise ValueError("pyb_freq must be a multiple of ctrl_freq")
    p.PYB_STEPS_PER_CTRL = p.PYB_FREQ // p.CTRL_FREQ
    p.CTRL_TIMESTEP = 1.0 / p.CTRL_FREQ
    p.PYB_TIMESTEP = 1.0 / p.PYB_FREQ
    p.Z_FLOOR = p.COLLISION_H / 2 - p.COLLISION_Z_OFFSET
    # App. A.4 regularisation of the downwash singularity at dz -> 0+: clip dz where alpha would exceed the weight
    p.DW_DZ_CLIP = 0.25 * p.PROP_RADIUS * math.sqrt(p.DW_COEFF_1 / p.GRAVITY)
    return p


__all__ = ["DroneModel", "Physics", "drone_params", "G"]
