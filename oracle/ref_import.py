"""Import the reference's own Python packages in the BUILD container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  ``/root/reference`` does not
exist on the GPU box, so this module is used solely by ``oracle/make_golden.py``
(fixture generation) and by ``-m "not gpu"`` tests that are skipped when the
reference tree is absent.

The reference imports four third-party packages that are not installed here
(SURVEY.md App. C): ``gym_pybullet_drones`` (enums + ``BaseControl``),
``pybullet``, ``cvxopt`` and ``matplotlib``.  We register in-memory stand-ins
for exactly the names the reference touches at import time:

* ``gym_pybullet_drones.utils.enums.DroneModel`` / ``.envs.BaseAviary.DroneModel``
  (used by control/low_level/thrust_omega_ctrl.py:5,34,47,54)
* ``gym_pybullet_drones.control.BaseControl.BaseControl``
  (base class of ThrustOmegaController, thrust_omega_ctrl.py:9,33)
* ``cvxopt.matrix`` / ``cvxopt.solvers.qp`` (cbf/qptracker.py:7,106) -- the QP
  is delegated to ``oracle.qp.solve_qp`` (exact active-set, KKT-certified)
"""
from __future__ import annotations

import enum
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("MDS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cbf"))


class DroneModel(enum.Enum):
    CF2X = "cf2x"
    CF2P = "cf2p"
    RACE = "racer"


class Physics(enum.Enum):
    PYB = "pyb"
    DYN = "dyn"
    PYB_GND = "pyb_gnd"
    PYB_DRAG = "pyb_drag"
    PYB_DW = "pyb_dw"
    PYB_GND_DRAG_DW = "pyb_gnd_drag_dw"
    # new composite defined by this project; not upstream
    DYN_GND_DRAG_DW = "dyn_gnd_drag_dw"


class _BaseControl:
    """Stand-in for gym_pybullet_drones.control.BaseControl."""

    def __init__(self, drone_model, g: float = 9.8):
        self.DRONE_MODEL = drone_model
        self.GRAVITY = g * 0.027
        self.KF = 3.16e-10
        self.KM = 7.94e-12
        self.reset()

    def reset(self):
        self.control_counter = 0


def _fake_cvxopt_qp(P, q, G, h, *a, **k):
    from oracle.qp import solve_qp

    P = np.asarray(P, dtype=float)
    q = np.asarray(q, dtype=float).reshape(-1)
    G = np.asarray(G, dtype=float)
    h = np.asarray(h, dtype=float).reshape(-1)
    x, _lam, status, _it = solve_qp(P, q, G, h)
    if status != 0:
        raise ValueError("oracle QP: infeasible or iteration cap (status %d)" % status)
    return {"x": x.reshape(-1, 1), "status": "optimal"}


_installed = False


def install_stubs() -> None:
    """Register the stand-in modules and put the reference tree on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    gpd = mod("gym_pybullet_drones")
    envs = mod("gym_pybullet_drones.envs")
    base_av = mod("gym_pybullet_drones.envs.BaseAviary")
    ctrl_av = mod("gym_pybullet_drones.envs.CtrlAviary")
    gutils = mod("gym_pybullet_drones.utils")
    gutils_utils = mod("gym_pybullet_drones.utils.utils")
    genums = mod("gym_pybullet_drones.utils.enums")
    gctrl = mod("gym_pybullet_drones.control")
    gbase = mod("gym_pybullet_drones.control.BaseControl")
    gdsl = mod("gym_pybullet_drones.control.DSLPIDControl")
    gpd.envs, gpd.utils, gpd.control = envs, gutils, gctrl
    genums.DroneModel = DroneModel
    genums.Physics = Physics
    base_av.DroneModel = DroneModel
    base_av.Physics = Physics
    ctrl_av.CtrlAviary = object
    gutils_utils.sync = lambda *a, **k: None
    gutils_utils.str2bool = lambda s: str(s).lower() in ("1", "true", "yes")
    gbase.BaseControl = _BaseControl
    gdsl.DSLPIDControl = _BaseControl

    mod("pybullet")
    mpl = mod("matplotlib")
    mpl.pyplot = mod("matplotlib.pyplot")

    cvx = mod("cvxopt")
    cvx.matrix = lambda a, *x, **k: np.asarray(a, dtype=float)
    solvers = types.SimpleNamespace(options={}, qp=_fake_cvxopt_qp)
    cvx.solvers = solvers

    if REFERENCE_ROOT not in sys.path:
        # APPEND, never prepend: the reference has top-level packages named
        # ``utils``/``control``/``model`` that must not shadow anything of ours.
        sys.path.append(REFERENCE_ROOT)
    _installed = True


def load():
    """Return a namespace with the reference packages (imports them once)."""
    import warnings

    install_stubs()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import utils as r_utils  # noqa: F401  (reference package)
        import model as r_model
        import trajectories as r_traj
        import control as r_control
        import cbf as r_cbf
        import obstacles as r_obstacles
        import utils.model_conversions as r_conv
    return types.SimpleNamespace(
        utils=r_utils, conv=r_conv, model=r_model, traj=r_traj,
        control=r_control, cbf=r_cbf, obstacles=r_obstacles,
        DroneModel=DroneModel, Physics=Physics,
    )
