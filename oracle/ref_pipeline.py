"""The reference's OWN classes in its own per-step loop -- TEST INFRASTRUCTURE / CPU baseline only.

``run_cbf_reference`` is simulations/CBFTestOrd3.py:272-360 (``do_control`` with ``--controller lqr`` and a
``DroneQPTracker``) and, for order 2, simulations/CBFTest.py:269-360, with every object the loop touches constructed
from the reference's packages (imported through oracle/ref_import.py from ``/root/reference`` or from the copy
``oracle/build_ref.py`` leaves in ``oracle/_ref``):

    trajectories.Lemniscate, control.LQRYankOmegaController / LQROmegaController (+ YankOmegaController /
    ThrustOmegaController inner loops), model.LinearizedYankOmegaModel / LinearizedOmegaModel,
    cbf.DroneCBF._build_ineq_const (dense Jacobian / Hessian builder), cbf.DroneQPTracker.compute_control

Only two pieces are not the reference's, because they are not in its tree: ``env.step`` (oracle/aviary.py, upstream
gym-pybullet-drones restated) and ``cvxopt.solvers.qp`` (oracle/qp.py behind the stand-in of ref_import).
"""
from __future__ import annotations

import contextlib
import io
import os

import numpy as np

from oracle import ref_import

_REF = None


def reference_root():
    """``/root/reference`` if present, else the shipped copy ``oracle/_ref``, else None."""
    for root in (os.environ.get("MDS_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")):
        if root and os.path.isdir(os.path.join(root, "cbf")):
            return root
    return None


def load_reference():
    global _REF
    if _REF is None:
        root = reference_root()
        if root is None:
            raise RuntimeError("the reference's packages are neither at /root/reference nor in oracle/_ref (python -m oracle.build_ref)")
        ref_import.REFERENCE_ROOT = root
        with contextlib.redirect_stdout(io.StringIO()):
            _REF = ref_import.load()
    return _REF


class ReferenceLoop:
    """One environment driven by the reference's objects; keeps them between calls so a run can be advanced in pieces."""

    def __init__(self, env, order, traj_specs, obstacles=None):
        ref = load_reference()
        self.env, self.order, N = env, order, env.NUM_DRONES
        with contextlib.redirect_stdout(io.StringIO()):
            if order == 3:
                self.models = [ref.model.LinearizedYankOmegaModel(env) for _ in range(N)]
                self.ctrl = [ref.control.LQRYankOmegaController(env, self.models[j], ref.control.YankOmegaController(env), use_noisy_model=False)
                             for j in range(N)]
                cbf = ref.cbf.DroneCBF(env, self.models, safety_radius=0.125, zscale=2, order=3, cbf_poles=np.array([-3.0, -3.6, -5.6]))
                self.tracker = ref.cbf.DroneQPTracker(cbf, num_robots=N, xdim=10, env=env, order=3)
            else:
                self.models = [ref.model.LinearizedOmegaModel(env) for _ in range(N)]
                self.ctrl = [ref.control.LQROmegaController(env, self.models[j], ref.control.ThrustOmegaController(env), use_noisy_model=False)
                             for j in range(N)]
                cbf = ref.cbf.DroneCBF(env, self.models, safety_radius=0.1, zscale=1, order=2, cbf_poles=np.array([-2.2, -2.4]))
                self.tracker = ref.cbf.DroneQPTracker(cbf, num_robots=N, xdim=9, env=env, order=2)
        self.trajs = [ref.traj.Lemniscate(**sp) for sp in traj_specs]
        if obstacles is None or len(obstacles) == 0:
            self.x_obs, self.r_obs = None, None
        else:  # (order, 3) per obstacle: position, then zero derivatives (simulations/CBFTest.py:421-424)
            self.x_obs = np.array([np.vstack([np.asarray(o[:3], float)] + [np.zeros(3)] * (order - 1)) for o in obstacles])
            self.r_obs = [float(o[3]) for o in obstacles]
        self.obs = env._compute_obs()
        self.t = 0.0
        self.qp_fallbacks = 0

    def adopt_inner_loop_state(self, oracle_ctrls):
        """Continue from a run of the oracle port (oracle/pipeline.py): copy its rate-PID state into the reference's inner loops."""
        for mine, theirs in zip(self.ctrl, oracle_ctrls):
            toc = mine.yo_controller.thrust_omega_ctrl if self.order == 3 else mine.to_controller
            pid = theirs.low.inner if self.order == 3 else theirs.low
            toc.last_omega, toc.integral_omega_e = np.array(pid.last_omega, float), np.array(pid.integral, float)

    def run(self, steps):
        env, N, mg = self.env, self.env.NUM_DRONES, self.env.M * self.env.G
        action, nominal = np.zeros((N, 4)), np.zeros((N, 4))
        xdim = 10 if self.order == 3 else 9
        sink = io.StringIO()
        for _ in range(steps):
            xdes = np.zeros((N, xdim))
            for j in range(N):
                pos, vel, acc, yaw, omega = self.trajs[j](self.t)
                self.ctrl[j].set_desired_trajectory(j, desired_pos=pos, desired_vel=vel, desired_acc=acc, desired_yaw=yaw, desired_omega=omega)
                action[j, :], u = self.ctrl[j].compute(self.obs[j], skip_low_level=True)
                nominal[j, :] = u
                xdes[j] = np.hstack([0, 0, yaw, mg, vel, pos]) if self.order == 3 else np.hstack([0, 0, yaw, vel, pos])
            nominal[:, 0] = nominal[:, 0] - mg
            with contextlib.redirect_stdout(sink):  # the tracker prints when the QP fails and returns the nominal input
                u_safe = np.array(self.tracker.compute_control(self.obs, xdes, nominal, x_obs=self.x_obs, obs_r_list=self.r_obs), dtype=float)
            if self.order == 2:
                u_safe[:, 0] = u_safe[:, 0] + mg
            for j in range(N):
                action[j, :] = self.ctrl[j].compute_low_level(u_safe[j, :], self.obs[j], j)
            self.obs = env.step(action)[0]
            self.t += env.CTRL_TIMESTEP
        self.qp_fallbacks += sink.getvalue().count("cannot find")
        return self.obs
