"""Hover-linearised models and the nonlinear SE(3) derivative of the reference
(model/linearized.py, linear_omega.py, linear_yank_omega.py, dynamics.py).

A, B (and the deliberately perturbed Ahat, Bhat) are constants built on the host at init
time exactly as the reference does; ``calc_xdot_from_obs`` / ``QuadrotorDynamics.dynamics_from_obs``
evaluate whole observation batches on device (``mds_xdot_linear`` / ``mds_xdot_nonlinear``),
which is what simulations/CompareModels.py:48-55 loops over."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


class _LinearBase:
    DIM = None

    def __init__(self, env, debug=False):
        self.env = env
        self.mass, self.g = env.M, env.G
        n = self.DIM
        self.A, self.B = np.zeros((n, n)), np.zeros((n, 4))
        self.Ahat, self.Bhat = np.zeros((n, n)), np.zeros((n, 4))
        self.C = np.eye(12)
        self.init_matrices()

    def calc_xdot_from_obs(self, obs, out=None):
        """xdot = A (x - x_eq) + B (u - u_eq) for a batch obs [..., 20] -> [..., DIM] on device.
        DIM 12 follows model/linearized.py:83-104; for DIM 9 / 10 the reference's own method raises
        (quirk B23) and this is the right-sized definition documented in DESIGN.md."""
        env = self.env
        obs = _lib.require_cuda(obs, "obs", env.dtype)
        D = obs.numel() // _lib.OBS_DIM
        if out is None:
            out = torch.empty(*obs.shape[:-1], self.DIM, device=obs.device, dtype=obs.dtype)
        _lib.call("mds_xdot_linear", env.dtype, env._prm, self.DIM, _lib.ptr(obs), _lib.ptr(out), D, _lib.stream_ptr(obs.device))
        return out


class LinearizedModel(_LinearBase):
    """x = [rpy, rates, v, p], u = [f, tx, ty, tz] (model/linearized.py:24-104)."""
    DIM = 12

    def roll_out(self, obs_log, dt=None, out=None):
        """simulations/CompareModels.py:82-95 (``roll_out_linear_system``) for every drone of a logged flight:
        obs_log [T, E, N, 20] (e.g. ``FusedRollout.run(..., obs_log=...)``) -> x [T, E, N, 12], the linear model driven by
        the logged RPMs as zero-order-hold inputs from the first logged state.  Each log interval ``dt`` (default: the
        env's control period) is advanced exactly; the reference integrates the same ODE with scipy's RK45."""
        env = self.env
        obs_log = _lib.require_cuda(obs_log, "obs_log", env.dtype)
        T = int(obs_log.shape[0])
        D = obs_log[0].numel() // _lib.OBS_DIM
        if out is None:
            out = torch.empty(*obs_log.shape[:-1], 12, device=obs_log.device, dtype=obs_log.dtype)
        _lib.call("mds_linear_rollout", env.dtype, env._prm, _lib.ptr(obs_log), float(env.CTRL_TIMESTEP if dt is None else dt), _lib.ptr(out), T, D,
                  _lib.stream_ptr(obs_log.device))
        return out

    def __init__(self, env, debug=False):
        self.Ixx, self.Iyy, self.Izz = env.J[0, 0], env.J[1, 1], env.J[2, 2]
        super().__init__(env, debug)

    def init_matrices(self):
        A, B, g, m = self.A, self.B, self.g, self.mass
        A[0:3, 3:6] = np.eye(3)
        A[9:, 6:9] = np.eye(3)
        A[6, 1], A[7, 0] = g, -g
        B[8, 0] = 1.0 / m
        B[3:6, 1:] = np.diag([1 / self.Ixx, 1 / self.Iyy, 1 / self.Izz])
        self.D = np.zeros((12, 6))
        self.D[:, 2:] = B.copy()
        self.D[7, 1] = self.D[6, 0] = 1.0 / m
        self.Ahat, self.Bhat = A.copy(), B.copy()
        self.Bhat[3:6, 1:] = B[3:6, 1:] * 0.75
        self.Bhat[8, 0] = 1.0 / (m * .75)


class LinearizedOmegaModel(_LinearBase):
    """x = [rpy, v, p], u = [f, wx, wy, wz] (model/linear_omega.py:24-82)."""
    DIM = 9

    def init_matrices(self):
        A, B, g, m = self.A, self.B, self.g, self.mass
        A[6:, 3:6] = np.eye(3)
        A[3, 1], A[4, 0] = g, -g
        B[5, 0] = 1.0 / m
        B[:3, 1:] = np.eye(3)
        self.Ahat, self.Bhat = A.copy(), B.copy()
        self.Ahat[3, 1], self.Ahat[4, 0] = g * 1.2, -g * 1.2
        self.Bhat[5, 0] = 1.0 / (m * 0.8)


class LinearizedYankOmegaModel(_LinearBase):
    """x = [rpy, F, v, p], u = [yank, wx, wy, wz] (model/linear_yank_omega.py:20-81)."""
    DIM = 10

    def __init__(self, env, debug=False):
        self.m, self.n = 10, 4
        super().__init__(env, debug)

    def init_matrices(self):
        A, B, g, m = self.A, self.B, self.g, self.mass
        A[7:, 4:7] = np.eye(3)
        A[4, 1], A[5, 0] = g, -g
        A[6, 3] = 1.0 / m
        B[:3, 1:] = np.eye(3)
        B[3, 0] = 1.0
        self.Ahat, self.Bhat = A.copy(), B.copy()
        self.Ahat[4, 1], self.Ahat[5, 0] = g * 1.2, -g * 1.2
        self.Ahat[6, 3] = 1.0 / (m * 0.8)


class QuadrotorDynamics:
    """Continuous-time rigid-body derivative (model/dynamics.py:83-106).  Hummingbird defaults;
    ``load_env_params`` updates m, g, kf but -- like the reference (finding 6 / quirk B22) -- leaves
    J at the constructor values unless ``update_inertia=True``."""

    def __init__(self, sim_freq, init_position=None, init_rpys=None):
        self.m, self.Jxx, self.Jyy, self.Jzz, self.g = 6.77, 1.05, 1.05, 2.05, 9.81
        self.sim_freq = int(sim_freq)
        self.dt = 1.0 / sim_freq
        self.kf, self.km = 3.16e-10, 7.94e-12
        self.J = np.diag([self.Jxx, self.Jyy, self.Jzz])
        self.env = None

    def load_env_params(self, env, update_inertia=False):
        self.env = env
        self.m, self.g, self.kf = env.M, env.G, env.KF
        self.Ixx, self.Iyy, self.Izz = env.J[0, 0], env.J[1, 1], env.J[2, 2]
        self.sim_freq, self.dt = env.PYB_FREQ, 1.0 / env.PYB_FREQ
        if update_inertia:
            self.Jxx, self.Jyy, self.Jzz = self.Ixx, self.Iyy, self.Izz
        self.J = np.diag([self.Jxx, self.Jyy, self.Jzz])

    def dynamics_from_obs(self, obs, out=None):
        """geo_x_dot_to_linear(dynamics(None, obs_to_geo_model(obs), action_to_input(obs[16:]))) for a
        batch obs [..., 20] -> [..., 12] = (w, wdot, vdot, v), i.e. CompareModels.py:52-54 fused."""
        env = self.env
        if env is None:
            raise _lib.MdsError("call load_env_params(env) first")
        obs = _lib.require_cuda(obs, "obs", env.dtype)
        D = obs.numel() // _lib.OBS_DIM
        if out is None:
            out = torch.empty(*obs.shape[:-1], 12, device=obs.device, dtype=obs.dtype)
        _lib.call("mds_xdot_nonlinear", env.dtype, env._prm, self.Jxx, self.Jyy, self.Jzz, _lib.ptr(obs), _lib.ptr(out), D,
                  _lib.stream_ptr(obs.device))
        return out


__all__ = ["LinearizedModel", "LinearizedOmegaModel", "LinearizedYankOmegaModel", "QuadrotorDynamics"]
