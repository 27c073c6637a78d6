"""Continuous-time LQR controllers of the reference (control/lqr/lqr_controller.py:11-114,
lqr_omega_controller.py:11-119, lqr_YO_controller.py:12-130), batched on device.

The gain K comes from the CARE exactly as in the reference (scipy, init time, host); the
per-step work (error state, u = -K e, caps, inner loop) is one launch of ``mds_lqr_ctrl``."""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as la
import torch

from .. import _lib
from .base_controller import BaseController


class _LqrBase(BaseController):
    VARIANT = None
    DIM = None

    def __init__(self, env, lin_model, low_level=None, debug=False, use_noisy_model=False, Q=None, R=None):
        super().__init__(env)
        self.lin_model = lin_model
        self.low_level = low_level
        self.use_noisy_model = use_noisy_model
        self.A = lin_model.Ahat if use_noisy_model else lin_model.A
        self.B = lin_model.Bhat if use_noisy_model else lin_model.B
        self.Q, self.R = self._weights(Q, R)
        self.compute_gain_matrix()
        E, N = env.NUM_ENVS, env.NUM_DRONES
        self.u = torch.zeros(E, N, 4, device=env.device, dtype=env.dtype)
        self.action = torch.zeros(E, N, 4, device=env.device, dtype=env.dtype)

    def compute_gain_matrix(self):
        self.P = la.solve_continuous_are(self.A, self.B, self.Q, self.R, e=None, s=None, balanced=True)
        self.K = la.solve(self.R, self.B.T @ self.P)
        self._gains = None

    def c_gains(self):
        if self._gains is None:
            g = _lib.LqrGains()
            flat = np.asarray(self.K, dtype=float).reshape(-1)
            for i, v in enumerate(flat):
                g.K[i] = v
            g.dim = self.DIM
            self._gains = g
        return self._gains

    def _pid(self):
        return self.low_level.pid_struct() if self.low_level is not None else _lib.PidState(None, None)

    def compute(self, obs, skip_low_level=False):
        """-> (action [E,N,4] or None, u [E,N,4]) like the reference's compute()."""
        env = self.env
        run_inner = not skip_low_level
        if run_inner and self.VARIANT != _lib.CTRL_LQR_TORQUE and self.low_level is None:
            raise _lib.MdsError("this LQR variant needs a low-level controller unless skip_low_level=True")
        _lib.call("mds_lqr_ctrl", env.dtype, env._prm, self.c_gains(), self.VARIANT, _lib.ptr(self._obs(obs)),
                  _lib.ptr(self._ref_view), _lib.ptr(self.u), _lib.ptr(self.action) if run_inner else None,
                  self._pid(), env.NUM_TOTAL, _lib.stream_ptr(env.device))
        return (self.action if run_inner else None), self.u

    def compute_low_level(self, u, obs, idx=0):
        return self.low_level.compute_from_obs(u, obs)


class LQRController(_LqrBase):
    """12-dim state, torque inputs -> PLUS-frame mixer (lqr_controller.py)."""
    VARIANT, DIM = _lib.CTRL_LQR_TORQUE, 12

    def __init__(self, env, lin_model, Q=None, R=None, debug=False, use_noisy_model=False):
        super().__init__(env, lin_model, None, debug, use_noisy_model, Q, R)

    def _weights(self, Q, R):
        mt = self.env.MAX_THRUST  # Bryson's rule, lqr_controller.py:17-37 (the Q, R arguments are overwritten there)
        r = [1 / mt ** 2, 1 / 0.001 ** 2, 1 / 0.001 ** 2, 1 / 0.001 ** 2]
        q = [1 / (math.pi / 40) ** 2] * 3 + [1 / .25 ** 2] * 3 + [1 / .15 ** 2] * 3 + [1 / .05 ** 2] * 3
        return np.diag(q), np.diag(r)

    def compute(self, obs, skip_low_level=False):
        return super().compute(obs, skip_low_level=False)


class LQROmegaController(_LqrBase):
    """9-dim state, thrust + body-rate inputs, ThrustOmega inner loop (lqr_omega_controller.py)."""
    VARIANT, DIM = _lib.CTRL_LQR_OMEGA, 9

    def __init__(self, env, lin_model, to_controller, debug=False, use_noisy_model=False):
        super().__init__(env, lin_model, to_controller, debug, use_noisy_model)
        self.to_controller = to_controller

    def _weights(self, Q, R):
        mt = self.env.MAX_THRUST  # lqr_omega_controller.py:15-31
        r = [1 / mt ** 2, 100.0, 100.0, 100.0]
        q = [1 / (math.pi / 20) ** 2] * 2 + [1 / (math.pi / 40) ** 2] + [1 / .15 ** 2] * 3 + [1 / .05 ** 2] * 3
        return np.diag(q), np.diag(r)


class LQRYankOmegaController(_LqrBase):
    """10-dim state with thrust as a state, yank + body-rate inputs (lqr_YO_controller.py)."""
    VARIANT, DIM = _lib.CTRL_LQR_YANK, 10

    def __init__(self, env, lin_model, yo_controller, debug=False, use_noisy_model=False, Q=None, R=None):
        super().__init__(env, lin_model, yo_controller, debug, use_noisy_model, Q, R)
        self.yo_controller = yo_controller

    def _weights(self, Q, R):
        env = self.env  # lqr_YO_controller.py:18-39
        if R is None:
            max_yank = (env.MAX_THRUST / env.CTRL_TIMESTEP) / 200
            R = np.diag([1 / max_yank ** 2, 100.0, 100.0, 100.0])
        if Q is None:
            Q = np.diag([1 / (math.pi / 20) ** 2] * 2 + [1 / (math.pi / 40) ** 2]
                        + [1 / (env.MAX_THRUST - env.M * env.G) ** 2] + [1 / .15 ** 2] * 3 + [1 / .05 ** 2] * 3)
        return Q, R
