"""Inner-loop controllers emulating the flight controller (reference
control/low_level/thrust_omega_ctrl.py:9-132, yank_omega_ctrl.py:9-55), batched; PID state
(last_omega, integral) lives in two SoA planes on device."""
from __future__ import annotations

import torch

from .. import _lib


class ThrustOmegaController:
    VARIANT = _lib.CTRL_LQR_OMEGA

    def __init__(self, env):
        self.env = env
        D = env.NUM_TOTAL
        self._a = torch.zeros(D, 4, device=env.device, dtype=env.dtype)
        self._b = torch.zeros(D, 2, device=env.device, dtype=env.dtype)
        self.action = torch.zeros(env.NUM_ENVS, env.NUM_DRONES, 4, device=env.device, dtype=env.dtype)
        self.control_counter = 0

    def reset(self):
        self._a.zero_()
        self._b.zero_()
        self.control_counter = 0

    def pid_struct(self):
        return _lib.PidState(self._a.data_ptr(), self._b.data_ptr())

    @property
    def last_omega(self):
        return self._a[:, :3]

    @property
    def integral_omega_e(self):
        return torch.cat([self._a[:, 3:4], self._b], dim=1)

    def compute_from_obs(self, u, obs):
        """u [E,N,4] = [thrust | yank, body-rate targets]; world->body conversion of obs rates and the
        PID run in one launch (the reference splits this between compute_low_level and
        computeControlFromInput)."""
        env = self.env
        self.control_counter += 1
        _lib.call("mds_lowlevel", env.dtype, env._prm, self.VARIANT, _lib.ptr(_lib.require_cuda(u, "u", env.dtype)),
                  _lib.ptr(_lib.require_cuda(obs, "obs", env.dtype)), self.pid_struct(), _lib.ptr(self.action),
                  env.NUM_TOTAL, _lib.stream_ptr(env.device))
        return self.action


class YankOmegaController(ThrustOmegaController):
    """thrust = current thrust (from obs RPMs) + yank * dt, then the ThrustOmega loop."""
    VARIANT = _lib.CTRL_LQR_YANK

    def __init__(self, env):
        super().__init__(env)
        self.hover_thrust = env.G * env.M
        self.thrust_omega_ctrl = self
