"""Geometric SE(3) tracking controller (reference control/geometric.py:7-115) on device."""
from __future__ import annotations

import math

import torch

from .. import _lib
from .base_controller import BaseController


class GeometricControl(BaseController):
    def __init__(self, env):
        super().__init__(env)
        self.m, self.J = env.M, env.J
        # gains and limits: control/geometric.py:14-23 (g = 9.81 on purpose, quirk B2)
        self.Kp, self.Kv, self.KR, self.Kw = 2.25, 3.5, 125.0, 10.0
        self.g = 9.81
        self.max_tilt_angle = 40 * math.pi / 180
        E, N = env.NUM_ENVS, env.NUM_DRONES
        self.action = torch.zeros(E, N, 4, device=env.device, dtype=env.dtype)
        self.u = torch.zeros(E, N, 4, device=env.device, dtype=env.dtype)

    def c_gains(self):
        return _lib.GeoGains(self.Kp, self.Kv, self.KR, self.Kw, self.g, self.max_tilt_angle)

    def compute(self, obs, skip_low_level=False):
        """-> RPM [E,N,4] only, like the reference (quirk B7); ``self.u`` holds [f, tau]."""
        env = self.env
        _lib.call("mds_geometric_ctrl", env.dtype, env._prm, self.c_gains(), _lib.ptr(self._obs(obs)),
                  _lib.ptr(self._ref_view), _lib.ptr(self.action), _lib.ptr(self.u), env.NUM_TOTAL,
                  _lib.stream_ptr(env.device))
        return self.action
