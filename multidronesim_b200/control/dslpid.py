"""Upstream gym-pybullet-drones ``DSLPIDControl`` on device (external to the reference; it is what
MultiDroneExample.py:85-92,111-114 drives its two drones with).  One object controls ALL drones of a
``BatchedCtrlAviary``; the PID state (integral_pos_e, last_rpy, integral_rpy_e) lives in three SoA planes."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .base_controller import BaseController


class DSLPIDControl(BaseController):
    def __init__(self, env=None, drone_model=None, g=9.8, gain_scale=1.0):
        """``DSLPIDControl(env)``.  ``gain_scale=0.5`` reproduces MultiDroneExample.py:87-92 (all six gain vectors
        halved); the six ``*_COEFF_*`` attributes can also be overwritten, like the reference does."""
        if env is None:
            raise _lib.MdsError("DSLPIDControl needs the batched env (device buffers are sized from it)")
        super().__init__(env)
        self.DRONE_MODEL = env.DRONE_MODEL if drone_model is None else drone_model
        self.GRAVITY = g * env.M
        self.KF, self.KM = env.KF, env.KM
        s = float(gain_scale)
        self.P_COEFF_FOR = s * np.array([.4, .4, 1.25])
        self.I_COEFF_FOR = s * np.array([.05, .05, .05])
        self.D_COEFF_FOR = s * np.array([.2, .2, .5])
        self.P_COEFF_TOR = s * np.array([70000., 70000., 60000.])
        self.I_COEFF_TOR = s * np.array([.0, .0, 500.])
        self.D_COEFF_TOR = s * np.array([20000., 20000., 12000.])
        D, kw = env.NUM_TOTAL, dict(device=env.device, dtype=env.dtype)
        self._a, self._b, self._c = torch.zeros(D, 4, **kw), torch.zeros(D, 4, **kw), torch.zeros(D, **kw)
        E, N = env.NUM_ENVS, env.NUM_DRONES
        self.action = torch.zeros(E, N, 4, **kw)
        self.pos_e = torch.zeros(E, N, 3, **kw)
        self._target = torch.zeros(D, 12, **kw)
        self.control_counter = 0

    def reset(self):
        for t in (self._a, self._b, self._c):
            t.zero_()
        self.control_counter = 0

    def c_gains(self):
        g = _lib.DslPidGains()
        for name, arr in (("p_for", self.P_COEFF_FOR), ("i_for", self.I_COEFF_FOR), ("d_for", self.D_COEFF_FOR),
                          ("p_tor", self.P_COEFF_TOR), ("i_tor", self.I_COEFF_TOR), ("d_tor", self.D_COEFF_TOR)):
            for i in range(3):
                getattr(g, name)[i] = float(arr[i])
        return g

    def state_struct(self):
        return _lib.DslPidState(self._a.data_ptr(), self._b.data_ptr(), self._c.data_ptr())

    def computeControlFromState(self, control_timestep, state, target_pos, target_rpy=None, target_vel=None, target_rpy_rates=None):
        """Batched upstream signature: ``state`` = obs [E,N,20]; targets [E,N,3] / [N,3] / [3] (None = zeros).
        Returns (rpm [E,N,4], pos_e [E,N,3], yaw error [E,N]) like upstream's (rpm, pos_e, yaw_e)."""
        env = self.env
        if abs(control_timestep - env.CTRL_TIMESTEP) > 1e-12:
            raise _lib.MdsError("control_timestep must equal env.CTRL_TIMESTEP (it is a launch constant)")
        tgt = self._target.view(env.NUM_ENVS, env.NUM_DRONES, 12)
        for k, v in enumerate((target_pos, target_rpy, target_vel, target_rpy_rates)):
            if v is None:
                tgt[..., 3 * k:3 * k + 3].zero_()
            else:
                self._assign(tgt[..., 3 * k:3 * k + 3], v)
        obs = self._obs(state)
        self.control_counter += 1
        _lib.call("mds_dslpid_ctrl", env.dtype, env._prm, self.c_gains(), _lib.ptr(obs), _lib.ptr(self._target), self.state_struct(),
                  _lib.ptr(self.action), _lib.ptr(self.pos_e), env.NUM_TOTAL, _lib.stream_ptr(env.device))
        return self.action, self.pos_e, tgt[..., 5] - obs[..., 9]

    def compute(self, obs, skip_low_level=False):
        """BaseController form: track the current reference (pos, vel, yaw, yaw rate) -> RPM [E,N,4]."""
        ref = self._ref_view.view(self.env.NUM_ENVS, self.env.NUM_DRONES, _lib.REF_DIM)
        zeros = torch.zeros_like(ref[..., 0:2])
        rpm, _, _ = self.computeControlFromState(self.env.CTRL_TIMESTEP, obs, ref[..., 0:3], torch.cat([zeros, ref[..., 9:10]], -1),
                                                 ref[..., 3:6], torch.cat([zeros, ref[..., 10:11]], -1))
        return rpm
