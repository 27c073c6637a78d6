"""Controller interface of the reference (control/base_controller.py:1-12), batched.

One controller object drives ALL drones of a ``BatchedCtrlAviary`` (the reference keeps
one Python object per drone and loops; simulations/EnvGeometric.py:435-451).  References
live in a device buffer ``ref [D, 11] = pos3, vel3, acc3, yaw, yaw_rate``."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


class BaseController:
    def __init__(self, env):
        self.env = env
        D = env.NUM_TOTAL
        self.ref = torch.zeros(D, _lib.REF_DIM, device=env.device, dtype=env.dtype)
        self._ref_view = self.ref

    def _assign(self, target, x):
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x, dtype=np.float64))
        x = x.to(device=self.env.device, dtype=self.env.dtype)
        if x.numel() == target.numel():
            x = x.reshape(target.shape)
        target.copy_(x.expand_as(target))

    def set_desired_trajectory(self, robot_idx, desired_pos, desired_vel, desired_acc, desired_yaw, desired_omega):
        """Set the reference.  ``robot_idx`` None / slice(None): every drone (arrays [D,3] / [D], or one
        [3] / scalar broadcast to all); an int: that drone in every environment ([3] / scalar or
        [E,3] / [E])."""
        ref = self.ref.view(self.env.NUM_ENVS, self.env.NUM_DRONES, _lib.REF_DIM)
        sel = ref if robot_idx is None or isinstance(robot_idx, slice) else ref[:, int(robot_idx)]
        self._assign(sel[..., 0:3], desired_pos)
        self._assign(sel[..., 3:6], desired_vel)
        self._assign(sel[..., 6:9], desired_acc)
        self._assign(sel[..., 9], desired_yaw)
        self._assign(sel[..., 10], desired_omega)
        self._ref_view = self.ref

    def set_reference(self, ref):
        """Zero-copy form: use a device buffer [D, 11] (e.g. ``TrajectorySet.eval(t)``) as the reference."""
        self._ref_view = _lib.require_cuda(ref, "ref", self.env.dtype, (self.env.NUM_TOTAL, _lib.REF_DIM))

    def compute(self, obs, skip_low_level=False):
        raise NotImplementedError

    def _obs(self, obs):
        return _lib.require_cuda(obs, "obs", self.env.dtype, (self.env.NUM_ENVS, self.env.NUM_DRONES, _lib.OBS_DIM))
