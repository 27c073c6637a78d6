"""Batched ``CtrlAviary``: E independent environments of N drones on one GPU.

Mirrors the upstream gym-pybullet-drones ``CtrlAviary`` / ``BaseAviary`` contract the
reference is written against (constructor kwargs: reference
simulations/EnvGeometric.py:89-100; ``step(action) -> obs, reward, terminated,
truncated, info``: EnvGeometric.py:431,469; attributes ``M, J, G, KF, KM, L, MAX_RPM,
MAX_THRUST, CTRL_TIMESTEP, ...``: SURVEY.md section 1) with a leading env axis.  The
state lives in HBM as 128-bit SoA planes (include/mds_b200.h) owned by PyTorch; the
update is one launch of ``mds_physics_step`` -- there is no CPU path.

In-place conventions: ``step`` overwrites the state planes and the library-owned
``obs`` buffer it returns (a view, valid until the next ``step``/``reset``/rollout).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib
from ..constants import DroneConstants
from ..enums import DroneModel, Physics


def _rpy_to_quat(rpy: torch.Tensor) -> torch.Tensor:
    """Bullet getQuaternionFromEuler (xyzw), float64 on host -- init-time only."""
    h = rpy.double() * 0.5
    cr, sr = torch.cos(h[..., 0]), torch.sin(h[..., 0])
    cp, sp = torch.cos(h[..., 1]), torch.sin(h[..., 1])
    cy, sy = torch.cos(h[..., 2]), torch.sin(h[..., 2])
    return torch.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                        cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy], dim=-1)


class BatchedCtrlAviary(DroneConstants):
    def __init__(self, drone_model=DroneModel.CF2X, num_drones=1, neighbourhood_radius=np.inf,
                 initial_xyzs=None, initial_rpys=None, physics=Physics.DYN, pyb_freq=240, ctrl_freq=240,
                 gui=False, record=False, obstacles=False, user_debug_gui=True, output_folder="results",
                 num_envs=1, device="cuda", dtype=torch.float32,
                 cf2x_torque_sign=-1, renormalize_quat=False, ground_clamp=None, dw_dz_clip=None, x_frame_mixer=False):
        super().__init__(drone_model, physics, pyb_freq, ctrl_freq, cf2x_torque_sign, renormalize_quat, ground_clamp, dw_dz_clip, x_frame_mixer)
        _lib.load_library()  # fail loudly if the CUDA extension is not built
        if not torch.cuda.is_available():
            raise _lib.MdsError("BatchedCtrlAviary needs a CUDA device: there is no CPU fallback")
        self.NUM_DRONES, self.NUM_ENVS = int(num_drones), int(num_envs)
        if not 1 <= self.NUM_DRONES <= _lib.MAX_DRONES_PER_ENV:
            raise ValueError(f"num_drones must be in [1, {_lib.MAX_DRONES_PER_ENV}]")
        self.device, self.dtype = torch.device(device), dtype
        _lib.suffix(dtype)
        self.NEIGHBOURHOOD_RADIUS = neighbourhood_radius
        self.OUTPUT_FOLDER = output_folder
        E, N = self.NUM_ENVS, self.NUM_DRONES
        D = E * N
        self.NUM_TOTAL = D
        if initial_xyzs is None:
            k = torch.arange(N, dtype=torch.float64) * 4 * self.L
            initial_xyzs = torch.stack([k, k, torch.full((N,), self.COLLISION_H / 2 - self.COLLISION_Z_OFFSET + .1,
                                                         dtype=torch.float64)], dim=-1)
        if initial_rpys is None:
            initial_rpys = torch.zeros(N, 3, dtype=torch.float64)
        self.INIT_XYZS = self._bcast(initial_xyzs, 3)
        self.INIT_RPYS = self._bcast(initial_rpys, 3)
        kw = dict(device=self.device, dtype=dtype)
        self._pos_wx = torch.zeros(D, 4, **kw)
        self._quat = torch.zeros(D, 4, **kw)
        self._vel_wy = torch.zeros(D, 4, **kw)
        self._rpm = torch.zeros(D, 4, **kw)
        self._wz = torch.zeros(D, **kw)
        self._obs = torch.zeros(E, N, _lib.OBS_DIM, **kw)
        self._action = torch.zeros(E, N, 4, **kw)
        self._ext_force = None
        self._reward = torch.full((E,), -1.0, **kw)
        self._false = torch.zeros(E, dtype=torch.bool, device=self.device)
        self._prm = self.c_params()
        self.step_counter = 0
        self.reset()

    # ------------------------------------------------------------------ helpers
    def _bcast(self, a, last):
        t = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a, dtype=torch.float64).cpu()
        E, N = self.NUM_ENVS, self.NUM_DRONES
        if t.shape == (N, last):
            t = t.unsqueeze(0).expand(E, N, last)
        if t.shape != (E, N, last):
            raise ValueError(f"expected shape ({N},{last}) or ({E},{N},{last}), got {tuple(t.shape)}")
        return t.contiguous()

    def _state_struct(self):
        return _lib.State(self._pos_wx.data_ptr(), self._quat.data_ptr(), self._vel_wy.data_ptr(),
                          self._rpm.data_ptr(), self._wz.data_ptr())

    def _stage(self, src, dst, name):
        """Copy a host (numpy / CPU tensor) or device input into a library-owned device buffer."""
        if isinstance(src, np.ndarray):
            src = torch.from_numpy(np.ascontiguousarray(src))
        if not isinstance(src, torch.Tensor):
            src = torch.as_tensor(src)
        if src.numel() != dst.numel():
            raise ValueError(f"{name} has {src.numel()} elements, expected {dst.numel()}")
        dst.copy_(src.reshape(dst.shape), non_blocking=True)
        return dst

    # ------------------------------------------------------------------ gym-style API
    def reset(self, seed=None, options=None):
        E, N = self.NUM_ENVS, self.NUM_DRONES
        self.set_state(self.INIT_XYZS, _rpy_to_quat(self.INIT_RPYS), torch.zeros(E, N, 3), torch.zeros(E, N, 3),
                       torch.zeros(E, N, 4))
        self.step_counter = 0
        return self._obs, {"answer": 42}

    def set_state(self, pos, quat, vel, rpy_rates, last_rpm=None):
        """Overwrite the full state (shapes [E,N,k] or [N,k]); obs is rebuilt with ang_vel = R w."""
        D = self.NUM_TOTAL
        pos, quat = self._bcast(pos, 3).reshape(D, 3), self._bcast(quat, 4).reshape(D, 4)
        vel, w = self._bcast(vel, 3).reshape(D, 3), self._bcast(rpy_rates, 3).reshape(D, 3)
        kw = dict(device=self.device, dtype=self.dtype)
        self._pos_wx.copy_(torch.cat([pos, w[:, 0:1]], dim=1).to(**kw))
        self._quat.copy_(quat.to(**kw))
        self._vel_wy.copy_(torch.cat([vel, w[:, 1:2]], dim=1).to(**kw))
        self._wz.copy_(w[:, 2].to(**kw))
        if last_rpm is not None:
            self._rpm.copy_(self._bcast(last_rpm, 4).reshape(D, 4).to(**kw))
        self.refresh_obs()

    def refresh_obs(self):
        _lib.call("mds_obs_from_state", self.dtype, self._prm, self._state_struct(), _lib.ptr(self._obs),
                  self.NUM_TOTAL, _lib.stream_ptr(self.device))
        return self._obs

    def set_external_force(self, force):
        """Per-drone world-frame force [E,N,3] applied on every physics sub-step until changed
        (stand-in for p.applyExternalForce wind, reference EnvGeometric.py:463-467); None clears."""
        if force is None:
            self._ext_force = None
            return
        if self._ext_force is None:
            self._ext_force = torch.zeros(self.NUM_ENVS, self.NUM_DRONES, 3, device=self.device, dtype=self.dtype)
        self._stage(force, self._ext_force, "force")

    def step(self, action, obs_out=None):
        """action: RPM [E,N,4] (or [N,4] when E == 1), device tensor or host array.
        Returns (obs [E,N,20], reward [E], terminated [E], truncated [E], info).
        ``obs_out``: optional device buffer [E,N,20] that receives the new observation and BECOMES the env's obs
        buffer (lets a caller alternate two buffers so that a device->host copy of step k overlaps step k+1)."""
        if obs_out is not None:
            _lib.require_cuda(obs_out, "obs_out", self.dtype, (self.NUM_ENVS, self.NUM_DRONES, _lib.OBS_DIM))
            self._obs = obs_out
        if isinstance(action, torch.Tensor) and action.is_cuda and action.dtype == self.dtype and action.is_contiguous() \
                and action.numel() == self._action.numel():
            act = action
        else:
            act = self._stage(action, self._action, "action")
        _lib.call("mds_physics_step", self.dtype, self._prm, self._state_struct(), _lib.ptr(act),
                  _lib.ptr(self._ext_force), _lib.ptr(self._obs), self.NUM_ENVS, self.NUM_DRONES,
                  _lib.stream_ptr(self.device))
        self.step_counter += self.PYB_STEPS_PER_CTRL
        return self._obs, self._reward, self._false, self._false, {"answer": 42}

    def step_host(self, action_host: torch.Tensor, obs_host: torch.Tensor):
        """Host-buffer form of ``step``: H2D of a (pinned) action, one launch, D2H of obs into
        ``obs_host`` (pinned).  Asynchronous on the current stream; caller synchronises."""
        self._action.copy_(action_host.reshape(self._action.shape), non_blocking=True)
        self.step(self._action)
        obs_host.copy_(self._obs.reshape(obs_host.shape), non_blocking=True)
        return obs_host

    # ------------------------------------------------------------------ state views (reference attribute names)
    @property
    def obs(self):
        return self._obs

    @property
    def pos(self):
        return self._pos_wx[:, :3].reshape(self.NUM_ENVS, self.NUM_DRONES, 3)

    @property
    def quat(self):
        return self._quat.reshape(self.NUM_ENVS, self.NUM_DRONES, 4)

    @property
    def vel(self):
        return self._vel_wy[:, :3].reshape(self.NUM_ENVS, self.NUM_DRONES, 3)

    @property
    def rpy_rates(self):
        return torch.stack([self._pos_wx[:, 3], self._vel_wy[:, 3], self._wz], dim=-1).reshape(self.NUM_ENVS, self.NUM_DRONES, 3)

    @property
    def rpy(self):
        return self._obs[..., 7:10]

    @property
    def ang_v(self):
        return self._obs[..., 13:16]

    @property
    def last_clipped_action(self):
        return self._rpm.reshape(self.NUM_ENVS, self.NUM_DRONES, 4)

    # ------------------------------------------------------------------ no-op parts of the upstream surface
    def close(self):
        pass

    def render(self, *a, **k):
        pass

    def getPyBulletClient(self):
        return -1

    def getDroneIds(self):
        return np.arange(1, self.NUM_DRONES + 1)

    def _showDroneLocalAxes(self, nth_drone):
        pass

    def state_dict(self):
        """Checkpoint (torch.save-able): SoA planes + obs + counters."""
        return {"pos_wx": self._pos_wx.clone(), "quat": self._quat.clone(), "vel_wy": self._vel_wy.clone(),
                "rpm": self._rpm.clone(), "wz": self._wz.clone(), "obs": self._obs.clone(), "step_counter": self.step_counter}

    def load_state_dict(self, sd):
        for k, t in (("pos_wx", self._pos_wx), ("quat", self._quat), ("vel_wy", self._vel_wy), ("rpm", self._rpm),
                     ("wz", self._wz), ("obs", self._obs)):
            t.copy_(sd[k])
        self.step_counter = int(sd["step_counter"])


CtrlAviary = BatchedCtrlAviary
