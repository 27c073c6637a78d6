"""Batched tensor forms of the reference's utils/model_conversions.py helpers.

These are layout conveniences for callers that want the intermediate vectors (the kernels fuse
them and never call these).  Inputs are device tensors with a trailing obs / input axis."""
from __future__ import annotations

import torch

MIN_RPM = 9440.3


def calc_z_thrust(env, obs):
    """utils/model_conversions.py:137-143."""
    return (env.KF * obs[..., 16:20] ** 2).sum(-1)


def obs_to_lin_model(obs, dim=12, env=None):
    """utils/model_conversions.py:20-58."""
    rpy, vel, pos = obs[..., 7:10], obs[..., 10:13], obs[..., 0:3]
    if dim == 12:
        return torch.cat([rpy, obs[..., 13:16], vel, pos], dim=-1)
    if dim == 9:
        return torch.cat([rpy, vel, pos], dim=-1)
    if dim == 10:
        assert env is not None, "env must be provided for 10 dim model to calculate the thrust"
        return torch.cat([rpy, calc_z_thrust(env, obs).unsqueeze(-1), vel, pos], dim=-1)
    raise ValueError("Invalid dim for linear model")


def _quat_to_rot(q):
    q = q / q.norm(dim=-1, keepdim=True)
    x, y, z, w = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1)


def obs_to_geo_model(obs):
    """utils/model_conversions.py:105-114 -> [..., 18] = p3, R9 row-major, v3, w3."""
    return torch.cat([obs[..., 0:3], _quat_to_rot(obs[..., 3:7]), obs[..., 10:13], obs[..., 13:16]], dim=-1)


def _mixer(env, like):
    r, L = env.KM / env.KF, env.L
    return torch.tensor([[1.0, 1.0, 1.0, 1.0], [0.0, L, 0.0, -L], [-L, 0.0, L, 0.0], [-r, r, -r, r]],
                        device=like.device, dtype=like.dtype)


def action_to_input(env, action, cap_rpm=True):
    """utils/model_conversions.py:69-83."""
    if cap_rpm:
        action = action.clamp(0, env.MAX_RPM)
    return (env.KF * action ** 2) @ _mixer(env, action).T


def input_to_action(env, u):
    """utils/model_conversions.py:85-103 (clamps u[..., 0] >= 0 in place, like the reference)."""
    u[..., 0].clamp_(min=0)
    thrusts = u @ torch.linalg.inv(_mixer(env, u)).T
    return (thrusts.clamp(MIN_RPM ** 2 * env.KF, env.MAX_THRUST) / env.KF).sqrt()


def geo_x_dot_to_linear(g):
    """utils/model_conversions.py:124-135."""
    return torch.cat([g[..., 3:6], g[..., 9:12], g[..., 6:9], g[..., 0:3]], dim=-1)
