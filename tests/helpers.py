"""Shared helpers for the parity tests (oracle-side input generation and comparisons)."""
import numpy as np


def rel_err(a, b, floor=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


def scaled_err(a, b):
    """max |a-b| / max(1, max|b|) -- for vectors whose components share a scale"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def euler_to_quat(rpy):
    """extrinsic xyz (Bullet getQuaternionFromEuler), vectorised, xyzw"""
    h = np.asarray(rpy, dtype=np.float64) * 0.5
    cr, sr, cp, sp, cy, sy = np.cos(h[..., 0]), np.sin(h[..., 0]), np.cos(h[..., 1]), np.sin(h[..., 1]), np.cos(h[..., 2]), np.sin(h[..., 2])
    return np.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy], -1)


def random_state(rng, env, shape, pos_scale=2.0, z_min=None):
    """(pos, quat, vel, body rates, last rpm) with moderate attitudes"""
    rpy = rng.uniform(-0.5, 0.5, (*shape, 3))
    pos = rng.uniform(-pos_scale, pos_scale, (*shape, 3))
    if z_min is not None:
        pos[..., 2] = np.abs(pos[..., 2]) + z_min
    return pos, euler_to_quat(rpy), rng.normal(0, 1, (*shape, 3)), rng.normal(0, 1, (*shape, 3)), rng.uniform(9440.3, env.MAX_RPM, (*shape, 4))
