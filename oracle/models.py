"""Linearised hover models and the nonlinear SE(3) derivative -- TEST INFRASTRUCTURE ONLY.

numpy fp64 restatement of the reference's model/*.py; pinned against the
imported reference by tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import numpy as np

from oracle import conversions as cv


def linear_model_matrices(env, kind):
    """(A, B, Ahat, Bhat) for kind in {'torque12', 'omega9', 'yank10'}.

    * torque12: model/linearized.py:52-79, x=[rpy, w, v, p], u=[f, tx, ty, tz]
    * omega9:   model/linear_omega.py:46-61, x=[rpy, v, p], u=[f, wx, wy, wz]
    * yank10:   model/linear_yank_omega.py:45-57, x=[rpy, F, v, p], u=[Fdot, wx, wy, wz]
    """
    g, m = env.G, env.M
    if kind == "torque12":
        A, B = np.zeros((12, 12)), np.zeros((12, 4))
        A[0:3, 3:6] = np.eye(3)
        A[9:12, 6:9] = np.eye(3)
        A[6, 1], A[7, 0] = g, -g
        B[8, 0] = 1.0 / m
        B[3, 1], B[4, 2], B[5, 3] = 1 / env.J[0, 0], 1 / env.J[1, 1], 1 / env.J[2, 2]
        Ah, Bh = A.copy(), B.copy()
        Bh[3:6, 1:] = B[3:6, 1:] * 0.75
        Bh[8, 0] = 1.0 / (m * 0.75)
    elif kind == "omega9":
        A, B = np.zeros((9, 9)), np.zeros((9, 4))
        A[6:9, 3:6] = np.eye(3)
        A[3, 1], A[4, 0] = g, -g
        B[5, 0] = 1.0 / m
        B[0:3, 1:4] = np.eye(3)
        Ah, Bh = A.copy(), B.copy()
        Ah[3, 1], Ah[4, 0] = g * 1.2, -g * 1.2
        Bh[5, 0] = 1.0 / (m * 0.8)
    elif kind == "yank10":
        A, B = np.zeros((10, 10)), np.zeros((10, 4))
        A[7:10, 4:7] = np.eye(3)
        A[4, 1], A[5, 0] = g, -g
        A[6, 3] = 1.0 / m
        B[0:3, 1:4] = np.eye(3)
        B[3, 0] = 1.0
        Ah, Bh = A.copy(), B.copy()
        Ah[4, 1], Ah[5, 0] = g * 1.2, -g * 1.2
        Ah[6, 3] = 1.0 / (m * 0.8)
    else:
        raise ValueError(kind)
    return A, B, Ah, Bh


def xdot_linear12_from_obs(env, obs):
    """model/linearized.py:83-104 -- the only ``calc_xdot`` variant that runs in
    the reference (quirk B23): x_eq = [0..0, p], u_eq = [m g, 0, 0, 0]."""
    A, B, _, _ = linear_model_matrices(env, "torque12")
    x = cv.obs_to_lin_model(obs, 12)
    u = cv.action_to_input(env, obs[16:20])
    xe = np.zeros(12)
    xe[9:] = x[9:]
    return A @ (x - xe) + B @ (u - np.array([env.M * env.G, 0, 0, 0]))


def xdot_linear_generic(env, obs, kind):
    """Builder-defined right-sized xdot = A (x - x_eq) + B (u - u_eq) for the 9/10-dim
    models (quirk B23: the reference's own method raises; parity is against its
    MATRICES).  Inputs: omega9 u=[f, body rates]; yank10 u=[0 yank, body rates]."""
    A, B, _, _ = linear_model_matrices(env, kind)
    dim = {"omega9": 9, "yank10": 10}[kind]
    x = cv.obs_to_lin_model(obs, dim, env)
    xe = np.zeros(dim)
    xe[-3:] = x[-3:]
    Rm = cv.quat_to_rot(obs[3:7])
    wb = Rm.T @ obs[13:16]
    f = cv.calc_z_thrust(env, obs)
    if kind == "omega9":
        u, ue = np.array([f, *wb]), np.array([env.M * env.G, 0, 0, 0])
    else:
        xe[3] = env.M * env.G
        u, ue = np.array([0.0, *wb]), np.zeros(4)
    return A @ (x - xe) + B @ (u - ue)


def xdot_nonlinear(env, state18, u, use_env_inertia=False):
    """model/dynamics.py:83-106 -> (v, w, vdot, wdot).  Quirk B22/finding 6:
    after ``load_env_params`` m, g come from the env but J stays at the Hummingbird
    default diag(1.05, 1.05, 2.05) unless ``use_env_inertia``."""
    J = env.J if use_env_inertia else np.diag([1.05, 1.05, 2.05])
    Rm = np.asarray(state18[3:12], dtype=float).reshape(3, 3)
    v, w = state18[12:15], state18[15:18]
    e3 = np.array([0.0, 0.0, 1.0])
    vdot = Rm @ (u[0] * e3) / env.M - env.G * e3
    wdot = np.linalg.inv(J) @ (u[1:4] - np.cross(w, J @ w))
    return np.concatenate([v, w, vdot, wdot])


def xdot_nonlinear_from_obs(env, obs, use_env_inertia=False):
    """simulations/CompareModels.py:52-54: action_to_input -> dynamics -> geo_x_dot_to_linear."""
    u = cv.action_to_input(env, obs[16:20])
    return cv.geo_x_dot_to_linear(xdot_nonlinear(env, cv.obs_to_geo_model(obs), u, use_env_inertia))


def roll_out_linear_system(env, observations, obs_ts, rtol=1e-11, atol=1e-13):
    """simulations/CompareModels.py:82-95: integrate the 12-dim linear model from the first logged state with the logged
    RPMs as inputs ("closest observation in the past", i.e. zero-order hold), sampled at ``obs_ts``.  The reference calls
    ``solve_ivp`` with its default tolerances (rtol 1e-3); the oracle integrates interval by interval with tight
    tolerances so that it can pin the device kernel's exact per-interval update.  -> y [T, 12]."""
    from scipy.integrate import solve_ivp
    A, B, _, _ = linear_model_matrices(env, "torque12")
    ueq = np.array([env.M * env.G, 0, 0, 0])
    x = cv.obs_to_lin_model(observations[0], 12)
    out = [x.copy()]
    for k in range(len(obs_ts) - 1):
        u = cv.action_to_input(env, observations[k][16:20]) - ueq

        def f(t, xx):
            xe = np.zeros(12)
            xe[9:] = xx[9:]
            return A @ (xx - xe) + B @ u
        res = solve_ivp(f, [obs_ts[k], obs_ts[k + 1]], x, rtol=rtol, atol=atol)
        x = res.y[:, -1]
        out.append(x.copy())
    return np.array(out)
