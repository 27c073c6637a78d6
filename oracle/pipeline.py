"""Reference-style closed loops built from the oracle pieces -- TEST INFRASTRUCTURE ONLY.

Each function restates one of the reference's per-step Python loops for ONE environment:

* ``run_tracking``  -- simulations/EnvGeometric.py:404-481 (``do_control`` with the geometric
  or 12-dim LQR controller; the reference's leading ``env.step(zeros)`` at :431 is kept)
* ``run_cbf``       -- simulations/CBFTest.py:269-360 (order 2, LQR-omega nominal) and
  simulations/CBFTestOrd3.py:305-360 (order 3, LQR-yank-omega nominal), including the
  callers' ``nominal_us[:,0] -= M*G`` before the QP and ``+= M*G`` after it (order 2 only)

They are the checker for the fused rollout kernel and the timed CPU baseline of bench.py.
"""
from __future__ import annotations

import numpy as np

from oracle import cbf as ocbf
from oracle import controllers as octl
from oracle.aviary import OracleCtrlAviary


def make_controllers(env, kind):
    """One controller object per drone, as the reference constructs them (simulations/EnvGeometric.py:420-426, CBFTest.py:291-292, CBFTestOrd3.py:294-295)."""
    N = env.NUM_DRONES
    if kind == "geometric":
        return [octl.Geometric(env) for _ in range(N)]
    if kind == "torque12":
        K = octl.lqr_gain(env, "torque12")
        return [octl.Lqr(env, "torque12", None, K=K) for _ in range(N)]
    if kind == "omega9":
        K = octl.lqr_gain(env, "omega9")
        return [octl.Lqr(env, "omega9", octl.ThrustOmegaPid(env), K=K) for _ in range(N)]
    if kind == "yank10":
        K = octl.lqr_gain(env, "yank10")
        return [octl.Lqr(env, "yank10", octl.YankOmegaPid(env), K=K) for _ in range(N)]
    raise ValueError(kind)


def run_tracking(env: OracleCtrlAviary, trajs, ctrl_kind, steps, ctrls=None, t0=0.0, initial_zero_step=False, log=True):
    """-> (obs_log [steps, N, 20], final obs).  One env, per-drone Python loop like the reference."""
    N = env.NUM_DRONES
    ctrls = make_controllers(env, ctrl_kind) if ctrls is None else ctrls
    obs = env.step(np.zeros((N, 4)))[0] if initial_zero_step else env._compute_obs()
    action = np.zeros((N, 4))
    out = []
    t = t0
    for _ in range(steps):
        for j in range(N):
            pos, vel, acc, yaw, om = trajs[j](t)
            ctrls[j].set_desired_trajectory(j, pos, vel, acc, yaw, om)
            if ctrl_kind == "geometric":
                action[j] = ctrls[j].compute(obs[j])
            else:
                action[j], _u = ctrls[j].compute(obs[j])
        obs = env.step(action)[0]
        if log:
            out.append(obs.copy())
        t += env.CTRL_TIMESTEP
    return (np.array(out) if log else None), obs


def run_cbf(env: OracleCtrlAviary, trajs, order, steps, cbf_prm=None, obstacles=None, ctrls=None, t0=0.0, log=True):
    """-> (obs_log, final obs, info) with info = dict(status counts, iterations, min barrier)."""
    N = env.NUM_DRONES
    kind = "omega9" if order == 2 else "yank10"
    ctrls = make_controllers(env, kind) if ctrls is None else ctrls
    if cbf_prm is None:
        cbf_prm = (ocbf.CbfParams(env, 2, 1.0, 0.1, (-2.2, -2.4)) if order == 2
                   else ocbf.CbfParams(env, 3, 2.0, 0.125, (-3.0, -3.6, -5.6)))
    x_obs = None if obstacles is None or len(obstacles) == 0 else [np.asarray(o[:3], float) for o in obstacles]
    r_obs = None if x_obs is None else [float(o[3]) for o in obstacles]
    obs = env._compute_obs()
    action = np.zeros((N, 4))
    nominal = np.zeros((N, 4))
    info = {"status": [0, 0, 0], "iters": 0, "solves": 0}
    out = []
    t = t0
    mg = env.M * env.G
    for _ in range(steps):
        xdes = np.zeros((N, cbf_prm.xdim))
        for j in range(N):
            pos, vel, acc, yaw, om = trajs[j](t)
            ctrls[j].set_desired_trajectory(j, pos, vel, acc, yaw, om)
            _, u = ctrls[j].compute(obs[j], skip_low_level=True)
            nominal[j] = u
            xdes[j] = np.hstack([0, 0, yaw, vel, pos]) if order == 2 else np.hstack([0, 0, yaw, mg, vel, pos])
        nominal[:, 0] -= mg
        u_safe, status, iters = ocbf.safety_filter(cbf_prm, env, obs, xdes, nominal.copy(), x_obs, r_obs)
        info["status"][status] += 1
        info["iters"] += iters
        info["solves"] += int(iters > 0 or status != 0)
        u_safe = np.array(u_safe, dtype=float)
        if order == 2:
            u_safe[:, 0] += mg
        for j in range(N):
            action[j] = ctrls[j].compute_low_level(u_safe[j].copy(), obs[j], j)
        obs = env.step(action)[0]
        if log:
            out.append(obs.copy())
        t += env.CTRL_TIMESTEP
    return (np.array(out) if log else None), obs, info
