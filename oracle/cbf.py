"""Exponential-CBF constraint builder and safety filter -- TEST INFRASTRUCTURE ONLY.

numpy fp64 restatement of the reference's cbf/cbf.py + cbf/qptracker.py.  The
reference materialises dense 2*xdim Jacobians / Hessians / a 20x20x20 tensor per
pair (cbf/cbf.py:194-283); only the three position components and the rows of
the hover-linearised A, B that touch them survive, so this file evaluates the
same Lie derivatives in closed form (derivation in DESIGN.md section "CBF rows").
Pinned row-for-row against the imported reference's ``_build_ineq_const`` through
tests/golden/cbf_rows.npz (oracle/make_golden.py; tests/test_oracle_golden.py).

Row order of G u <= h (cbf/cbf.py:308-367):
  pairs (i<j, lexicographic) | +I box, -I box (2*4N) | force bound (2N, order 3) | obstacles (i*N_obs + j)
"""
from __future__ import annotations

import math

import numpy as np

from oracle import conversions as cv
from oracle.qp import solve_qp


def cbf_gain(poles):
    """``scipy.signal.place_poles`` on the integrator chain (cbf/cbf.py:119-124);
    single input => unique gain = reversed coefficients of prod(s - p_k)."""
    c = np.poly(np.asarray(poles, dtype=float))          # s^n + c1 s^(n-1) + ... + cn
    return np.real(c[1:][::-1]).copy()                   # [cn, ..., c1]


class CbfParams:
    """What ``DroneCBF.__init__`` precomputes (cbf/cbf.py:545-580)."""

    def __init__(self, env, order=2, zscale=2.0, safety_radius=1.0, cbf_poles=(-2.2, -2.4),
                 omega_max=(10.0, 10.0, 10.0)):
        if order not in (2, 3):
            raise ValueError("order must be 2 (thrust-omega, xdim 9) or 3 (yank-omega, xdim 10)")
        if len(cbf_poles) != order:
            raise AssertionError("Number of specified CBF poles does not match order")
        self.order, self.xdim = order, 9 if order == 2 else 10
        self.c, self.rs = float(zscale), float(safety_radius)
        self.K = cbf_gain(cbf_poles)
        self.m, self.g = env.M, env.G
        self.Fmin, self.Fmax = -env.M * env.G, env.MAX_THRUST
        u0 = env.MAX_THRUST if order == 2 else (env.MAX_THRUST / env.CTRL_TIMESTEP) / 100
        self.umax = np.array([u0, omega_max[0], omega_max[1], omega_max[2]], dtype=float)


def _split(prm, x):
    """(roll, pitch, F, v3, p3) of a linear-model state of dim 9 / 10 (App. D)."""
    if prm.order == 2:
        return x[0], x[1], 0.0, x[3:6], x[6:9]
    return x[0], x[1], x[3], x[4:7], x[7:10]


def pair_row(prm, xi, xj, xi_des, xj_des, Ds, cylinder=False):
    """One ECBF row: returns (a3, rhs) with a = LgL_f^{r-1}h[:3] on drone i's
    input block (columns [u0, wx, wy]) and rhs = Kcbf . [h, hdot, ..] + L_f^r h.

    Follows cbf/cbf.py:135-178 (custom_hdots, incl. quirk B12) and :194-283
    (custom_control_affine_terms)."""
    c4 = math.inf if cylinder else prm.c ** 4   # vertical cylinder = zscale -> infinity (builder extension, parity unpinned)
    g, m = prm.g, prm.m
    ri, pi_, Fi, vi, pi3 = _split(prm, xi)
    rj, pj_, Fj, vj, pj3 = _split(prm, xj)
    rid, pid, Fid, vid, _ = _split(prm, xi_des)
    rjd, pjd, Fjd, vjd, _ = _split(prm, xj_des)
    ex, ey, ez = pi3 - pj3
    dv = (vi - vid) - (vj - vjd)                             # relative velocity error
    acc_i = np.array([g * (pi_ - pid), -g * (ri - rid), (Fi - Fid) / m if prm.order == 3 else 0.0])
    acc_j = np.array([g * (pj_ - pjd), -g * (rj - rjd), (Fj - Fjd) / m if prm.order == 3 else 0.0])
    da = acc_i - acc_j                                      # relative linearised acceleration
    rho = ex * ex + ey * ey
    d = np.array([4 * ex * rho, 4 * ey * rho, 4 * ez ** 3 / c4])
    Hxx, Hxy, Hyy, Hzz = 12 * ex * ex + 4 * ey * ey, 8 * ex * ey, 4 * ex * ex + 12 * ey * ey, 12 * ez * ez / c4
    Hdv = np.array([Hxx * dv[0] + Hxy * dv[1], Hxy * dv[0] + Hyy * dv[1], Hzz * dv[2]])
    h0 = rho * rho + (0.0 if cylinder else (ez / prm.c) ** 4) - Ds ** 4
    h1 = d @ dv
    if prm.order == 2:
        Lf = d @ da + dv @ Hdv
        a = np.array([d[2] / m, 0.0, 0.0])
        return a, prm.K[0] * h0 + prm.K[1] * h1 + Lf
    # order 3.  hdots[2] uses the reference's hard-coded indices 6,7,8 (vz, px, py of the 10-dim state)
    h2 = (d[1] * da[0] + d[2] * da[1]
          + Hxx * da[2] ** 2 + 2 * Hxy * da[2] * dv[0] + Hyy * dv[0] ** 2 + Hzz * dv[1] ** 2)
    q = np.array([24 * ex * dv[0] ** 2 + 16 * ey * dv[0] * dv[1] + 8 * ex * dv[1] ** 2,
                  8 * ey * dv[0] ** 2 + 16 * ex * dv[0] * dv[1] + 24 * ey * dv[1] ** 2,
                  24 * ez / c4 * dv[2] ** 2])
    Lf = 3.0 * (da @ Hdv) + q @ dv
    a = np.array([d[2] / m, -g * d[1], g * d[0]])
    return a, prm.K[0] * h0 + prm.K[1] * h1 + prm.K[2] * h2 + Lf


def build_ineq(prm, x, xdes, x_obs=None, obs_r=None, allow_extra_obstacles=False, do_state_bounds=True):
    """Dense (G, h) exactly as ``CBF._build_ineq_const`` (cbf/cbf.py:308-367)."""
    x, xdes = np.asarray(x, float), np.asarray(xdes, float)
    N = x.shape[0]
    rows_G, rows_h = [], []
    for i in range(N - 1):
        for j in range(i + 1, N):
            a, rhs = pair_row(prm, x[i], x[j], xdes[i], xdes[j], 2 * prm.rs)
            g_row = np.zeros(4 * N)
            g_row[4 * i:4 * i + 3] = -a
            g_row[4 * j:4 * j + 3] = a          # cbf.py:299-300 reuses drone i's block, sign flipped
            rows_G.append(g_row)
            rows_h.append(rhs)
    eye = np.eye(4 * N)
    rows_G += list(eye) + list(-eye)            # cbf.py:400-412
    rows_h += list(np.tile(prm.umax, 2 * N))
    if prm.order == 3 and do_state_bounds:      # cbf.py:446-476: acts on column 4i+3 (wz), quirk kept; only with do_state_bounds
        for i in range(N):
            gp, gm = np.zeros(4 * N), np.zeros(4 * N)
            gp[4 * i + 3], gm[4 * i + 3] = 1.0, -1.0
            rows_G += [gp, gm]
            rows_h += [prm.K[-1] * (prm.Fmax - x[i][3]), prm.K[-1] * (x[i][3] - prm.Fmin)]
    if x_obs is not None and obs_r is not None:
        if len(x_obs) != len(obs_r):
            raise AssertionError("The lists for Obstacle positions and radii must have the same length")
        if len(x_obs) > N and not allow_extra_obstacles:
            # builder extension when allowed: the row itself only ever uses drone i's own model
            raise IndexError("reference indexes agent blocks by obstacle id: N_obs <= N (quirk B14)")
        for i in range(N):
            for j in range(len(x_obs)):
                xo = np.zeros(prm.xdim)
                xo[-3:] = np.asarray(x_obs[j], float).reshape(-1, 3)[0]
                a, rhs = pair_row(prm, x[i], xo, xdes[i], xo, prm.rs + abs(obs_r[j]), cylinder=obs_r[j] < 0)
                g_row = np.zeros(4 * N)
                g_row[4 * i:4 * i + 3] = -a
                rows_G.append(g_row)
                rows_h.append(rhs)
    return np.array(rows_G).reshape(-1, 4 * N), np.array(rows_h, dtype=float)


def safety_filter(prm, env, obs, xdes, u_nominal, x_obs=None, obs_r=None):
    """``DroneQPTracker.compute_control`` (cbf/qptracker.py:22-34) ->
    (u_safe (N,4), status, iterations).  status 0 optimal; otherwise the nominal
    input is returned, as the reference's except-branch does (:30-34,105-114)."""
    obs = np.asarray(obs, float)
    N = obs.shape[0]
    x = np.array([cv.obs_to_lin_model(obs[i], prm.xdim, env) for i in range(N)])
    Gm, h = build_ineq(prm, x, xdes, x_obs, obs_r)
    uhat = np.asarray(u_nominal, float).reshape(4 * N)
    u, _lam, status, iters = solve_qp(np.eye(4 * N), -uhat, Gm, h)
    if status != 0:
        return np.asarray(u_nominal, float).reshape(N, 4).copy(), status, iters
    return u.reshape(N, 4), 0, iters
