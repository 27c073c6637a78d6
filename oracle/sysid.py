"""Recursive-least-squares model learning of the reference's decentralised LQR, restated per drone in numpy fp64.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
  control/dlqr/decentralized_lqr_omega.py:110-139   (theta_update2, theta_update; m = 9)
  control/dlqr/decentralized_lqr_yank_omega.py:112-126 (theta_update; m = 10)
  control/dlqr/decentralized_lqr.py:157-240          (theta_update, est_x_dot, approx_theta_update, project_theta; m = 12)
  control/dlqr/decentralized_yolqr_crazyflie.py:228-290 (est_x_dot, project_theta, approx_theta_update; m = 10)
Pinned by tests/golden/dlqr.npz, which oracle/make_golden.py produces by running those reference classes themselves.

The one deliberate difference: ``forward_predict`` (scipy ``solve_ivp`` RK45 over one control step in the reference) is
evaluated exactly with a matrix exponential; the two agree to ~1e-11 over 1/240 s (checked against the golden file).
State per drone: ``theta`` [(m+4), m] = [Ahat, Bhat]^T and ``P`` [(m+4), (m+4)].
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as la

TARGET_PREDICT, TARGET_XDOT = 0, 1
PROJECT_NONE, PROJECT_AFTER, PROJECT_LOOP = 0, 1, 2


def forward_predict(theta, e0, u, dt):
    """decentralized_lqr_omega.py:87-98: integrate e' = Ahat e + Bhat u over dt (exactly, via an augmented expm)."""
    m = theta.shape[1]
    A, B = theta[:m].T, theta[m:].T
    aug = np.zeros((m + 1, m + 1))
    aug[:m, :m] = A
    aug[:m, m] = B @ u
    return (la.expm(aug * dt) @ np.append(e0, 1.0))[:m]


def est_x_dot(m, x_tp1, phi, dt):
    """decentralized_lqr.py:185-198 (m = 12), decentralized_yolqr_crazyflie.py:228-243 (m = 10)."""
    xd = np.zeros(m)
    if m == 12:
        xd[0:3] = x_tp1[3:6]
        xd[3:6] = (x_tp1[3:6] - phi[3:6]) / dt
        xd[6:9] = (x_tp1[6:9] - phi[6:9]) / dt
        xd[9:] = x_tp1[6:9]
    elif m == 10:
        xd[0:3] = phi[m + 1:]
        xd[3] = phi[m]
        xd[4:7] = (x_tp1[4:7] - phi[4:7]) / dt
        xd[7:] = x_tp1[4:7]
    else:
        raise ValueError("the reference defines est_x_dot for m = 10 and m = 12 only")
    return xd


def project_codes(m):
    """project_theta's masks as one code per entry of theta [(m+4), m]: 0 -> zero, 1 -> keep, 2 -> force one.
    decentralized_lqr.py:70-87,230-240 (m = 12); decentralized_yolqr_crazyflie.py:164-186,245-257 (m = 10)."""
    A = np.zeros((m, m), int)
    B = np.zeros((m, 4), int)
    if m == 12:
        A[(6, 7), (1, 0)] = 1
        A[(0, 1, 2), (3, 4, 5)] = 2   # A_keep_2 and A_keep_3 are both forced to one (:237-238)
        A[(9, 10, 11), (6, 7, 8)] = 2
        B[3:6, 1:] = 1
        B[8, 0] = 1
    elif m == 10:
        A[(4, 5), (1, 0)] = 1
        A[(7, 8, 9), (4, 5, 6)] = 2   # only A_keep_2 is forced (:254-255)
        A[6, 3] = 1
        B[(0, 1, 2), (1, 2, 3)] = 1
        B[3, 0] = 1
    else:
        raise ValueError("the reference defines project_theta for m = 10 and m = 12 only")
    return np.vstack([A.T, B.T])  # theta = [A, B]^T


def project_theta(theta, codes):
    out = np.where(codes == 0, 0.0, theta)
    return np.where(codes == 2, 1.0, out)


def rls_update(theta, P, phi, x_tp1, dt, target=TARGET_PREDICT, predict_from_xtp1=False, normalize_gain=True,
               project=PROJECT_NONE, codes=None, first_of_env=True):
    """One update of one drone -> (theta_new, P_new, residual).  ``normalize_gain=False`` is theta_update2's information
    form with ``P`` holding V^-1 (V_new = V + phi phi' <=> P_new = P - P phi phi' P / (1 + phi' P phi))."""
    m = theta.shape[1]
    phi, x_tp1 = np.asarray(phi, float), np.asarray(x_tp1, float)
    if project == PROJECT_LOOP and not first_of_env:
        theta = project_theta(theta, codes)
    w = P @ phi
    s = 1.0 + phi @ w
    if target == TARGET_XDOT:
        r = est_x_dot(m, x_tp1, phi, dt) - theta.T @ phi
    else:
        e0 = x_tp1 if predict_from_xtp1 else phi[:m]
        r = x_tp1 - forward_predict(theta, e0, phi[m:], dt)
    L = w / s if normalize_gain else w
    theta_new = theta + np.outer(L, r)
    if project != PROJECT_NONE:
        theta_new = project_theta(theta_new, codes)
    P_new = P - np.outer(w / s, phi @ P)
    return theta_new, P_new, r
