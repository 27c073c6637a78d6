"""Exact strictly-convex QP solver -- TEST INFRASTRUCTURE ONLY.

Stands in for ``cvxopt.solvers.qp`` (cvxopt 1.3.2, reference environment.yaml:52,
call site cbf/qptracker.py:106), which is not installed and cannot be fetched.
PARITY UNPINNED against cvxopt itself: because P is symmetric positive definite
the minimiser is unique, so any exact solver is a valid oracle; every solution
is validated by the KKT certificate ``kkt_residuals`` (tests require <= 1e-9).

Algorithm: Goldfarb & Idnani (1983) dual active-set method,
    minimise 1/2 x'Px + q'x   subject to   G x <= h,
dense fp64, refactorising the (small) active-set Gram matrix every iteration.
"""
from __future__ import annotations

import numpy as np

STATUS_OPTIMAL, STATUS_INFEASIBLE, STATUS_ITER_CAP = 0, 1, 2


def solve_qp(P, q, G, h, max_iter=None, tol=1e-11):
    """-> (x, lam, status, iterations).  ``lam`` are multipliers of G x <= h."""
    P = np.asarray(P, float)
    q = np.asarray(q, float).reshape(-1)
    G = np.asarray(G, float).reshape(-1, q.size)
    h = np.asarray(h, float).reshape(-1)
    n, m = q.size, h.size
    max_iter = 10 * (m + n) + 50 if max_iter is None else max_iter
    Pinv = np.linalg.inv(P)
    x = -Pinv @ q
    lam = np.zeros(m)
    active: list[int] = []
    scale = 1.0 + np.abs(h) + np.linalg.norm(G, axis=1) * (1.0 + np.linalg.norm(x))
    it = 0
    while True:
        slack = h - G @ x
        viol = slack / scale
        viol[active] = 0.0
        p = int(np.argmin(viol))
        if viol[p] >= -tol:
            # certify: rows held active must still be satisfied (guards numerical breakdown
            # on nearly dependent active sets); otherwise report failure, never a wrong optimum
            if active and np.min((slack / scale)[active]) < -1e-7:
                return x, lam, STATUS_ITER_CAP, it
            return x, lam, STATUS_OPTIMAL, it
        n_p = G[p]
        lam_p = 0.0
        while True:
            it += 1
            if it > max_iter:
                return x, lam, STATUS_ITER_CAP, it
            if active:
                Na = G[active]                               # q x n
                PiNt = Pinv @ Na.T
                gram = Na @ PiNt
                r = np.linalg.solve(gram, Na @ (Pinv @ n_p))
                z = Pinv @ n_p - PiNt @ r
            else:
                r = np.zeros(0)
                z = Pinv @ n_p
            zn = float(n_p @ z)
            # partial step: keep the active multipliers non-negative
            t1, k_drop = np.inf, -1
            for idx, rj in enumerate(r):
                if rj > tol:
                    cand = lam[active[idx]] / rj
                    if cand < t1:
                        t1, k_drop = cand, idx
            s_p = h[p] - n_p @ x
            t2 = -s_p / zn if zn > 1e-9 * (n_p @ n_p) else np.inf
            t = min(t1, t2)
            if not np.isfinite(t):
                return x, lam, STATUS_INFEASIBLE, it
            if np.isfinite(t2):
                x = x - t * z                                # G x <= h  => move against the normal
            for idx, rj in enumerate(r):
                lam[active[idx]] -= t * rj
            lam_p += t
            if t == t2:
                lam[p] = lam_p
                active.append(p)
                break
            dropped = active.pop(k_drop)
            lam[dropped] = 0.0
    # unreachable


def kkt_residuals(P, q, G, h, x, lam):
    """(stationarity, primal infeasibility, dual infeasibility, complementarity), all inf-norms."""
    P, G = np.asarray(P, float), np.asarray(G, float)
    stat = np.max(np.abs(P @ x + q + G.T @ lam)) if x.size else 0.0
    slack = h - G @ x
    prim = max(0.0, float(np.max(-slack))) if h.size else 0.0
    dual = max(0.0, float(np.max(-lam))) if h.size else 0.0
    comp = float(np.max(np.abs(lam * slack))) if h.size else 0.0
    return stat, prim, dual, comp
