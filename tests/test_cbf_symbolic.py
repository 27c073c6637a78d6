"""CPU: the closed-form CBF rows (oracle/cbf.py:pair_row, which the CUDA rows are tested against) versus the symbolic derivation
the reference itself carries -- cbf/symb_lie_deriv.ipynb, cells 1, 8 and 17: h(e) = |E_xy e|^4 + (e_z / c)^4 - Ds^4 on the stacked
state of two agents, x' = A xhat + B u with the hover-linearised yank-omega model, h' = dh/dx . x', h'' = dh'/dx . x',
h''' = dh''/dx . x'.  SURVEY.md section 4 calls that notebook the only correctness asset in the reference; here sympy redoes its
automatic differentiation and the results are compared at random points.

The barrier value, h', L_f^r h and L_g L_f^(r-1) h must agree.  The h'' that enters the right-hand side does NOT: the reference's
custom_hdots reads hard-coded state indices (quirk B12, cbf/cbf.py:135-178), and parity is with the reference, so the oracle
reproduces that expression (pinned by tests/golden/cbf_rows.npz); this file asserts both facts."""
import numpy as np
import pytest

sp = pytest.importorskip("sympy")

from oracle import cbf as ocbf
from oracle.aviary import OracleCtrlAviary
from oracle.constants import DroneModel, Physics


def symbolic_lie(order):
    """-> (syms, h, hdots [h', h'', (h''')] as functions of x, xdes, u) for two agents; state dim 9 (order 2) or 10 (order 3)."""
    n = 9 if order == 2 else 10
    c, Ds, g, m = sp.symbols("c Ds g m", positive=True)
    x = sp.Matrix(sp.symbols(f"x:{2 * n}", real=True))
    xd = sp.Matrix(sp.symbols(f"xd:{2 * n}", real=True))
    u = sp.Matrix(sp.symbols("u:8", real=True))
    A1, B1 = sp.zeros(n, n), sp.zeros(n, 4)
    if order == 3:      # cell 8 == model/linear_yank_omega.py:45-57: x = [rpy, F, v, p], u = [yank, w]
        A1[7:, 4:7] = sp.eye(3); A1[4, 1] = g; A1[5, 0] = -g; A1[6, 3] = 1 / m
        B1[:3, 1:] = sp.eye(3); B1[3, 0] = 1
    else:               # model/linear_omega.py:46-61: x = [rpy, v, p], u = [f, w]
        A1[6:, 3:6] = sp.eye(3); A1[3, 1] = g; A1[4, 0] = -g
        B1[:3, 1:] = sp.eye(3); B1[5, 0] = 1 / m
    A, B = sp.diag(A1, A1), sp.diag(B1, B1)
    e = x[n - 3:n, 0] - x[2 * n - 3:2 * n, 0]
    h = (e[0] ** 2 + e[1] ** 2) ** 2 + (e[2] / c) ** 4 - Ds ** 4
    xdot = A * (x - xd) + B * u
    hs = [sp.Matrix([h])]
    for _ in range(order):
        hs.append(hs[-1].jacobian(x) * xdot)
    return dict(n=n, c=c, Ds=Ds, g=g, m=m, x=x, xd=xd, u=u), hs


@pytest.mark.parametrize("order", [2, 3])
def test_rows_match_the_symbolic_lie_derivatives(order):
    S, hs = symbolic_lie(order)
    n, u = S["n"], S["u"]
    args = [S["c"], S["Ds"], S["g"], S["m"], *S["x"], *S["xd"], *S["u"]]
    top = hs[order][0]                                   # h^(r) = L_f^r h + L_g L_f^(r-1) h . u
    Lg = sp.Matrix([top]).jacobian(u)
    Lf = top.subs({ui: 0 for ui in u})
    f_h, f_h1 = sp.lambdify(args, hs[0][0]), sp.lambdify(args, hs[1][0].subs({ui: 0 for ui in u}))
    f_h2 = sp.lambdify(args, hs[2][0].subs({ui: 0 for ui in u})) if order == 3 else None
    f_Lf, f_Lg = sp.lambdify(args, Lf), sp.lambdify(args, Lg)
    assert all(sp.simplify(hs[k][0].diff(ui)) == 0 for k in range(order) for ui in u)   # relative degree = order

    env = OracleCtrlAviary(DroneModel.CF2P, 2, physics=Physics.DYN)
    rng = np.random.default_rng(order)
    quirk_seen = False
    for trial in range(40):
        zs, rs = rng.uniform(0.8, 2.5), rng.uniform(0.05, 0.3)
        prm = ocbf.CbfParams(env, order, zs, rs, (-2.2, -2.4) if order == 2 else (-3.0, -3.6, -5.6))
        xi, xj, xid, xjd = (rng.normal(0, 0.6, n) for _ in range(4))
        Ds = 2 * rs
        vals = [zs, Ds, prm.g, prm.m, *xi, *xj, *xid, *xjd, *np.zeros(8)]
        # components of the oracle row through the gain vector: rhs = K . [h, h', (h'')] + L_f^r h
        comp = []
        for K in np.vstack([np.zeros(order), np.eye(order)]):
            prm.K = K
            a, rhs = ocbf.pair_row(prm, xi, xj, xid, xjd, Ds)
            comp.append(rhs)
        Lf_o, h_o, h1_o = comp[0], comp[1] - comp[0], comp[2] - comp[0]
        scale = lambda v: 1.0 + abs(v)
        assert abs(h_o - f_h(*vals)) < 1e-10 * scale(f_h(*vals))
        assert abs(h1_o - f_h1(*vals)) < 1e-10 * scale(f_h1(*vals))
        assert abs(Lf_o - f_Lf(*vals)) < 1e-9 * scale(f_Lf(*vals)), (trial, Lf_o, f_Lf(*vals))
        Lg_s = np.asarray(f_Lg(*vals), float).reshape(-1)
        # row coefficients: G = -a on agent i's block, +a on agent j's, and the row is  -h^(r) <= ... i.e. G = -L_g L_f^(r-1) h
        assert np.allclose(a, Lg_s[0:3], rtol=1e-10, atol=1e-12), (trial, a, Lg_s)
        assert np.allclose(-a, Lg_s[4:7], rtol=1e-10, atol=1e-12) and abs(Lg_s[3]) + abs(Lg_s[7]) == 0
        if order == 3:
            h2_o = comp[3] - comp[0]
            if abs(h2_o - f_h2(*vals)) > 1e-6 * scale(f_h2(*vals)):
                quirk_seen = True
    if order == 3:
        assert quirk_seen   # the reference's h'' (indices 6:9 of the 10-dim state) is not the symbolic one: reproduced on purpose
