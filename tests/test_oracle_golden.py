"""CPU: the numpy oracle reproduces every fixture generated from the REFERENCE's own modules
(tests/golden/*.npz, made by oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import cbf as ocbf
from oracle import controllers as octl
from oracle import conversions as cv
from oracle import models as omd
from oracle import trajectories as otj
from oracle.constants import drone_params
from oracle.qp import kkt_residuals, solve_qp
from helpers import rel_err

TOL = 1e-9


def unpack(r):
    return r[0:3], r[3:6], r[6:9], r[9], r[10]


def pack(s):
    p, v, a, y, w = s
    return np.hstack([np.asarray(p, float), np.asarray(v, float), np.asarray(a, float), y, w])


def test_trajectories(golden):
    g = golden["trajectories"]
    Rz = np.array([[np.cos(.7), -np.sin(.7), 0], [np.sin(.7), np.cos(.7), 0], [0, 0, 1]])
    gens = {
        "circle": otj.Circle(r=1, v=.5, center=(0, 0, 1), yaw_rate=.1),
        "circle2": otj.Circle(r=0.7, v=1.3, center=(0.2, -0.4, 0.8), yaw_rate=-0.35),
        "lemniscate": otj.Lemniscate(center=(0, 0, .5), omega=1.5, yaw_rate=.1, phase_shift=-np.pi / 4),
        "lemniscate2": otj.Lemniscate(a=1.4, center=(0.3, 0.1, 1.5), omega=0.5, yaw_rate=0, phase_shift=2.1),
        "wait": otj.Wait((0.5, -0.2, 1.0), 3.0, yaw=0.4),
        "line": otj.Line((0, 0, 0.5), (2.0, 1.0, 1.5), speed=0.8),
        "line_short": otj.Line((0, 0, 0.5), (0.2, 0.1, 0.6), speed=1.5),
        "line_s0": otj.Line((1.0, 0, 0.5), (-2.0, 1.0, 0.5), speed=1.0, s0=0.3, sf=0.2),
        "rotate": otj.Rotate(otj.Lemniscate(center=(0, 0, .5), omega=0.8), Rz, (0.1, 0.2, 0.5)),
    }
    for name, gen in gens.items():
        got = np.array([pack(gen(float(t))) for t in g["t"]])
        assert np.allclose(got, g[name], rtol=TOL, atol=1e-12), name
    comp = otj.Compound([otj.Wait((0, 0, 0.5), 1.0), otj.Line((0, 0, 0.5), (1.5, 0.5, 1.0), speed=0.7),
                         otj.Circle(r=0.5, v=0.4, center=(1.0, 0.5, 1.0), duration=4.0), otj.Wait((1.5, 0.5, 1.0), 2.0, yaw=0.0)])
    got = np.array([pack(comp(float(t))) for t in g["compound_t"]])
    assert np.allclose(got, g["compound"], rtol=TOL, atol=1e-12)


def test_circle_yaw_wrap_quirk():
    # quirk B18: yaw in [pi, 3 pi); SURVEY anchor 6.358185307179586 at t = 0.75, rate 0.1
    assert abs(otj.Circle(r=1, v=.5, center=(0, 0, 1), yaw_rate=.1)(0.75)[3] - 6.358185307179586) < 1e-14


def test_line_requires_speed():
    with pytest.raises((AssertionError, TypeError)):
        otj.Line((0, 0, 0), (1, 0, 0), speed=None, duration=2.0)


@pytest.mark.parametrize("model", ["cf2p", "cf2x"])
def test_controllers(golden, model):
    g = golden["controllers"]
    env = drone_params(model, 240, 240)
    obs, refs = g[f"{model}_obs"], g[f"{model}_ref"]
    geo = octl.Geometric(env)
    acts = []
    for o, r in zip(obs, refs):
        geo.set_desired_trajectory(0, *unpack(r))
        acts.append(geo.compute(o.copy()))
    assert rel_err(acts, g[f"{model}_geometric_action"]) < TOL
    for kind in ("torque12", "omega9", "yank10"):
        low = None if kind == "torque12" else (octl.ThrustOmegaPid(env) if kind == "omega9" else octl.YankOmegaPid(env))
        c = octl.Lqr(env, kind, low)
        assert np.allclose(c.K, g[f"{model}_{kind}_K"], rtol=1e-9, atol=1e-9 * np.abs(c.K).max())
        acts, us, us_skip = [], [], []
        for o, r in zip(obs, refs):
            c.set_desired_trajectory(0, *unpack(r))
            if low is not None:
                low.reset()
                us_skip.append(c.compute(o.copy(), skip_low_level=True)[1].copy())
            a, u = c.compute(o.copy())
            acts.append(a)
            us.append(u.copy())
        assert rel_err(acts, g[f"{model}_{kind}_action"]) < 1e-8, kind
        assert np.allclose(us, g[f"{model}_{kind}_u"], rtol=1e-8, atol=1e-9), kind
        if us_skip:
            assert np.allclose(us_skip, g[f"{model}_{kind}_u_skip"], rtol=1e-8, atol=1e-9), kind
    pid = octl.ThrustOmegaPid(env)
    for k in range(obs.shape[0]):
        pid.reset()
        a1 = pid.compute(g[f"{model}_pid_u"][k].copy(), env.CTRL_TIMESTEP, g[f"{model}_pid_w"][0, k])
        a2 = pid.compute(g[f"{model}_pid_u"][k].copy(), env.CTRL_TIMESTEP, g[f"{model}_pid_w"][1, k])
        got = np.hstack([a1, a2, pid.last_omega, pid.integral])
        assert np.allclose(got, g[f"{model}_pid_out"][k], rtol=1e-12, atol=1e-12)


def test_models_and_conversions(golden):
    g = golden["models"]
    env = drone_params("cf2p", 240, 240)
    obs = g["obs"]
    assert np.allclose([omd.xdot_linear12_from_obs(env, o) for o in obs], g["xdot_linear12"], rtol=TOL, atol=1e-12)
    assert np.allclose([omd.xdot_nonlinear_from_obs(env, o) for o in obs], g["xdot_nonlinear"], rtol=TOL, atol=1e-12)
    assert np.allclose(g["J_dynamics"], [1.05, 1.05, 2.05])  # finding 6: load_env_params leaves J at Hummingbird values
    for kind in ("torque12", "omega9", "yank10"):
        A, B, Ah, Bh = omd.linear_model_matrices(env, kind)
        for got, key in ((A, "A"), (B, "B"), (Ah, "Ahat"), (Bh, "Bhat")):
            assert np.array_equal(got, g[f"{kind}_{key}"]), (kind, key)
    assert np.allclose([cv.obs_to_lin_model(o, 9) for o in obs], g["lin9"], rtol=0, atol=0)
    assert np.allclose([cv.obs_to_lin_model(o, 10, env) for o in obs], g["lin10"], rtol=1e-15)
    assert np.allclose([cv.obs_to_geo_model(o) for o in obs], g["geo18"], rtol=TOL, atol=1e-15)
    assert np.allclose([cv.action_to_input(env, o[16:]) for o in obs], g["action_to_input"], rtol=TOL, atol=1e-18)
    assert np.allclose([cv.input_to_action(env, u.copy()) for u in g["input_to_action_in"]], g["input_to_action"], rtol=TOL)


CBF_CASES = [("o2_n2_obs1", 2, 2), ("o2_n8_obs1", 2, 8), ("o3_n7_obs0", 3, 7), ("o3_n8_obs1", 3, 8), ("o3_n4_obs3", 3, 4)]


def cbf_params(env, order):
    return ocbf.CbfParams(env, order, 1.0 if order == 2 else 2.0, 0.1 if order == 2 else 0.125,
                          (-2.2, -2.4) if order == 2 else (-3.0, -3.6, -5.6))


@pytest.mark.parametrize("name,order,N", CBF_CASES)
def test_cbf_rows_and_qp(golden, name, order, N):
    g = golden["cbf_rows"]
    env = drone_params("cf2p", 240, 240)
    prm = cbf_params(env, order)
    assert np.allclose(prm.K, g[f"{name}_Kcbf"].reshape(-1), rtol=1e-12)
    assert np.allclose(prm.umax, g[f"{name}_umax"], rtol=1e-12)
    obs, xdes, obst, unom = g[f"{name}_obs"], g[f"{name}_xdes"], g[f"{name}_obstacles"], g[f"{name}_unom"]
    n_active = 0
    for e in range(obs.shape[0]):
        x = np.array([cv.obs_to_lin_model(obs[e, i], prm.xdim, env) for i in range(N)])
        G, h = ocbf.build_ineq(prm, x, xdes[e], obst[:, :3] if len(obst) else None, list(obst[:, 3]) if len(obst) else None)
        Gr, hr = g[f"{name}_G"][e], g[f"{name}_h"][e]
        assert G.shape == Gr.shape
        assert np.max(np.abs(G - Gr)) <= 1e-9 * (1 + np.max(np.abs(Gr)))
        assert np.max(np.abs(h - hr) / (1 + np.abs(hr))) <= 1e-9
        # the QP answer stored in the fixture came from the reference's tracker + this oracle solver: re-derive and certify
        uhat = unom[e].reshape(-1)
        u, lam, status, _ = solve_qp(np.eye(4 * N), -uhat, Gr, hr)
        usafe_ref = g[f"{name}_usafe_oracle_qp"][e]
        if status == 0:
            assert max(kkt_residuals(np.eye(4 * N), -uhat, Gr, hr, u, lam)) < 1e-9
            assert np.allclose(u.reshape(N, 4), usafe_ref, rtol=1e-9, atol=1e-10)
            n_active += int(np.sum(lam > 0))
        else:
            assert np.allclose(usafe_ref, unom[e])  # nominal fallback (cbf/qptracker.py:30-34)
    assert n_active > 0  # fixtures must exercise the active-set path


def test_qp_anchor():
    """SURVEY App. C anchor: descending drone over the sphere obstacle."""
    env = drone_params("cf2p", 100, 100)
    prm = ocbf.CbfParams(env, 2, 1, 0.1, (-2.2, -2.4))

    def mk(p, v):
        return np.concatenate([p, [0, 0, 0, 1], [0, 0, 0], v, [0, 0, 0], [env.HOVER_RPM] * 4])

    obs = np.array([mk([0, 0, .75], [0, 0, -.8]), mk([1, 1, 1.5], [0, 0, 0])])
    xdes = np.array([np.hstack([0, 0, 0, [0, 0, 0], o[:3]]) for o in obs])
    unom = np.array([[-0.2, 0, 0, 0], [0, 0, 0, 0.]])
    u, st, _ = ocbf.safety_filter(prm, env, obs, xdes, unom, x_obs=[[0, 0, .5]], obs_r=[.1])
    assert st == 0 and abs(u[0, 0] - (-0.113260464)) < 1e-8 and np.allclose(u.reshape(-1)[1:], 0)


def test_qp_random_kkt():
    rng = np.random.default_rng(1)
    for _ in range(100):
        n, m = 24, 80
        G = rng.normal(size=(m, n)) * (rng.uniform(size=(m, n)) < 0.25)
        h, q = rng.uniform(0.0, 2, m), rng.normal(size=n) * 3
        x, lam, st, _ = solve_qp(np.eye(n), q, G, h)
        assert st == 0 and max(kkt_residuals(np.eye(n), q, G, h, x, lam)) < 1e-9


def test_qp_infeasible_detected():
    G = np.array([[1.0, 0.0], [-1.0, 0.0]])
    h = np.array([-1.0, -1.0])  # x <= -1 and x >= 1
    _, _, st, _ = solve_qp(np.eye(2), np.zeros(2), G, h)
    assert st == 1


def test_cylinder_row_is_the_zscale_limit():
    """Builder extension (parity unpinned: the reference has spheres only): a vertical-cylinder obstacle row equals the
    sphere row of the same radius in the limit zscale -> infinity, and differs from it at the reference's zscale."""
    from oracle import cbf as ocbf
    from oracle.aviary import OracleCtrlAviary
    from oracle.constants import DroneModel, Physics
    env = OracleCtrlAviary(DroneModel.CF2P, 2, physics=Physics.DYN)
    rng = np.random.default_rng(9)
    for order, xdim in ((2, 9), (3, 10)):
        x, xdes = rng.normal(0, 0.4, (2, xdim)), rng.normal(0, 0.4, (2, xdim))
        xo = [np.array([0.3, -0.2, 0.7])]
        prm = ocbf.CbfParams(env, order, 2.0, 0.125, (-2.2, -2.4) if order == 2 else (-3.0, -3.6, -5.6))
        Gc, hc = ocbf.build_ineq(prm, x, xdes, xo, [-0.1])
        Gs, hs = ocbf.build_ineq(prm, x, xdes, xo, [0.1])
        big = ocbf.CbfParams(env, order, 1e9, 0.125, (-2.2, -2.4) if order == 2 else (-3.0, -3.6, -5.6))
        Gl, hl = ocbf.build_ineq(big, x, xdes, xo, [0.1])
        assert np.allclose(Gc[-2:], Gl[-2:], rtol=1e-12, atol=1e-15) and np.allclose(hc[-2:], hl[-2:], rtol=1e-12, atol=1e-15)
        assert not np.allclose(hc[-2:], hs[-2:])
        assert np.array_equal(Gc[:-2], Gs[:-2]) and np.array_equal(hc[:-2], hs[:-2])  # only the obstacle rows change


def test_linear_roll_out_matches_reference_style_integration():
    """CompareModels.py:82-95: the oracle's interval-by-interval roll-out equals one solve_ivp call over the whole log
    with the reference's 'closest observation in the past' input lookup (tight tolerances), and stays within the
    reference's own default-tolerance integration error."""
    from scipy.integrate import solve_ivp
    from oracle import conversions as cv
    from oracle import models as om
    from oracle.aviary import OracleCtrlAviary
    from oracle.constants import DroneModel, Physics
    env = OracleCtrlAviary(DroneModel.CF2P, 1, physics=Physics.DYN)
    rng = np.random.default_rng(4)
    T, dt = 40, 1.0 / 240
    obs = np.zeros((T, 20))
    obs[:, 0:3] = rng.normal(0, 1, (T, 3)); obs[:, 6] = 1.0
    obs[0, 7:10] = rng.uniform(-0.2, 0.2, 3); obs[0, 10:16] = rng.normal(0, 0.3, 6)
    ts = dt * np.arange(T)
    obs[:, 16:20] = env.HOVER_RPM + 300 * np.sin(2 * np.pi * 3 * ts[:, None] + rng.uniform(0, 6, 4)[None, :])  # a smooth logged flight
    y = om.roll_out_linear_system(env, obs, ts)
    A, B, _, _ = om.linear_model_matrices(env, "torque12")

    def f(t, x):  # the reference's closure, CompareModels.py:85-92
        idx = int(np.argmin(np.abs(ts - t)))
        if ts[idx] > t:
            idx -= 1
        xe = np.zeros(12); xe[9:] = x[9:]
        return A @ (x - xe) + B @ (cv.action_to_input(env, obs[idx][16:20]) - np.array([env.M * env.G, 0, 0, 0]))
    ref_default = solve_ivp(f, [0, ts[-1]], cv.obs_to_lin_model(obs[0], 12), t_eval=ts).y.T
    assert np.max(np.abs(y - ref_default)) < 2e-2 * (1 + np.max(np.abs(y)))   # the reference's own rtol = 1e-3 across input jumps
    tight = solve_ivp(f, [0, ts[-1]], cv.obs_to_lin_model(obs[0], 12), t_eval=ts, rtol=1e-10, atol=1e-12, max_step=dt).y.T
    assert np.max(np.abs(y - tight)) < 1e-7 * (1 + np.max(np.abs(y)))


def test_sysid_oracle_vs_reference_classes(golden):
    """oracle/sysid.py against theta / P produced by the reference's own Decentralized* classes over four successive
    updates of three robots (exact matrix exponential vs the reference's RK45 forward_predict included)."""
    from dlqr_cases import CASES
    from oracle import sysid
    g = golden["dlqr"]
    dt = float(g["dt"])
    for tag, (m, target, from_x1, normalize, project, _cls, _method, _kw) in CASES.items():
        info = not normalize
        codes = sysid.project_codes(m) if project else None
        th, P = g[tag + "_theta0"].copy(), g[tag + "_P0"].copy()
        if info:
            P = np.array([np.linalg.inv(p) for p in P])
        T, N = g[tag + "_phi"].shape[:2]
        for t in range(T):
            for i in range(N):
                th[i], P[i], _ = sysid.rls_update(th[i], P[i], g[tag + "_phi"][t, i], g[tag + "_x1"][t, i], dt, target, from_x1, normalize,
                                                  project, codes, first_of_env=(i == 0))
            Pc = np.array([np.linalg.inv(p) for p in P]) if info else P
            assert np.abs(th - g[tag + "_theta"][t]).max() <= 1e-10 * max(1.0, np.abs(g[tag + "_theta"][t]).max()), (tag, t)
            assert np.abs(Pc - g[tag + "_P"][t]).max() <= 1e-10 * np.abs(g[tag + "_P"][t]).max(), (tag, t)


def test_golden_fixtures_reproduce_from_the_reference(golden, tmp_path):
    """Where the reference tree is present (the build container; never the GPU box), regenerate every fixture from the
    reference's own modules (oracle/make_golden.py) and compare with the committed files: tests/golden/ is what the
    reference computes today, not a stale copy."""
    import os
    from oracle import make_golden, ref_import
    if not ref_import.reference_available():
        pytest.skip("reference tree not present")
    saved = make_golden.OUT
    make_golden.OUT = str(tmp_path)
    import warnings
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # the reference's own numpy deprecations (cbf/cbf.py:395)
            ref = ref_import.load()
            for fn in (make_golden.golden_trajectories, make_golden.golden_controllers, make_golden.golden_models, make_golden.golden_cbf,
                       make_golden.golden_dlqr):
                fn(ref)
    finally:
        make_golden.OUT = saved
    fresh = {n[:-4]: np.load(os.path.join(str(tmp_path), n)) for n in os.listdir(str(tmp_path)) if n.endswith(".npz")}
    assert sorted(fresh) == sorted(golden)
    for name, g in golden.items():
        assert sorted(g.files) == sorted(fresh[name].files), name
        for k in g.files:
            a, b = fresh[name][k], g[k]
            assert a.shape == b.shape, (name, k)
            if a.size and a.dtype.kind == "f":
                assert np.abs(a - b).max() <= 1e-12 * max(1.0, float(np.abs(b).max())), (name, k)
            else:
                assert np.array_equal(a, b), (name, k)


def test_reference_loop_equals_port_closed_loop():
    """oracle/ref_pipeline.ReferenceLoop (the reference's own Lemniscate / LQRYankOmegaController / YankOmegaController / DroneCBF /
    DroneQPTracker objects in the loop of simulations/CBFTestOrd3.py:305-360) and the numpy port oracle/pipeline.run_cbf, continued
    from the same state in steady flight, produce the same observations: the port IS the reference's closed loop (both around the
    oracle env step and QP solver).  Runs wherever the reference packages are importable (/root/reference or oracle/_ref)."""
    import copy
    import bench
    from oracle import pipeline as opl, ref_pipeline as rpl
    if rpl.reference_root() is None:
        pytest.skip("reference packages not available (python -m oracle.build_ref)")
    env, trajs = bench._oracle_env(0)
    ctrls = opl.make_controllers(env, "yank10")
    opl.run_cbf(env, trajs, 3, 240, obstacles=bench.SWARM_OBSTACLES, ctrls=ctrls, t0=0.0, log=False)
    loop = rpl.ReferenceLoop(copy.deepcopy(env), 3, bench._lem_specs(), bench.SWARM_OBSTACLES)
    loop.adopt_inner_loop_state(ctrls)
    loop.t = 240 * env.CTRL_TIMESTEP
    o_ref = loop.run(30)
    _, o_port, info = opl.run_cbf(env, trajs, 3, 30, obstacles=bench.SWARM_OBSTACLES, ctrls=ctrls, t0=240 * env.CTRL_TIMESTEP, log=False)
    assert info["solves"] > 0 and loop.qp_fallbacks == 0
    assert np.max(np.abs(o_ref - o_port)) < 1e-8
