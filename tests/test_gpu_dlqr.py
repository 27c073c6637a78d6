"""GPU parity of the decentralised-LQR path: mds_rls_update / mds_error_state / mds_dlqr_ctrl through the
host mirror multidronesim_b200.control.dlqr against (a) theta, P, K, action, u produced by the REFERENCE's own classes
(tests/golden/dlqr.npz) and (b) oracle/sysid.py on a larger random batch.  fp64 1e-9, fp32 2e-4 of the matrix scale."""
import numpy as np
import pytest
import torch

from dlqr_cases import CASES
from helpers import rel_err, scaled_err

pytestmark = pytest.mark.gpu
DTYPES = [torch.float64, torch.float32]
TOL = {torch.float64: 1e-9, torch.float32: 2e-4}


def make_env(E, N, dtype):
    import multidronesim_b200 as mds
    return mds.BatchedCtrlAviary(drone_model=mds.DroneModel("cf2p"), num_drones=N, num_envs=E, dtype=dtype)


def make_ctrl(env, cls_name):
    import multidronesim_b200 as mds
    from multidronesim_b200.control import dlqr
    model = {"DecentralizedLQROmega": mds.model.LinearizedOmegaModel, "DecentralizedLQR": mds.model.LinearizedModel}.get(cls_name, mds.model.LinearizedYankOmegaModel)
    models = [model(env) for _ in range(env.NUM_DRONES)]
    if cls_name == "DecentralizedYOLQRCrazyflie":
        return dlqr.DecentralizedYOLQRCrazyflie(env, models, np.eye(10), np.eye(4))
    return getattr(dlqr, cls_name)(env, models)


def dev(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda", dtype=dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("tag", list(CASES))
def test_rls_vs_reference_golden(golden, tag, dtype, lib_built):
    g = golden["dlqr"]
    m, _target, _fx, normalize, _proj, cls_name, method, kw = CASES[tag]
    T, N = g[tag + "_phi"].shape[:2]
    E = 5  # the same three robots replicated in five environments: every environment must reproduce the reference
    env = make_env(E, N, dtype)
    c = make_ctrl(env, cls_name)
    th0, P0 = g[tag + "_theta0"], g[tag + "_P0"]
    assert np.abs(c.theta.cpu().numpy()[:N] - th0).max() <= 1e-6 * np.abs(th0).max()  # the constructor builds the same prior
    if normalize:
        assert np.abs(c.P.cpu().numpy()[:N] - P0).max() <= 1e-6 * np.abs(P0).max()
    c.set_theta(dev(np.tile(th0, (E, 1, 1)), dtype))
    c.set_P(dev(np.tile(P0 if normalize else np.array([np.linalg.inv(p) for p in P0]), (E, 1, 1)), dtype))
    for t in range(T):
        getattr(c, method)(dev(np.tile(g[tag + "_phi"][t], (E, 1)), dtype), dev(np.tile(g[tag + "_x1"][t], (E, 1)), dtype), **kw)
        th = c.theta.cpu().numpy().astype(np.float64).reshape(E, N, m + 4, m)
        P = c.P.cpu().numpy().astype(np.float64).reshape(E, N, m + 4, m + 4)
        if not normalize:
            P = np.linalg.inv(P)
        want_th, want_P = g[tag + "_theta"][t], g[tag + "_P"][t]
        for e in range(E):
            assert np.abs(th[e] - want_th).max() <= TOL[dtype] * max(1.0, np.abs(want_th).max()), (tag, t, e)
            assert np.abs(P[e] - want_P).max() <= TOL[dtype] * np.abs(want_P).max(), (tag, t, e)


@pytest.mark.parametrize("tag", list(CASES))
def test_rls_random_batch_vs_oracle(tag, lib_built):
    """1 021 drones (not a multiple of the block), random well-conditioned P and perturbed theta, two successive updates."""
    from oracle import sysid
    m, target, from_x1, normalize, project, cls_name, method, kw = CASES[tag]
    E, N = 1021 // 3 + 1, 3
    env = make_env(E, N, torch.float64)
    D = E * N
    c = make_ctrl(env, cls_name)
    rng = np.random.default_rng(5)
    th = c.theta.cpu().numpy().copy() * (1.0 + 0.05 * rng.normal(size=(D, m + 4, m)))
    Lr = rng.normal(0, 0.3, (D, m + 4, m + 4))
    P = np.eye(m + 4) + Lr @ Lr.transpose(0, 2, 1)
    c.set_theta(dev(th, torch.float64))
    c.set_P(dev(P, torch.float64))
    codes = sysid.project_codes(m) if project else None
    for step in range(2):
        phi = np.concatenate([rng.normal(0, 0.1, (D, m)), rng.normal(0, 0.05, (D, 4))], axis=1)
        x1 = phi[:, :m] + rng.normal(0, 0.01, (D, m))
        resid = getattr(c, method)(dev(phi, torch.float64), dev(x1, torch.float64), **kw).cpu().numpy()
        for d in range(D):
            th[d], P[d], r = sysid.rls_update(th[d], P[d], phi[d], x1[d], env.CTRL_TIMESTEP, target, from_x1, normalize, project, codes,
                                              first_of_env=(d % N == 0))
            assert np.abs(resid[d] - r).max() <= 1e-9 * max(1.0, np.abs(r).max()), (tag, step, d)
        assert np.abs(c.theta.cpu().numpy() - th).max() <= 1e-9 * np.abs(th).max(), (tag, step)
        assert np.abs(c.P.cpu().numpy() - P).max() <= 1e-9 * np.abs(P).max(), (tag, step)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("tag", ["omega9", "torque12"])
def test_dlqr_compute_vs_reference_golden(golden, tag, dtype, lib_built):
    """compute_controller (CARE on the learned models; the 12-dim variant couples robots 0 and 1 through Q) and compute(obs)."""
    g = golden["dlqr"]
    N = g["ctrl_obs"].shape[0]
    E = 4
    env = make_env(E, N, dtype)
    c = make_ctrl(env, "DecentralizedLQROmega" if tag == "omega9" else "DecentralizedLQR")
    m = c.m
    c.set_theta(dev(np.tile(g[f"ctrl_{tag}_theta"], (E, 1, 1)), dtype))
    K = c.compute_controller()  # omega9: device Riccati solve (fp32: from the fp32 copy of theta); torque12: coupled -> host scipy
    Kg = g[f"ctrl_{tag}_K"]
    if tag == "omega9":
        ktol = 1e-9 if dtype == torch.float64 else 2e-5
        for i in range(N):  # block-diagonal reference K: drone i keeps block (i, i); the other blocks are zero there
            Ki = c.K_matrix(i, env_idx=E - 1).double().cpu().numpy()
            assert np.abs(Ki - Kg[4 * i:4 * i + 4, m * i:m * i + m]).max() <= ktol * np.abs(Kg).max()
        assert int(c.care_status.sum()) == 0
        off = Kg.copy()
        for i in range(N):
            off[4 * i:4 * i + 4, m * i:m * i + m] = 0
        assert np.abs(off).max() <= 1e-9 * np.abs(Kg).max()
    else:
        assert np.allclose(K[0], Kg, rtol=1e-6, atol=1e-8 * np.abs(Kg).max())
        assert np.abs(Kg[0:4, m:2 * m]).max() > 1e-5 * np.abs(Kg).max()  # the coupling really is there (2e-4 of the largest gain)
    ref = dev(np.tile(g["ctrl_ref"], (E, 1)), dtype)
    obs = dev(np.tile(g["ctrl_obs"], (E, 1)).reshape(E, N, 20), dtype)
    c.set_reference(ref)
    a, u = c.compute(obs)
    tol = 10 * (1e-9 if dtype == torch.float64 else 1e-5)
    a, u = a.cpu().numpy().astype(np.float64), u.cpu().numpy().astype(np.float64)
    ug = g[f"ctrl_{tag}_u"].reshape(N, 4).copy()
    if tag == "torque12":
        ug[:, 0] += env.M * env.G  # the 12-dim reference returns the flat u without the hover offset (decentralized_lqr.py:336-342)
    for e in range(E):
        assert rel_err(a[e], g[f"ctrl_{tag}_action"]) < tol, (tag, e)
        for col in range(4):  # thrust O(0.3), rates O(1), torques O(1e-3): compare column-wise against the column scale
            scale = max(np.abs(ug[:, col]).max(), 1e-3)
            assert np.abs(u[e][:, col] - ug[:, col]).max() <= tol * scale, (tag, e, col)


@pytest.mark.parametrize("dtype", DTYPES)
def test_error_state_vs_oracle(golden, dtype, lib_built):
    """mds_error_state for the three parametrisations against oracle.controllers.Lqr.error_state on the reference-pinned inputs."""
    from oracle import controllers as oc
    from oracle.constants import drone_params
    g = golden["controllers"]
    obs, refs = g["cf2p_obs"], g["cf2p_ref"]
    n = obs.shape[0]
    env = make_env(n, 1, dtype)
    oenv = drone_params("cf2p", 240, 240)
    for cls_name, kind in (("DecentralizedLQR", "torque12"), ("DecentralizedLQROmega", "omega9"), ("DecentralizedLQRYankOmega", "yank10")):
        c = make_ctrl(env, cls_name)
        c.set_reference(dev(refs, dtype))
        got = c.error_state(dev(obs.reshape(n, 1, 20), dtype)).cpu().numpy().reshape(n, -1)
        o = oc.Lqr(oenv, kind, K=np.zeros((4, c.m)))
        for k in range(n):
            o.set_desired_trajectory(0, refs[k, 0:3], refs[k, 3:6], refs[k, 6:9], refs[k, 9], refs[k, 10])
            want = o.error_state(obs[k])
            assert np.abs(got[k] - want).max() <= (1e-9 if dtype == torch.float64 else 2e-5) * max(1.0, np.abs(want).max()), (kind, k)


def test_dlqr_argument_errors(lib_built):
    from multidronesim_b200 import _lib
    env = make_env(2, 2, torch.float64)
    c = make_ctrl(env, "DecentralizedLQROmega")
    with pytest.raises(_lib.MdsError):
        c.compute(torch.zeros(2, 2, 20, device="cuda", dtype=torch.float64))  # no gains yet
    with pytest.raises(_lib.MdsError):
        c.project_theta()  # the reference defines no projection for the 9-dim model


def test_learning_loop_vs_oracle(lib_built):
    """The warm-up learning loop of simulations/CBFTestOrd3.py:153-198 (random_warmup): random inputs -> inner loop -> env.step
    -> error states -> theta_update, 40 control steps, every stage on device (mds_error_state, mds_lowlevel,
    mds_physics_step, mds_rls_update) against the same loop over the oracle (aviary + controllers + sysid), fp64."""
    import multidronesim_b200 as mds
    from oracle import controllers as oc
    from oracle import conversions as cv
    from oracle import sysid
    from oracle.aviary import OracleCtrlAviary
    from oracle.constants import DroneModel as ODM, Physics as OPH
    E, N, steps, dtype = 3, 2, 40, torch.float64
    rng = np.random.default_rng(21)
    init = rng.uniform(-0.5, 0.5, (E, N, 3)) + np.array([0, 0, 1.0])
    env = mds.BatchedCtrlAviary(drone_model=mds.DroneModel.CF2P, num_drones=N, num_envs=E, dtype=dtype, initial_xyzs=init, physics=mds.Physics.DYN)
    dl = mds.control.DecentralizedLQROmega(env, [mds.model.LinearizedOmegaModel(env) for _ in range(N)])
    dl.set_desired_trajectory(None, dev(init.reshape(-1, 3), dtype), np.zeros(3), np.zeros(3), 0.0, 0.0)
    mg = env.M * env.G
    us = np.concatenate([rng.uniform(0.7 * mg, 1.5 * mg, (steps, E, N, 1)), rng.uniform(-0.2, 0.2, (steps, E, N, 3))], axis=-1)  # sigma1-like draws
    # oracle side, one reference-style loop per environment
    oenvs = [OracleCtrlAviary(ODM.CF2P, N, initial_xyzs=init[e], physics=OPH.DYN) for e in range(E)]
    octl = [[oc.Lqr(oenvs[e], "omega9", low_level=oc.ThrustOmegaPid(oenvs[e]), K=np.zeros((4, 9))) for _ in range(N)] for e in range(E)]
    th = dl.theta.cpu().numpy().copy().reshape(E, N, 13, 9)
    P = dl.P.cpu().numpy().copy().reshape(E, N, 13, 13)
    oobs = [o.reset()[0] for o in oenvs]
    obs = env.reset()[0]
    for i in range(steps):
        u = dev(us[i], dtype)
        e_t = dl.error_state(obs).clone()
        phis = torch.cat([e_t, u], dim=-1)
        action = dl.compute_low_level(u, obs)
        obs = env.step(action)[0]
        e_tp1 = dl.error_state(obs).clone()
        if i != 0:
            dl.theta_update(phis, e_tp1)
        for e in range(E):
            ophis, act = [], np.zeros((N, 4))
            for j in range(N):
                c = octl[e][j]
                c.set_desired_trajectory(j, init[e, j], np.zeros(3), np.zeros(3), 0.0, 0.0)
                ophis.append(np.hstack([c.error_state(oobs[e][j]), us[i, e, j]]))
                act[j] = c.compute_low_level(us[i, e, j].copy(), oobs[e][j])
            oobs[e] = oenvs[e].step(act)[0]
            if i != 0:
                for j in range(N):
                    th[e, j], P[e, j], _ = sysid.rls_update(th[e, j], P[e, j], ophis[j], octl[e][j].error_state(oobs[e][j]), env.CTRL_TIMESTEP,
                                                            sysid.TARGET_PREDICT, True, True)
    got_th = dl.theta.cpu().numpy().reshape(E, N, 13, 9)
    got_P = dl.P.cpu().numpy().reshape(E, N, 13, 13)
    assert np.abs(obs.cpu().numpy() - np.array(oobs))[..., :16].max() < 1e-8
    assert np.abs(got_th - th).max() <= 1e-8 * np.abs(th).max(), np.abs(got_th - th).max()
    assert np.abs(got_P - P).max() <= 1e-8 * np.abs(P).max()
    prior = np.hstack([mds.model.LinearizedOmegaModel(env).Ahat, mds.model.LinearizedOmegaModel(env).Bhat]).T
    assert np.abs(th - prior).max() > 1e-3  # the models really moved


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cls_name", ["DecentralizedLQROmega", "DecentralizedLQRYankOmega", "DecentralizedLQR", "DecentralizedYOLQRCrazyflie"])
def test_care_device_vs_scipy(cls_name, dtype, lib_built):
    """mds_care_gains (one warp per drone, matrix sign function) against scipy.linalg.solve_continuous_are -- the reference's
    compute_controller (decentralized_lqr_omega.py:185-204) -- on 150 DISTINCT learned models: the prior scaled entry-wise
    by 5 % noise, half of them with a dense perturbation on top (what RLS produces)."""
    E, N = 50, 3
    env = make_env(E, N, dtype)
    c = make_ctrl(env, cls_name)
    m, D = c.m, E * N
    rng = np.random.default_rng(9)
    th = c.theta.double().cpu().numpy() * (1.0 + 0.05 * rng.normal(size=(D, m + 4, m)))
    th[D // 2:, :m, :] += 0.05 * rng.normal(size=(D - D // 2, m, m))
    c.set_theta(dev(th, dtype))
    K_host = np.array(c.compute_controller(force_diagonal=True, solver="host"))          # [D, 4, m], scipy per drone
    c.K_planes.zero_()
    c.compute_controller(force_diagonal=True, solver="device")
    assert int(c.care_status.sum()) == 0
    K_dev = c.K_planes.double().cpu().numpy().T.reshape(D, 4, m)
    tol = 1e-9 if dtype == torch.float64 else 1e-5  # fp32: K is rounded to fp32 on store; theta is the same fp32 copy on both sides
    for d in range(D):
        assert np.abs(K_dev[d] - K_host[d]).max() <= tol * np.abs(K_host[d]).max(), (cls_name, d)
    # a model with no stabilising solution (A = 0, B = 0: the uncontrollable integrators) is reported, its gain left alone
    bad = th.copy()
    bad[0] = 0.0
    c.set_theta(dev(bad, dtype))
    c.K_planes.fill_(7.0)
    c.compute_controller(force_diagonal=True, solver="device")
    st = c.care_status.cpu().numpy()
    assert st[0] == 1 and st[1:].sum() == 0
    assert float(c.K_planes[:, 0].min()) == 7.0


@pytest.mark.parametrize("dtype", DTYPES)
def test_fedce_wrapper_vs_reference_golden(golden, dtype, lib_built):
    """multidronesim_b200.fedce.FederatedLearning against the reference's own FedCE/FederatedLearning.py driven with YOState
    lists: five update() calls (the first only stores), then calc_controller() and lqr_control()."""
    import multidronesim_b200 as mds
    from multidronesim_b200.fedce import FederatedLearning
    g = golden["dlqr"]
    T, N = g["fed_x"].shape[:2]
    E = 3
    env = make_env(E, N, dtype)
    fl = FederatedLearning(env, [mds.model.LinearizedYankOmegaModel(env) for _ in range(N)], np.eye(10), np.eye(4), num_drones=N)
    tol = TOL[dtype]
    for t in range(T):
        rep = lambda a: dev(np.tile(a[None], (E, 1, 1)), dtype)
        out = fl.update(rep(g["fed_x"][t]), rep(g["fed_xdes"][t]), rep(g["fed_u"][t]))
        assert (out[0] is None) == (t == 0)
        th = fl.dLQR.theta.double().cpu().numpy().reshape(E, N, 14, 10)
        P = fl.dLQR.P.double().cpu().numpy().reshape(E, N, 14, 14)
        for e in range(E):
            assert np.abs(th[e] - g["fed_theta"][t]).max() <= tol * max(1.0, np.abs(g["fed_theta"][t]).max()), (t, e)
            assert np.abs(P[e] - g["fed_P"][t]).max() <= tol * np.abs(g["fed_P"][t]).max(), (t, e)
    fl.calc_controller()
    Kg = g["fed_K"]
    ktol = 1e-9 if dtype == torch.float64 else 1e-3  # fp32: the CARE of an fp32-rounded learned model
    for i in range(N):
        assert np.abs(fl.dLQR.K_matrix(i).double().cpu().numpy() - Kg[4 * i:4 * i + 4, 10 * i:10 * i + 10]).max() <= ktol * np.abs(Kg).max()
    u = fl.lqr_control(rep(g["fed_x"][-1]), rep(g["fed_xdes"][-1])).double().cpu().numpy()
    for e in range(E):
        assert np.abs(u[e] - g["fed_ctrl_u"]).max() <= (1e-8 if dtype == torch.float64 else 2e-3) * max(1.0, np.abs(g["fed_ctrl_u"]).max())
    sd = fl.make_desired_state(pos=[0.0, 0.0, 1.0])
    assert sd.shape == (E, N, 10) and abs(float(sd[0, 0, 3]) - env.M * env.G) < 1e-6 and float(sd[0, 0, 9]) == 1.0


@pytest.mark.parametrize("E,N", [(7, 5), (1, 1), (3, 32)])
def test_coupled_gain_law_odd_sizes(E, N, lib_built):
    """mds_dlqr_ctrl with robot-coupled gains at group sizes that do not fill a warp (N = 5: 125-thread blocks), a single drone
    and the largest group (N = 32): u_d = -sum_j K_dj e_j + [m g, 0, 0, 0] against numpy on the device's own error states."""
    dtype = torch.float64
    env = make_env(E, N, dtype)
    c = make_ctrl(env, "DecentralizedLQR")
    m, D = c.m, E * N
    rng = np.random.default_rng(3)
    Kfull = rng.normal(0, 1e-3, (E, 4 * N, m * N))                      # any gain will do: this tests the law, not the CARE
    c._coupled = True
    planes = Kfull.reshape(E, N, 4, N, m).transpose(3, 2, 4, 0, 1).reshape(N * 4 * m, D)
    c.K_planes = dev(planes, dtype)
    c.K = Kfull
    rpy = rng.uniform(-0.3, 0.3, (D, 3))
    from scipy.spatial.transform import Rotation
    obs = np.concatenate([rng.uniform(-1, 1, (D, 3)), Rotation.from_euler("xyz", rpy).as_quat(), rpy, rng.normal(0, 0.5, (D, 3)),
                          rng.normal(0, 0.5, (D, 3)), rng.uniform(12000, 17000, (D, 4))], axis=1).reshape(E, N, 20)
    ref = np.concatenate([rng.uniform(-1, 1, (D, 3)), rng.normal(0, 0.2, (D, 3)), np.zeros((D, 3)), rng.uniform(-1, 1, (D, 1)), rng.normal(0, 0.1, (D, 1))], axis=1)
    c.set_reference(dev(ref, dtype))
    e = c.error_state(dev(obs, dtype)).cpu().numpy().reshape(E, N * m)
    _, u = c.compute(dev(obs, dtype))
    want = -(Kfull @ e[:, :, None])[:, :, 0].reshape(E, N, 4)
    want[..., 0] += env.M * env.G
    want[..., 0] = np.maximum(want[..., 0], 0.0)                        # input_to_action clamps the thrust in place
    assert np.abs(u.cpu().numpy() - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def test_dlqr_as_cbf_nominal_controller_vs_oracle(lib_built):
    """The reference's ``--controller dlqr`` loop (simulations/CBFTest.py:302-358): DecentralizedLQROmega.compute(obs,
    skip_low_level=True) as the nominal controller of the order-2 CBF-QP, each drone with its own learned model and gain
    (device Riccati solve), through PerCallPipeline; 80 steps, fp64, against the oracle loop with the same per-drone gains."""
    import multidronesim_b200 as mds
    import multidronesim_b200.trajectories as T
    from oracle import controllers as oc
    from oracle import pipeline as opl
    from oracle import trajectories as otj
    from oracle.aviary import OracleCtrlAviary
    from oracle.constants import DroneModel as ODM, Physics as OPH
    E, N, steps, dtype = 2, 3, 80, torch.float64
    rng = np.random.default_rng(4)
    ph = np.pi / 2 - 0.3 + (2 * np.pi / (N + 0.25)) * np.arange(N)
    specs = [dict(a=1.0, center=np.array([0, 0, 0.5]), omega=0.5, yaw_rate=0.0, phase_shift=float(p)) for p in ph]
    obstacles = [[0.0, 0.0, 0.5, 0.1]]
    init = np.zeros((E, N, 3))
    for e in range(E):
        for j, sp in enumerate(specs):
            init[e, j] = otj.Lemniscate(**sp)(0.0)[0] + rng.normal(0, 0.02, 3) + np.array([0, 0, 0.04 * j])
    env = mds.BatchedCtrlAviary(drone_model=mds.DroneModel.CF2P, num_drones=N, physics=mds.Physics.DYN_GND_DRAG_DW, num_envs=E, dtype=dtype,
                                initial_xyzs=init)
    models = [mds.model.LinearizedOmegaModel(env) for _ in range(N)]
    dl = mds.control.DecentralizedLQROmega(env, models)
    th = dl.theta.cpu().numpy() * (1.0 + 0.03 * rng.normal(size=(E * N, 13, 9)))   # every drone its own "learned" model
    dl.set_theta(dev(th, dtype))
    dl.compute_controller()
    assert int(dl.care_status.sum()) == 0
    cbf = mds.cbf.DroneCBF(env, models, safety_radius=0.1, zscale=1.0, order=2, cbf_poles=np.array([-2.2, -2.4]))
    trk = mds.cbf.DroneQPTracker(cbf, order=2, num_robots=N, xdim=9, env=env)
    ts = mds.trajectories.TrajectorySet([T.Lemniscate(**sp) for sp in specs] * E, dtype=dtype)
    pipe = mds.PerCallPipeline(env, dl, trk, obstacles)
    obs = None
    for k in range(steps):
        obs = pipe.step(ts.eval(k * env.CTRL_TIMESTEP))
    got = obs.cpu().numpy()
    for e in range(E):
        o = OracleCtrlAviary(ODM.CF2P, N, initial_xyzs=init[e], physics=OPH.DYN_GND_DRAG_DW)
        ctrls = [oc.Lqr(o, "omega9", oc.ThrustOmegaPid(o), K=dl.K_matrix(j, env_idx=e).cpu().numpy()) for j in range(N)]
        _, want, info = opl.run_cbf(o, [otj.Lemniscate(**sp) for sp in specs], 2, steps, obstacles=obstacles, ctrls=ctrls)
        assert np.abs(got[e, :, 0:3] - want[:, 0:3]).max() < 1e-6, (e, info)


@pytest.mark.parametrize("cls_name,order", [("DecentralizedLQROmega", 2), ("DecentralizedLQROmega", None), ("DecentralizedLQR", None)])
def test_fused_rollout_with_per_drone_gains(cls_name, order, lib_built):
    """MdsRolloutCfg.lqr_gain_planes_dev: the K-steps-in-one-launch rollout (and the two-launch plan) with a gain per drone
    reproduce the per-call pipeline driven by the same decentralised controller (fp64; 12-dim: force_diagonal gains)."""
    import multidronesim_b200 as mds
    import multidronesim_b200.trajectories as T
    from oracle import trajectories as otj
    E, N, steps, dtype = 4, 3, 30, torch.float64
    rng = np.random.default_rng(8)
    ph = (2 * np.pi / (N + 0.25)) * np.arange(N)
    specs = [dict(a=1.0, center=np.array([0, 0, 0.5]), omega=0.5, yaw_rate=0.0, phase_shift=float(p)) for p in ph]
    obstacles = [[0.0, 0.0, 0.5, 0.1]] if order else None
    init = np.zeros((E, N, 3))
    for e in range(E):
        for j, sp in enumerate(specs):
            init[e, j] = otj.Lemniscate(**sp)(0.0)[0] + rng.normal(0, 0.02, 3) + np.array([0, 0, 0.04 * j])
    outs = []
    for mode in ("percall", "loop", "two"):
        env = mds.BatchedCtrlAviary(drone_model=mds.DroneModel.CF2P, num_drones=N, physics=mds.Physics.DYN_GND_DRAG_DW, num_envs=E, dtype=dtype,
                                    initial_xyzs=init)
        dl = make_ctrl(env, cls_name)
        th = dl.theta.cpu().numpy() * (1.0 + 0.03 * np.random.default_rng(1).normal(size=(E * N, dl.m + 4, dl.m)))
        dl.set_theta(dev(th, dtype))
        dl.compute_controller(force_diagonal=True)
        trk = None
        if order:
            cbf = mds.cbf.DroneCBF(env, [mds.model.LinearizedOmegaModel(env) for _ in range(N)], safety_radius=0.1, zscale=1.0, order=2,
                                   cbf_poles=np.array([-2.2, -2.4]))
            trk = mds.cbf.DroneQPTracker(cbf, order=2, num_robots=N, xdim=9, env=env)
        ts = mds.trajectories.TrajectorySet([T.Lemniscate(**sp) for sp in specs] * E, dtype=dtype)
        if mode == "percall":
            pipe = mds.PerCallPipeline(env, dl, trk, obstacles)
            for k in range(steps):
                obs = pipe.step(ts.eval(k * env.CTRL_TIMESTEP))
        else:
            ro = mds.FusedRollout(env, ts, dl, trk, obstacles)
            obs = ro.run(steps, stages=6 if mode == "loop" else 4)
        outs.append(obs.cpu().numpy().copy())
    assert np.abs(outs[0][..., 0:3]).max() > 0.1
    assert np.abs(outs[1] - outs[0])[..., :16].max() < 1e-9 and np.abs(outs[2] - outs[0])[..., :16].max() < 1e-9
