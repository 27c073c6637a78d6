"""GPU parity vs the REFERENCE-generated fixtures: trajectories, geometric / LQR controllers, inner
PID, linear / nonlinear xdot.  fp64 1e-9, fp32 1e-5 (relative, per call)."""
import numpy as np
import pytest
import torch

from helpers import rel_err, scaled_err

pytestmark = pytest.mark.gpu
TOL = {torch.float64: 1e-9, torch.float32: 1e-5}
DTYPES = [torch.float64, torch.float32]


def traj_gens(T):
    Rz = np.array([[np.cos(.7), -np.sin(.7), 0], [np.sin(.7), np.cos(.7), 0], [0, 0, 1]])
    return {
        "circle": T.CircleTrajectory(r=1, v=.5, center=np.array([0, 0, 1]), yaw_rate=.1),
        "circle2": T.CircleTrajectory(r=0.7, v=1.3, center=np.array([0.2, -0.4, 0.8]), yaw_rate=-0.35),
        "lemniscate": T.Lemniscate(center=np.array([0, 0, .5]), omega=1.5, yaw_rate=.1, phase_shift=-np.pi / 4),
        "lemniscate2": T.Lemniscate(a=1.4, center=np.array([0.3, 0.1, 1.5]), omega=0.5, yaw_rate=0, phase_shift=2.1),
        "wait": T.WaitTrajectory(np.array([0.5, -0.2, 1.0]), 3.0, yaw=0.4),
        "line": T.LineTrajectory(np.array([0, 0, 0.5]), np.array([2.0, 1.0, 1.5]), speed=0.8),
        "line_short": T.LineTrajectory(np.array([0, 0, 0.5]), np.array([0.2, 0.1, 0.6]), speed=1.5),
        "line_s0": T.LineTrajectory(np.array([1.0, 0, 0.5]), np.array([-2.0, 1.0, 0.5]), speed=1.0, s0=0.3, sf=0.2),
        "rotate": T.RotateTrajectory(T.Lemniscate(center=np.array([0, 0, .5]), omega=0.8), Rz, np.array([0.1, 0.2, 0.5])),
    }


@pytest.mark.parametrize("dtype", DTYPES)
def test_trajectories(golden, dtype, lib_built):
    import multidronesim_b200.trajectories as T
    g = golden["trajectories"]
    gens = traj_gens(T)
    names = list(gens)
    ts = T.TrajectorySet([gens[n] for n in names], dtype=dtype)
    for k, t in enumerate(g["t"]):
        got = ts.eval(float(t)).cpu().numpy()
        for i, n in enumerate(names):
            assert scaled_err(got[i, :9], g[n][k, :9]) < TOL[dtype], (n, t)
            assert scaled_err(got[i, 9:], g[n][k, 9:]) < TOL[dtype], (n, t)
    comp = T.CompoundTrajectory([T.WaitTrajectory(np.array([0, 0, 0.5]), 1.0),
                                 T.LineTrajectory(np.array([0, 0, 0.5]), np.array([1.5, 0.5, 1.0]), speed=0.7),
                                 T.CircleTrajectory(r=0.5, v=0.4, center=np.array([1.0, 0.5, 1.0]), duration=4.0),
                                 T.WaitTrajectory(np.array([1.5, 0.5, 1.0]), 2.0, yaw=0.0)])
    tc = T.TrajectorySet([comp, gens["circle"]], dtype=dtype)
    for k, t in enumerate(g["compound_t"]):
        got = tc.eval(float(t)).cpu().numpy()[0]
        assert scaled_err(got, g["compound"][k]) < TOL[dtype], t
    pos, vel, acc, yaw, om = ts(0.75)  # reference call contract
    assert pos.shape == (len(names), 3) and yaw.shape == (len(names),)
    assert abs(float(yaw[0]) - 6.358185307179586) < 1e-5  # quirk B18


def make_env(model, n, dtype):
    import multidronesim_b200 as mds
    return mds.BatchedCtrlAviary(drone_model=mds.DroneModel(model), num_drones=1, num_envs=n, dtype=dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("model", ["cf2p", "cf2x"])
def test_controllers(golden, dtype, model, lib_built):
    import multidronesim_b200 as mds
    g = golden["controllers"]
    obs_np, ref_np = g[f"{model}_obs"], g[f"{model}_ref"]
    n = obs_np.shape[0]
    env = make_env(model, n, dtype)
    obs = torch.as_tensor(obs_np, device="cuda", dtype=dtype).reshape(n, 1, 20).contiguous()
    ref = torch.as_tensor(ref_np, device="cuda", dtype=dtype).contiguous()
    geo = mds.control.GeometricControl(env)
    geo.set_reference(ref)
    act = geo.compute(obs).cpu().numpy().reshape(n, 4)
    assert rel_err(act, g[f"{model}_geometric_action"]) < TOL[dtype]
    # set_desired_trajectory path gives the same answer as the zero-copy reference buffer
    geo.set_desired_trajectory(None, ref_np[:, 0:3], ref_np[:, 3:6], ref_np[:, 6:9], ref_np[:, 9], ref_np[:, 10])
    assert np.array_equal(geo.compute(obs).cpu().numpy().reshape(n, 4), act)
    M = mds.model
    variants = {
        "torque12": lambda: mds.control.LQRController(env, M.LinearizedModel(env)),
        "omega9": lambda: mds.control.LQROmegaController(env, M.LinearizedOmegaModel(env), mds.control.ThrustOmegaController(env)),
        "yank10": lambda: mds.control.LQRYankOmegaController(env, M.LinearizedYankOmegaModel(env), mds.control.YankOmegaController(env)),
    }
    for name, mk in variants.items():
        c = mk()
        Kg = g[f"{model}_{name}_K"]
        assert np.allclose(c.K, Kg, rtol=1e-8, atol=1e-9 * np.abs(Kg).max())
        c.set_reference(ref)
        if name != "torque12":
            _, u_skip = c.compute(obs, skip_low_level=True)
            assert scaled_err(u_skip.cpu().numpy().reshape(n, 4), g[f"{model}_{name}_u_skip"]) < 10 * TOL[dtype], name
        a, u = c.compute(obs)
        # u mixes thrust O(0.3) with rates O(10): compare column-wise against the column scale
        ug, ud = g[f"{model}_{name}_u"], u.cpu().numpy().reshape(n, 4)
        for col in range(4):
            assert scaled_err(ud[:, col], ug[:, col]) < 10 * TOL[dtype], (name, col)
        assert rel_err(a.cpu().numpy().reshape(n, 4), g[f"{model}_{name}_action"]) < 10 * TOL[dtype], name
    # inner loop state update over two consecutive calls (quirk B11)
    pid = mds.control.ThrustOmegaController(env)
    u_seq = torch.as_tensor(g[f"{model}_pid_u"], device="cuda", dtype=dtype).reshape(n, 1, 4).contiguous()
    fake = torch.zeros(n, 1, 20, device="cuda", dtype=dtype)
    fake[..., 6] = 1.0  # identity attitude: body rates == world rates
    outs = []
    for k in range(2):
        fake[..., 13:16] = torch.as_tensor(g[f"{model}_pid_w"][k], device="cuda", dtype=dtype).reshape(n, 1, 3)
        outs.append(pid.compute_from_obs(u_seq, fake).cpu().numpy().reshape(n, 4).copy())
    want = g[f"{model}_pid_out"]
    assert rel_err(outs[0], want[:, 0:4]) < TOL[dtype] and rel_err(outs[1], want[:, 4:8]) < TOL[dtype]
    assert scaled_err(pid.last_omega.cpu().numpy(), want[:, 8:11]) < TOL[dtype]
    assert scaled_err(pid.integral_omega_e.cpu().numpy(), want[:, 11:14]) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_models(golden, dtype, lib_built):
    import multidronesim_b200 as mds
    g = golden["models"]
    n = g["obs"].shape[0]
    env = make_env("cf2p", n, dtype)
    obs = torch.as_tensor(g["obs"], device="cuda", dtype=dtype).contiguous()
    lin = mds.model.LinearizedModel(env)
    got = lin.calc_xdot_from_obs(obs).cpu().numpy()
    want = g["xdot_linear12"]
    for a, b in ((0, 3), (3, 6), (6, 9), (9, 12)):
        assert scaled_err(got[:, a:b], want[:, a:b]) < 20 * TOL[dtype]
    qd = mds.model.QuadrotorDynamics(env.PYB_FREQ)
    qd.load_env_params(env)
    assert np.allclose(np.diag(qd.J), g["J_dynamics"])
    got = qd.dynamics_from_obs(obs).cpu().numpy()
    want = g["xdot_nonlinear"]
    for a, b in ((0, 3), (3, 6), (6, 9), (9, 12)):
        assert scaled_err(got[:, a:b], want[:, a:b]) < 20 * TOL[dtype]
    for name, cls in (("torque12", mds.model.LinearizedModel), ("omega9", mds.model.LinearizedOmegaModel), ("yank10", mds.model.LinearizedYankOmegaModel)):
        m = cls(env)
        for key in ("A", "B", "Ahat", "Bhat"):
            assert np.array_equal(getattr(m, key), g[f"{name}_{key}"]), (name, key)
    # right-sized 9/10-dim xdot (builder-defined, quirk B23) against the oracle's definition
    from oracle.constants import drone_params
    from oracle.models import xdot_linear_generic
    oenv = drone_params("cf2p", 240, 240)
    for kind, cls in (("omega9", mds.model.LinearizedOmegaModel), ("yank10", mds.model.LinearizedYankOmegaModel)):
        got = cls(env).calc_xdot_from_obs(obs).cpu().numpy()
        want = np.array([xdot_linear_generic(oenv, o, kind) for o in g["obs"]])
        assert scaled_err(got, want) < 20 * TOL[dtype]
    # conversions
    U = mds.utils
    assert scaled_err(U.obs_to_lin_model(obs, 10, env).cpu().numpy(), g["lin10"]) < TOL[dtype]
    assert scaled_err(U.obs_to_geo_model(obs).cpu().numpy(), g["geo18"]) < TOL[dtype]
    assert rel_err(U.input_to_action(env, torch.as_tensor(g["input_to_action_in"], device="cuda", dtype=dtype)).cpu().numpy(), g["input_to_action"]) < 10 * TOL[dtype]


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_dslpid_config1_closed_loop(dtype, tol, lib_built):
    """BASELINE config 1 on the device: 2 CF2X drones, Physics.DYN, 240 Hz, upstream DSL PID with the halved gains of
    MultiDroneExample.py:85-92 climbing 1 m -- step for step against the oracle's DslPid + env (both restated from
    upstream, parity unpinned), 2 s closed loop, several environments with different starts."""
    import multidronesim_b200 as mds
    from oracle.aviary import OracleCtrlAviary
    from oracle.constants import DroneModel as ODM, Physics as OPH
    from oracle.controllers import DslPid
    E, N, steps = 3, 2, 480
    rng = np.random.default_rng(2)
    init = np.array([[[1.0, 0, 0.05], [-1.0, 0, 0.05]]]) + rng.normal(0, 0.05, (E, N, 3)) * np.array([1, 1, 0])
    target = init + np.array([0, 0, 1.0])
    env = mds.CtrlAviary(drone_model=mds.DroneModel.CF2X, num_drones=N, physics=mds.Physics.DYN, pyb_freq=240, ctrl_freq=240,
                         initial_xyzs=init, num_envs=E, dtype=dtype)
    ctrl = mds.control.DSLPIDControl(env, gain_scale=0.5)
    obs = env.step(torch.zeros(E, N, 4, device="cuda", dtype=dtype))[0]
    tgt = torch.as_tensor(target, device="cuda", dtype=dtype)
    for _ in range(steps):
        rpm, pos_e, _ = ctrl.computeControlFromState(env.CTRL_TIMESTEP, obs, tgt, torch.zeros(3))
        obs = env.step(rpm)[0]
    got = obs.double().cpu().numpy()
    for e in range(E):
        o = OracleCtrlAviary(ODM.CF2X, N, initial_xyzs=init[e], physics=OPH.DYN)
        cs = [DslPid(o, gain_scale=0.5) for _ in range(N)]
        ob = o.step(np.zeros((N, 4)))[0]
        for _ in range(steps):
            ob = o.step(np.array([cs[j].compute_from_state(o.CTRL_TIMESTEP, ob[j], target[e, j])[0] for j in range(N)]))[0]
        assert np.max(np.abs(got[e, :, 0:3] - ob[:, 0:3])) < tol
        assert np.max(np.abs(ob[:, 2] - target[e, :, 2])) < 0.6  # it is climbing towards the target
    # the same closed loop as ONE rollout call (MDS_CTRL_DSLPID, WaitTrajectory targets): identical to the per-call loop
    env2 = mds.CtrlAviary(drone_model=mds.DroneModel.CF2X, num_drones=N, physics=mds.Physics.DYN, initial_xyzs=init, num_envs=E, dtype=dtype)
    ctrl2 = mds.control.DSLPIDControl(env2, gain_scale=0.5)
    waits = [mds.trajectories.WaitTrajectory(target[e, j], 1e6) for e in range(E) for j in range(N)]
    ro = mds.FusedRollout(env2, mds.trajectories.TrajectorySet(waits, dtype=dtype), ctrl2)
    env2.step(torch.zeros(E, N, 4, device="cuda", dtype=dtype))
    obs2 = ro.run(steps).double().cpu().numpy()
    assert np.max(np.abs(obs2 - got)) < 1e-12 * (1 + np.max(np.abs(got)))


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-5)])
def test_linear_roll_out(dtype, tol, lib_built):
    """roll_out_linear_system (simulations/CompareModels.py:82-95) on device along a logged closed-loop flight: the exact
    per-interval update vs the oracle's tight-tolerance integration of the same ODE."""
    import multidronesim_b200 as mds
    from oracle import models as om
    from oracle.aviary import OracleCtrlAviary
    from oracle.constants import DroneModel as ODM, Physics as OPH
    E, N, T = 3, 2, 60
    env = mds.CtrlAviary(drone_model=mds.DroneModel.CF2P, num_drones=N, physics=mds.Physics.DYN, num_envs=E, dtype=dtype,
                         initial_xyzs=np.array([[0.0, 0, 1.0], [0.5, 0.2, 1.2]]))
    ctrl = mds.control.GeometricControl(env)
    trajs = mds.trajectories.TrajectorySet([mds.trajectories.CircleTrajectory(r=0.5, v=0.4, center=np.array([0.0, 0.0, 1.0]), yaw_rate=0.0),
                                            mds.trajectories.Lemniscate(a=0.5, center=np.array([0.5, 0.2, 1.2]), omega=1.0)] * E, dtype=dtype)
    log = torch.zeros(T, E, N, 20, device="cuda", dtype=dtype)
    mds.FusedRollout(env, trajs, ctrl).run(T, obs_log=log, log_every=1)
    x = mds.model.LinearizedModel(env).roll_out(log).double().cpu().numpy()
    obs = log.double().cpu().numpy()
    oenv = OracleCtrlAviary(ODM.CF2P, 1, physics=OPH.DYN)
    ts = env.CTRL_TIMESTEP * np.arange(T)
    worst = 0.0
    for e in range(E):
        for n in range(N):
            want = om.roll_out_linear_system(oenv, obs[:, e, n], ts)
            worst = max(worst, float(np.max(np.abs(x[:, e, n] - want) / (1 + np.abs(want)))))
    assert worst < tol, worst
    assert np.abs(x[-1] - x[0]).max() > 1e-3  # the roll-out moved
