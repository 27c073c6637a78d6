"""Generate tests/golden/*.npz from the REFERENCE's own Python modules.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Every array below is produced by code imported from /root/reference (through the stand-ins
of oracle/ref_import.py) on seeded inputs that are stored alongside the outputs, so the
fixtures pin both the oracle (``-m "not gpu"`` tests) and the CUDA path (``-m gpu`` tests)
without the reference tree being present.  The env step, DSLPID and cvxopt are NOT in the
reference tree, so nothing here pins them (see oracle/__init__.py).
"""
from __future__ import annotations

import contextlib
import io
import os

import numpy as np
from scipy.spatial.transform import Rotation

from oracle import ref_import
from oracle.constants import drone_params

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def random_obs(rng, env, n, pos_scale=2.0):
    rpy = rng.uniform(-0.5, 0.5, (n, 3))
    quat = Rotation.from_euler("xyz", rpy).as_quat()
    return np.hstack([rng.uniform(-pos_scale, pos_scale, (n, 3)), quat, rpy, rng.normal(0, 1, (n, 3)),
                      rng.normal(0, 1, (n, 3)), rng.uniform(9440.3, env.MAX_RPM, (n, 4))])


def pack_ref(samples):
    return np.array([np.hstack([np.asarray(p, float), np.asarray(v, float), np.asarray(a, float), y, w]) for p, v, a, y, w in samples])


def golden_trajectories(ref):
    T = ref.traj
    ts = np.array([0.0, 0.013, 0.75, 1.9, 3.3, 7.77, 12.5, 40.0])
    out = {"t": ts}
    gens = {
        "circle": T.CircleTrajectory(r=1, v=.5, center=np.array([0, 0, 1]), yaw_rate=.1),
        "circle2": T.CircleTrajectory(r=0.7, v=1.3, center=np.array([0.2, -0.4, 0.8]), yaw_rate=-0.35),
        "lemniscate": T.Lemniscate(center=np.array([0, 0, .5]), omega=1.5, yaw_rate=.1, phase_shift=-np.pi / 4),
        "lemniscate2": T.Lemniscate(a=1.4, center=np.array([0.3, 0.1, 1.5]), omega=0.5, yaw_rate=0, phase_shift=2.1),
        "wait": T.WaitTrajectory(np.array([0.5, -0.2, 1.0]), 3.0, yaw=0.4),
        "line": T.LineTrajectory(np.array([0, 0, 0.5]), np.array([2.0, 1.0, 1.5]), speed=0.8),
        "line_short": T.LineTrajectory(np.array([0, 0, 0.5]), np.array([0.2, 0.1, 0.6]), speed=1.5),
        "line_s0": T.LineTrajectory(np.array([1.0, 0, 0.5]), np.array([-2.0, 1.0, 0.5]), speed=1.0, s0=0.3, sf=0.2),
    }
    Rz = Rotation.from_euler("z", 0.7).as_matrix()
    gens["rotate"] = T.RotateTrajectory(T.Lemniscate(center=np.array([0, 0, .5]), omega=0.8), Rz, np.array([0.1, 0.2, 0.5]))
    for name, g in gens.items():
        out[name] = pack_ref([g(float(t)) for t in ts])
    # compound on a strictly increasing clock that avoids segment boundaries (cursor quirk B21)
    comp = T.CompoundTrajectory([T.WaitTrajectory(np.array([0, 0, 0.5]), 1.0),
                                 T.LineTrajectory(np.array([0, 0, 0.5]), np.array([1.5, 0.5, 1.0]), speed=0.7),
                                 T.CircleTrajectory(r=0.5, v=0.4, center=np.array([1.0, 0.5, 1.0]), duration=4.0),
                                 T.WaitTrajectory(np.array([1.5, 0.5, 1.0]), 2.0, yaw=0.0)])
    tc = np.linspace(0.05, comp.get_total_time() + 1.0, 57)
    out["compound_t"] = tc
    out["compound"] = pack_ref([comp(float(t)) for t in tc])
    np.savez(os.path.join(OUT, "trajectories.npz"), **out)


def golden_controllers(ref):
    rng = np.random.default_rng(11)
    out = {}
    for model in ("cf2p", "cf2x"):
        env = drone_params(model, 240, 240)
        n = 64
        obs = random_obs(rng, env, n)
        lem = ref.traj.Lemniscate(center=np.array([0, 0, .5]), omega=1.5, yaw_rate=.1, phase_shift=0.3)
        refs = pack_ref([lem(float(t)) for t in rng.uniform(0, 6, n)])
        refs[:, 0:3] += rng.normal(0, 0.3, (n, 3))
        out[f"{model}_obs"], out[f"{model}_ref"] = obs, refs
        geo = ref.control.GeometricControl(env)
        acts = []
        for o, r in zip(obs, refs):
            geo.set_desired_trajectory(0, r[0:3], r[3:6], r[6:9], r[9], r[10])
            acts.append(geo.compute(o.copy()))
        out[f"{model}_geometric_action"] = np.array(acts)
        with contextlib.redirect_stdout(io.StringIO()):
            variants = {
                "torque12": ref.control.LQRController(env, ref.model.LinearizedModel(env)),
                "omega9": ref.control.LQROmegaController(env, ref.model.LinearizedOmegaModel(env), ref.control.ThrustOmegaController(env)),
                "yank10": ref.control.LQRYankOmegaController(env, ref.model.LinearizedYankOmegaModel(env), ref.control.YankOmegaController(env)),
            }
        for name, c in variants.items():
            out[f"{model}_{name}_K"] = c.K
            acts, us, us_skip = [], [], []
            for o, r in zip(obs, refs):
                c.set_desired_trajectory(0, r[0:3], r[3:6], r[6:9], r[9], r[10])
                if name != "torque12":  # fresh inner-loop state per sample
                    (c.to_controller if name == "omega9" else c.yo_controller.thrust_omega_ctrl).reset()
                    _, u2 = c.compute(o.copy(), skip_low_level=True)
                    us_skip.append(u2.copy())
                a, u = c.compute(o.copy())
                acts.append(a)
                us.append(u.copy())
            out[f"{model}_{name}_action"], out[f"{model}_{name}_u"] = np.array(acts), np.array(us)
            if us_skip:
                out[f"{model}_{name}_u_skip"] = np.array(us_skip)
        # two consecutive inner-loop calls on one controller: pins the PID state update (quirk B11)
        toc = ref.control.ThrustOmegaController(env)
        u_seq = np.hstack([rng.uniform(0.1, 0.5, (n, 1)), rng.normal(0, 1.0, (n, 3))])
        w_seq = rng.normal(0, 1.0, (2, n, 3))
        seq = []
        for k in range(n):
            toc.reset()
            a1 = toc.computeControlFromInput(u_seq[k].copy(), env.CTRL_TIMESTEP, w_seq[0, k])
            a2 = toc.computeControlFromInput(u_seq[k].copy(), env.CTRL_TIMESTEP, w_seq[1, k])
            seq.append(np.hstack([a1, a2, toc.last_omega, toc.integral_omega_e]))
        out[f"{model}_pid_u"], out[f"{model}_pid_w"], out[f"{model}_pid_out"] = u_seq, w_seq, np.array(seq)
    np.savez(os.path.join(OUT, "controllers.npz"), **out)


def golden_models(ref):
    rng = np.random.default_rng(4)
    env = drone_params("cf2p", 240, 240)
    n = 128
    obs = random_obs(rng, env, n)
    lm = ref.model.LinearizedModel(env)
    qd = ref.model.QuadrotorDynamics(env.PYB_FREQ)
    qd.load_env_params(env)
    lin = np.array([lm.calc_xdot_from_obs(o) for o in obs])
    non = np.array([ref.conv.geo_x_dot_to_linear(qd.dynamics(None, ref.conv.obs_to_geo_model(o), ref.conv.action_to_input(env, o[16:])))
                    for o in obs])
    out = {"obs": obs, "xdot_linear12": lin, "xdot_nonlinear": non, "J_dynamics": np.diag(qd.J)}
    for name, M in (("torque12", ref.model.LinearizedModel), ("omega9", ref.model.LinearizedOmegaModel),
                    ("yank10", ref.model.LinearizedYankOmegaModel)):
        m = M(env)
        out[f"{name}_A"], out[f"{name}_B"], out[f"{name}_Ahat"], out[f"{name}_Bhat"] = m.A, m.B, m.Ahat, m.Bhat
    out["lin9"] = np.array([ref.conv.obs_to_lin_model(o, dim=9) for o in obs])
    out["lin10"] = np.array([ref.conv.obs_to_lin_model(o, dim=10, env=env) for o in obs])
    out["geo18"] = np.array([ref.conv.obs_to_geo_model(o) for o in obs])
    u = np.array([ref.conv.action_to_input(env, o[16:]) for o in obs])
    out["action_to_input"] = u
    u_in = u * rng.uniform(0.2, 1.5, u.shape)
    u_in[::7, 0] *= -1.0  # exercise the u[0] >= 0 clamp
    out["input_to_action_in"] = u_in
    out["input_to_action"] = np.array([ref.conv.input_to_action(env, ui.copy()) for ui in u_in])
    np.savez(os.path.join(OUT, "models.npz"), **out)


def golden_cbf(ref):
    rng = np.random.default_rng(3)
    env = drone_params("cf2p", 240, 240)
    out = {}
    cases = [("o2_n2_obs1", 2, 2, 1), ("o2_n8_obs1", 2, 8, 1), ("o3_n7_obs0", 3, 7, 0), ("o3_n8_obs1", 3, 8, 1), ("o3_n4_obs3", 3, 4, 3)]
    for name, order, N, nobs in cases:
        Mdl = ref.model.LinearizedOmegaModel if order == 2 else ref.model.LinearizedYankOmegaModel
        poles = np.array([-2.2, -2.4]) if order == 2 else np.array([-3.0, -3.6, -5.6])
        rs, zs = (0.1, 1.0) if order == 2 else (0.125, 2.0)
        cbf = ref.cbf.DroneCBF(env, [Mdl(env) for _ in range(N)], safety_radius=rs, zscale=zs, order=order, cbf_poles=poles)
        trk = ref.cbf.DroneQPTracker(cbf, order=order, num_robots=N, xdim=cbf.xdim, env=env)
        E = 12
        xdim = cbf.xdim
        obs = np.array([random_obs(rng, env, N, pos_scale=0.6) for _ in range(E)])
        # second half: gentle near-hover formations (ring, 0.3-0.5 m spacing, slow closing speeds) where the
        # barrier rows are active but the QP stays feasible
        for e in range(E // 2, E):
            ang = 2 * np.pi * np.arange(N) / N + rng.uniform(0, 1)
            rad = rng.uniform(0.25, 0.45)
            pos = np.stack([rad * np.cos(ang), rad * np.sin(ang), 0.5 + rng.uniform(-0.15, 0.15, N)], axis=1)
            rpy = rng.uniform(-0.08, 0.08, (N, 3))
            vel = -0.4 * pos * rng.uniform(0.2, 1.0, (N, 1)) + rng.normal(0, 0.05, (N, 3))
            obs[e, :, 0:3], obs[e, :, 7:10], obs[e, :, 10:13] = pos, rpy, vel
            obs[e, :, 3:7] = Rotation.from_euler("xyz", rpy).as_quat()
            obs[e, :, 13:16] = rng.normal(0, 0.1, (N, 3))
            obs[e, :, 16:20] = env.HOVER_RPM * rng.uniform(0.95, 1.05, (N, 4))
        xdes = np.zeros((E, N, xdim))
        for e in range(E):
            for i in range(N):
                vd = rng.normal(0, .5, 3) if e < E // 2 else rng.normal(0, .05, 3)
                tail = np.hstack([vd, obs[e, i, 0:3] + rng.normal(0, .1, 3)])
                xdes[e, i] = np.hstack([0, 0, rng.uniform(-1, 1), tail]) if order == 2 else np.hstack([0, 0, rng.uniform(-1, 1), env.G * env.M, tail])
        obstacles = np.hstack([rng.uniform(-0.6, 0.6, (nobs, 3)), rng.uniform(0.05, 0.2, (nobs, 1))]) if nobs else np.zeros((0, 4))
        x_obs = [np.vstack([o[:3]] + [np.zeros(3)] * (order - 1)) for o in obstacles] if nobs else None
        r_obs = [float(o[3]) for o in obstacles] if nobs else None
        unom = np.concatenate([rng.normal(0, 0.3, (E, N, 1)), rng.normal(0, 2.0, (E, N, 3))], axis=2)
        Gs, hs, us = [], [], []
        for e in range(E):
            x = np.array([ref.conv.obs_to_lin_model(obs[e, i], dim=xdim, env=env) for i in range(N)])
            cbf.set_xdes(xdes[e])
            G, h = cbf._build_ineq_const(x, False, x_obs, r_obs)
            Gs.append(G)
            hs.append(h)
            with contextlib.redirect_stdout(io.StringIO()):
                us.append(np.array(trk.compute_control(obs[e], xdes[e], unom[e].copy(), x_obs=x_obs, obs_r_list=r_obs)))
        out.update({f"{name}_obs": obs, f"{name}_xdes": xdes, f"{name}_obstacles": obstacles, f"{name}_unom": unom,
                    f"{name}_G": np.array(Gs), f"{name}_h": np.array(hs), f"{name}_usafe_oracle_qp": np.array(us),
                    f"{name}_Kcbf": cbf.Kcbf, f"{name}_umax": np.asarray(cbf.umax, float)})
    np.savez_compressed(os.path.join(OUT, "cbf_rows.npz"), **out)


def golden_dlqr(ref):
    """Decentralised-LQR model learning and control (control/dlqr/*.py), run by the reference classes themselves."""
    import importlib
    rng = np.random.default_rng(17)
    env = drone_params("cf2p", 240, 240)
    out = {"dt": env.CTRL_TIMESTEP}
    with contextlib.redirect_stdout(io.StringIO()):
        m_om = importlib.import_module("control.dlqr.decentralized_lqr_omega")
        m_yo = importlib.import_module("control.dlqr.decentralized_lqr_yank_omega")
        m_12 = importlib.import_module("control.dlqr.decentralized_lqr")
        m_cf = importlib.import_module("control.dlqr.decentralized_yolqr_crazyflie")
    N, T = 3, 4

    def thetas(d):
        return np.array([d.get_thetai(i) for i in range(N)])

    def draws(m, u_scale):
        phis = np.concatenate([rng.normal(0, 0.1, (T, N, m)), rng.normal(0, 1.0, (T, N, 4)) * u_scale], axis=2)
        x1 = phis[:, :, :m] + rng.normal(0, 0.01, (T, N, m))
        return phis, x1

    def run(tag, make, method, m, u_scale, **kw):
        with contextlib.redirect_stdout(io.StringIO()):
            d = make()
        phis, x1 = draws(m, u_scale)
        out[f"{tag}_phi"], out[f"{tag}_x1"] = phis, x1
        out[f"{tag}_theta0"], out[f"{tag}_P0"] = thetas(d), np.array(d.P, float).copy()
        th, Ps = [], []
        for t in range(T):
            with contextlib.redirect_stdout(io.StringIO()):
                getattr(d, method)([p.copy() for p in phis[t]], [x.copy() for x in x1[t]], **kw)
            th.append(thetas(d))
            Ps.append(np.array(d.P, float).copy())
        out[f"{tag}_theta"], out[f"{tag}_P"] = np.array(th), np.array(Ps)
        return d

    u9 = np.array([0.05, 0.1, 0.1, 0.1])
    u12 = np.array([0.05, 1e-4, 1e-4, 1e-4])
    u10 = np.array([2.0, 0.1, 0.1, 0.1])
    om = lambda: m_om.DecentralizedLQROmega(env, [ref.model.LinearizedOmegaModel(env) for _ in range(N)])
    yo = lambda: m_yo.DecentralizedLQRYankOmega(env, [ref.model.LinearizedYankOmegaModel(env) for _ in range(N)])
    t12 = lambda: m_12.DecentralizedLQR(env, [ref.model.LinearizedModel(env) for _ in range(N)])
    cf = lambda: m_cf.DecentralizedYOLQRCrazyflie(env, [ref.model.LinearizedYankOmegaModel(env) for _ in range(N)], np.eye(10), np.eye(4))
    d_om = run("omega9_update", om, "theta_update", 9, u9)
    run("omega9_update2", om, "theta_update2", 9, u9)            # P holds V (information form)
    run("yank10_update", yo, "theta_update", 10, u10)
    d_12 = run("torque12_update", t12, "theta_update", 12, u12)  # project_theta once after the loop
    run("torque12_approx", t12, "approx_theta_update", 12, u12)  # project_theta inside the loop
    run("cf10_approx", cf, "approx_theta_update", 10, u10, project=True)
    run("cf10_approx_noproj", cf, "approx_theta_update", 10, u10, project=False)
    # control law with the learned models: compute_controller (CARE on the learned theta) then compute(obs)
    obs = random_obs(rng, env, N, pos_scale=0.5)
    lem = ref.traj.Lemniscate(center=np.array([0, 0, .5]), omega=1.5, yaw_rate=.1)
    refs = pack_ref([lem(float(t)) for t in rng.uniform(0, 6, N)])
    out["ctrl_obs"], out["ctrl_ref"] = obs, refs
    for tag, d in (("omega9", d_om), ("torque12", d_12)):
        d.theta = np.hstack([d.Astar, d.Bstar]).T.copy()  # a stabilisable model: the noisy prior, perturbed per robot
        for i in range(N):
            th = d.get_thetai(i)
            d.overwrite_theta(th * (1.0 + 0.02 * (i + 1)), i)
        d.compute_controller()
        for i in range(N):
            d.set_desired_trajectory(i, refs[i, 0:3], refs[i, 3:6], refs[i, 6:9], refs[i, 9], refs[i, 10])
        action, u = d.compute(obs.copy())
        out[f"ctrl_{tag}_theta"], out[f"ctrl_{tag}_K"] = thetas(d), d.K
        out[f"ctrl_{tag}_action"], out[f"ctrl_{tag}_u"] = np.array(action), np.array(u, float)
    # FedCE wrapper (FedCE/FederatedLearning.py): YOState lists in, approx_theta_update(project=True) + u = -K e out
    fed_mod = importlib.import_module("FedCE.FederatedLearning")
    Nf, Tf = 2, 5
    with contextlib.redirect_stdout(io.StringIO()):
        fl = fed_mod.FederatedLearning(env, [ref.model.LinearizedYankOmegaModel(env) for _ in range(Nf)], np.eye(10), np.eye(4), num_drones=Nf)
    mg = env.M * env.G
    xs = np.concatenate([rng.uniform(-0.2, 0.2, (Tf, Nf, 3)), mg + rng.normal(0, 0.02, (Tf, Nf, 1)), rng.normal(0, 0.3, (Tf, Nf, 3)),
                         rng.uniform(-1, 1, (Tf, Nf, 3))], axis=2)
    xdes = np.zeros((Tf, Nf, 10))
    xdes[:, :, 2] = rng.uniform(-0.5, 0.5, (Tf, Nf))
    xdes[:, :, 3] = mg
    xdes[:, :, 4:7] = rng.normal(0, 0.1, (Tf, Nf, 3))
    xdes[:, :, 7:10] = rng.uniform(-1, 1, (Tf, Nf, 3))
    us = np.concatenate([rng.normal(0, 2.0, (Tf, Nf, 1)), rng.normal(0, 0.1, (Tf, Nf, 3))], axis=2)
    yo = lambda v: m_cf.YOState(*[float(a) for a in v])
    fed_th, fed_P = [], []
    for t in range(Tf):
        with contextlib.redirect_stdout(io.StringIO()):
            fl.update([yo(v) for v in xs[t]], [yo(v) for v in xdes[t]], [u.copy() for u in us[t]])
        fed_th.append(np.array([fl.dLQR.get_thetai(i) for i in range(Nf)]))
        fed_P.append(np.array(fl.dLQR.P, float).copy())
    with contextlib.redirect_stdout(io.StringIO()):
        fl.calc_controller()
        fed_u = fl.lqr_control([yo(v) for v in xs[-1]], [yo(v) for v in xdes[-1]])
    out["fed_x"], out["fed_xdes"], out["fed_u"] = xs, xdes, us
    out["fed_theta"], out["fed_P"], out["fed_K"], out["fed_ctrl_u"] = np.array(fed_th), np.array(fed_P), fl.dLQR.K, np.array(fed_u)
    np.savez_compressed(os.path.join(OUT, "dlqr.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_import.load()
    golden_trajectories(ref)
    golden_controllers(ref)
    golden_models(ref)
    golden_cbf(ref)
    golden_dlqr(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
