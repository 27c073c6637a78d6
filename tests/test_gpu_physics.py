"""GPU parity: mds_physics_step (through BatchedCtrlAviary.step) vs the oracle env step.
Tolerances from BASELINE.json north_star: fp64 1e-9 relative per step, fp32 1e-5 relative per
step, <= 1 mm over 1 s of hover."""
import numpy as np
import pytest
import torch

from helpers import random_state, scaled_err
from oracle.aviary import OracleCtrlAviary
from oracle.constants import DroneModel as ODM, Physics as OPH

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-9, torch.float32: 1e-5}


def make_pair(E, N, model, physics, dtype, pyb_freq=240, ctrl_freq=240, **kw):
    import multidronesim_b200 as mds
    env = mds.BatchedCtrlAviary(drone_model=mds.DroneModel(model), num_drones=N, physics=mds.Physics(physics), pyb_freq=pyb_freq,
                                ctrl_freq=ctrl_freq, num_envs=E, dtype=dtype, **kw)
    oracles = [OracleCtrlAviary(ODM(model), N, physics=OPH(physics), pyb_freq=pyb_freq, ctrl_freq=ctrl_freq, **kw) for _ in range(E)]
    return env, oracles


def set_both(env, oracles, state):
    pos, quat, vel, w, rpm = state
    env.set_state(pos, quat, vel, w, rpm)
    for e, o in enumerate(oracles):
        o.set_state(pos[e], quat[e], vel[e], w[e], rpm[e])


def obs_err(got, want):
    """per-field scaled error of a [.., 20] observation (fields share a scale: pos, quat, rpy, vel, ang vel, rpm)"""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return max(scaled_err(got[..., a:b], want[..., a:b]) for a, b in ((0, 3), (3, 7), (7, 10), (10, 13), (13, 16), (16, 20)))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("model", ["cf2p", "cf2x"])
@pytest.mark.parametrize("physics,N,freqs", [("dyn", 1, (240, 240)), ("dyn", 3, (480, 240)), ("dyn_gnd_drag_dw", 1, (240, 240)),
                                             ("dyn_gnd_drag_dw", 8, (240, 240)), ("dyn_gnd_drag_dw", 5, (720, 240))])
def test_single_step_parity(dtype, model, physics, N, freqs, lib_built):
    E = 7
    env, oracles = make_pair(E, N, model, physics, dtype, *freqs)
    rng = np.random.default_rng(5)
    state = random_state(rng, oracles[0], (E, N), pos_scale=0.5, z_min=0.02)
    if dtype == torch.float32:  # start both from the same representable state
        state = tuple(np.asarray(s, np.float32).astype(np.float64) for s in state)
    set_both(env, oracles, state)
    action = rng.uniform(0, 1.15 * oracles[0].MAX_RPM, (E, N, 4))  # some above MAX_RPM: exercises the clip
    if dtype == torch.float32:
        action = action.astype(np.float32).astype(np.float64)
    for step in range(3):
        obs, reward, term, trunc, info = env.step(torch.as_tensor(action, device="cuda", dtype=dtype))
        want = np.array([o.step(action[e])[0] for e, o in enumerate(oracles)])
        assert obs.shape == (E, N, 20)
        err = obs_err(obs.cpu().numpy(), want)
        assert err < TOL[dtype] * (step + 1), (step, err)
        if dtype == torch.float32:  # per-step tolerance: re-synchronise the oracle to the device state
            for e, o in enumerate(oracles):
                o.set_state(env.pos[e].cpu().numpy(), env.quat[e].cpu().numpy(), env.vel[e].cpu().numpy(), env.rpy_rates[e].cpu().numpy(),
                            env.last_clipped_action[e].cpu().numpy())
                o.ang_v = obs[e, :, 13:16].double().cpu().numpy()
    assert float(reward[0]) == -1 and not bool(term.any()) and not bool(trunc.any()) and info == {"answer": 42}


def test_default_spawn_and_reset(lib_built):
    import multidronesim_b200 as mds
    env = mds.BatchedCtrlAviary(num_drones=3, num_envs=2, dtype=torch.float64)
    o = OracleCtrlAviary(ODM.CF2X, 3)
    obs, info = env.reset()
    assert np.allclose(obs[0].cpu().numpy(), o.reset()[0], atol=1e-15) and info == {"answer": 42}
    assert env.M == o.M and env.MAX_RPM == o.MAX_RPM and np.allclose(env.J, o.J) and env.CTRL_TIMESTEP == 1 / 240


def test_ground_clamp_and_downwash(lib_built):
    """drone 1 directly above drone 0: downwash pushes 0 down; a falling drone stops at the floor."""
    import multidronesim_b200 as mds
    for dtype in (torch.float64, torch.float32):
        env, (o,) = make_pair(1, 2, "cf2p", "dyn_gnd_drag_dw", dtype)
        pos = np.array([[[0, 0, 0.3], [0.02, 0, 0.6]]])
        quat = np.tile([0, 0, 0, 1.0], (1, 2, 1))
        z = np.zeros((1, 2, 3))
        set_both(env, [o], (pos, quat, z, z, np.zeros((1, 2, 4))))
        act = np.full((1, 2, 4), o.HOVER_RPM)
        act[0, 0] = 0.0  # drone 0 free-falls to the floor
        for _ in range(150):
            obs = env.step(torch.as_tensor(act, device="cuda", dtype=dtype))[0]
            want = o.step(act[0])[0]
        assert abs(float(obs[0, 0, 2]) - o.Z_FLOOR) < 1e-6 and want[0, 2] == o.Z_FLOOR
        assert obs_err(obs[0].cpu().numpy(), want) < (1e-9 if dtype == torch.float64 else 2e-4)


def test_hover_1s_under_1mm_fp32(lib_built):
    """fp32, 240 steps of constant hover RPM with a small tilt: position within 1 mm of the fp64 oracle."""
    E, N = 4, 2
    env, oracles = make_pair(E, N, "cf2p", "dyn_gnd_drag_dw", torch.float32)
    rng = np.random.default_rng(9)
    pos = rng.uniform(-1, 1, (E, N, 3)); pos[..., 2] = 1.0 + 0.2 * np.arange(N)
    rpy = rng.uniform(-0.02, 0.02, (E, N, 3))
    from helpers import euler_to_quat
    state = tuple(np.asarray(s, np.float32).astype(np.float64) for s in (pos, euler_to_quat(rpy), np.zeros((E, N, 3)), np.zeros((E, N, 3)), np.zeros((E, N, 4))))
    set_both(env, oracles, state)
    act = np.full((E, N, 4), np.float32(oracles[0].HOVER_RPM)).astype(np.float64)
    a = torch.as_tensor(act, device="cuda", dtype=torch.float32)
    for _ in range(240):
        obs = env.step(a)[0]
    want = np.array([[o.step(act[e])[0] for _ in range(240)][-1] for e, o in enumerate(oracles)])
    assert np.max(np.abs(obs[..., 0:3].double().cpu().numpy() - want[..., 0:3])) < 1e-3


def test_host_buffers_and_errors(lib_built):
    import multidronesim_b200 as mds
    env = mds.BatchedCtrlAviary(num_drones=2, num_envs=3, dtype=torch.float32)
    act = np.full((3, 2, 4), env.HOVER_RPM)
    obs_np = env.step(act)[0].cpu().numpy().copy()           # numpy action is staged H2D
    env.reset()
    a_pin = torch.full((3, 2, 4), env.HOVER_RPM, dtype=torch.float32).pin_memory()
    o_pin = torch.empty(3, 2, 20, dtype=torch.float32).pin_memory()
    env.step_host(a_pin, o_pin)
    torch.cuda.synchronize()
    assert np.array_equal(o_pin.numpy(), obs_np)
    with pytest.raises(ValueError):
        env.step(np.zeros((3, 2, 3)))
    with pytest.raises(mds._lib.MdsError):
        mds._lib.call("mds_physics_step", torch.float32, env._prm, env._state_struct(), None, None, None, 3, 2, None)
    with pytest.raises(ValueError):
        mds.BatchedCtrlAviary(num_drones=2, pyb_freq=250, ctrl_freq=240)


def test_external_force(lib_built):
    env, (o,) = make_pair(1, 1, "cf2p", "dyn", torch.float64)
    f = np.array([[[0.00025, 0, 0]]])  # the reference's wind_force (EnvGeometric.py:34)
    env.set_external_force(f)
    o.ext_force = f[0]
    act = np.full((1, 1, 4), o.HOVER_RPM)
    for _ in range(10):
        obs = env.step(act)[0]
        want = o.step(act[0])[0]
    assert obs_err(obs[0].cpu().numpy(), want) < 1e-9 and want[0, 0] > 0


@pytest.mark.gpu
def test_plain_host_abi_numpy_only(lib_built):
    """The boundary as a numpy-only caller (the reference has no torch) would use it: device buffers from
    mds_device_alloc, state uploaded with mds_copy_to_device, `obs = step(action)` with HOST arrays through
    mds_physics_step_host_f64 -- compared with the oracle env step by step (INTEGRATION.md stub)."""
    import ctypes as C
    from multidronesim_b200 import _lib
    from multidronesim_b200.constants import DroneConstants
    from multidronesim_b200.enums import DroneModel, Physics
    lib = lib_built
    N, steps = 3, 12
    rng = np.random.default_rng(5)
    o = OracleCtrlAviary(ODM.CF2P, N, initial_xyzs=np.array([[0, 0, 0.5], [0.3, 0, 0.9], [0.1, 0.05, 1.4]]), physics=OPH.DYN_GND_DRAG_DW)
    prm = DroneConstants(DroneModel.CF2P, Physics.DYN_GND_DRAG_DW, 240, 240).c_params()

    def dev(nbytes):
        p = C.c_void_p()
        assert lib.mds_device_alloc(nbytes, C.byref(p)) == 0
        return p
    bufs = {k: dev(N * w * 8) for k, w in (("pos_wx", 4), ("quat", 4), ("vel_wy", 4), ("rpm", 4), ("wz", 1), ("action", 4), ("obs", 20))}
    pos_wx = np.zeros((N, 4)); pos_wx[:, :3] = o.pos
    quat = np.tile([0.0, 0, 0, 1], (N, 1))
    for name, arr in (("pos_wx", pos_wx), ("quat", quat)):
        assert lib.mds_copy_to_device(bufs[name], arr.ctypes.data_as(C.c_void_p), arr.nbytes, None) == 0
    assert lib.mds_stream_synchronize(None) == 0
    st = _lib.State(*(bufs[k].value for k in ("pos_wx", "quat", "vel_wy", "rpm", "wz")))
    obs = np.zeros((N, 20))
    for k in range(steps):
        action = rng.uniform(12000, 16000, (N, 4))
        rc = lib.mds_physics_step_host_f64(C.byref(prm), st, action.ctypes.data_as(C.c_void_p), bufs["action"], bufs["obs"],
                                           obs.ctypes.data_as(C.c_void_p), 1, N, None)
        assert rc == 0, lib.mds_last_error()
        want = o.step(action)[0]
        assert np.max(np.abs(obs - want) / (1 + np.abs(want))) < 1e-9
    for p in bufs.values():
        assert lib.mds_device_free(p) == 0
