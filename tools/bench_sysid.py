"""Timing of the decentralised-LQR kernels at swarm size: mds_rls_update and mds_dlqr_ctrl for E x 8 drones,
CUDA events on the launching stream, algorithmic HBM bytes against MEASURED_PEAKS.json:hbm_gbs.
usage: python tools/bench_sysid.py [envs=125000] [reps=20]   -> one JSON line per case"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multidronesim_b200 as mds  # noqa: E402
from multidronesim_b200.control import dlqr  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
N = 8
peak = 6551.7
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, cls, model, dtype, method in (
        ("rls theta_update m=9 f32", dlqr.DecentralizedLQROmega, mds.model.LinearizedOmegaModel, torch.float32, "theta_update"),
        ("rls theta_update m=9 f64", dlqr.DecentralizedLQROmega, mds.model.LinearizedOmegaModel, torch.float64, "theta_update"),
        ("rls approx_theta_update m=12 f32", dlqr.DecentralizedLQR, mds.model.LinearizedModel, torch.float32, "approx_theta_update"),
        ("rls theta_update m=12 f32", dlqr.DecentralizedLQR, mds.model.LinearizedModel, torch.float32, "theta_update")):
    env = mds.BatchedCtrlAviary(drone_model=mds.DroneModel.CF2P, num_drones=N, num_envs=E, dtype=dtype)
    c = cls(env, [model(env) for _ in range(N)])
    D, m = E * N, c.m
    g = torch.Generator(device="cuda").manual_seed(1)
    phi = torch.cat([0.1 * torch.randn(D, m, device="cuda", dtype=dtype, generator=g), 0.05 * torch.randn(D, 4, device="cuda", dtype=dtype, generator=g)], 1).contiguous()
    x1 = (phi[:, :m] + 0.01 * torch.randn(D, m, device="cuda", dtype=dtype, generator=g)).contiguous()
    if method == "theta_update" and m == 9:
        c.compute_controller()  # BEFORE the updates: identical priors share one CARE solve (distinct learned models cost one host solve each)
    ms = timed(lambda: getattr(c, method)(phi, x1), REPS)
    sz = 4 if dtype == torch.float32 else 8
    algo = (2 * (m + 4) ** 2 + 2 * (m + 4) * m + (m + 4) + 2 * m) * sz  # P and theta in + out, phi, x_{t+1} in, residual out
    gbs = algo * D / (ms * 1e-3) / 1e9
    print(json.dumps({"case": name, "drones": D, "ms": round(ms, 4), "drone_updates_per_s": D / (ms * 1e-3), "algorithmic_bytes_per_drone": algo,
                      "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3), "bound": "hbm"}))
    if method == "theta_update" and m == 9:
        env.reset()
        ref = torch.zeros(D, 11, device="cuda", dtype=dtype)
        ref[:, 2] = 1.0
        c.set_reference(ref)
        ms = timed(lambda: c.compute(env.obs), REPS)
        algo = (20 + 11 + 4 * m + 6 + 6 + 4 + 4) * sz  # obs, ref, K in; PID in + out; u, action out
        gbs = algo * D / (ms * 1e-3) / 1e9
        print(json.dumps({"case": f"dlqr compute m=9 {'f32' if sz == 4 else 'f64'}", "drones": D, "ms": round(ms, 4), "drone_steps_per_s": D / (ms * 1e-3),
                          "algorithmic_bytes_per_drone": algo, "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3), "bound": "hbm"}))
    if method == "theta_update":  # the Riccati solve for all (by now distinct) learned models, one warp per drone
        ms = timed(lambda: c.compute_controller(force_diagonal=True, solver="device"), 3)
        bad = int(c.care_status.sum())
        print(json.dumps({"case": f"care_gains m={m} {'f32' if sz == 4 else 'f64'}", "drones": D, "ms": round(ms, 3), "drone_solves_per_s": D / (ms * 1e-3),
                          "unsolved": bad, "bound": "fp64 pipe + shared memory (one warp per drone, ~9 sign-function iterations of a 2m x 2m Gauss-Jordan inverse)"}))
    del c, env
    torch.cuda.empty_cache()
