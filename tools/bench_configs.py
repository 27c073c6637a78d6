"""Secondary timings of the BASELINE.json configs that are not the bench line (device-resident, CUDA events):
C2 tracking swarm (geometric, N=1), C3 CBF swarm at 16 384 envs, C4 model comparison, C5 in fp64.  Prints JSON."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multidronesim_b200 as mds  # noqa: E402
from multidronesim_b200 import scenarios  # noqa: E402


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for E in (4096, 65536, 1000000):
    sc = scenarios.tracking_swarm(E)
    ro = sc["rollout"]
    ro.run(240)
    ms = timed(lambda: ro.run(240), 4) / 240
    out[f"C2_geometric_E{E}_N1_f32"] = {"ms_per_control_step": ms, "drone_steps_per_s": E / (ms * 1e-3), "plan": ro.plan(), "max_pos_err": ro.stats_dict()["max_pos_err"]}
    del sc, ro
for E, dt in ((16384, torch.float32), (125000, torch.float64), (16384, torch.float64)):
    sc = scenarios.cbf_swarm(E, 8, order=3, dtype=dt)
    ro = sc["rollout"]
    ro.run(240)
    ms = timed(lambda: ro.run(48), 5) / 48
    out[f"C5_cbf_order3_E{E}_N8_{'f32' if dt == torch.float32 else 'f64'}"] = {"ms_per_control_step": ms, "drone_steps_per_s": 8 * E / (ms * 1e-3), "plan": ro.plan()}
    del sc, ro
sc = scenarios.cbf_swarm(16384, 8, order=2)
ro = sc["rollout"]
ro.run(240)
ms = timed(lambda: ro.run(48), 5) / 48
out["C3_cbf_order2_E16384_N8_f32"] = {"ms_per_control_step": ms, "drone_steps_per_s": 8 * 16384 / (ms * 1e-3), "plan": ro.plan(), "stats": ro.stats_dict()}
# C4: 65 536 samples, fp64
env = mds.CtrlAviary(num_drones=1, num_envs=65536, dtype=torch.float64)
obs = torch.randn(65536, 1, 20, device="cuda", dtype=torch.float64)
obs[..., 16:20] = 15000.0
lin, dyn = mds.model.LinearizedModel(env), mds.model.QuadrotorDynamics(240)
dyn.load_env_params(env)
for name, fn in (("xdot_linear12", lambda: lin.calc_xdot_from_obs(obs)), ("xdot_nonlinear", lambda: dyn.dynamics_from_obs(obs))):
    ms = timed(fn, 50)
    out[f"C4_{name}_65536_f64"] = {"ms_per_call": ms, "samples_per_s": 65536 / (ms * 1e-3)}
print(json.dumps(out, indent=1))
