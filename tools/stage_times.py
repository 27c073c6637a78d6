"""Per-stage device times of the per-call path (CUDA events), C5-style swarm.  Prints one JSON line.
usage: python tools/stage_times.py [envs] [sim_steps_before_timing]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multidronesim_b200 as mds  # noqa: E402
from multidronesim_b200 import _lib, scenarios  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 96
order = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc = scenarios.cbf_swarm(E, 8, order=order)
env, ctrl, trk, trajs, ro = sc["env"], sc["ctrl"], sc["tracker"], sc["trajs"], sc["rollout"]
ro.run(pre)  # get into the regime the bench times
torch.cuda.synchronize()
pipe = mds.rollout.PerCallPipeline(env, ctrl, trk, sc["obstacles"])
t = ro.t
D = env.NUM_TOTAL
ev = lambda: torch.cuda.Event(enable_timing=True)
names = ["traj_eval", "lqr_ctrl", "cbf_prepare", "cbf_qp", "lowlevel", "physics_step"]
acc = {n: 0.0 for n in names}
reps = 10
for rep in range(reps + 2):
    marks = [ev() for _ in range(7)]
    marks[0].record()
    ref = trajs.eval(t); marks[1].record()
    ctrl.set_reference(ref)
    _, u = ctrl.compute(env.obs, skip_low_level=True); marks[2].record()
    _lib.call("mds_cbf_prepare", env.dtype, env._prm, trk.order, pipe.mg, _lib.ptr(ref), _lib.ptr(u), _lib.ptr(pipe.xdes), D, _lib.stream_ptr(env.device)); marks[3].record()
    us = trk.compute_control(env.obs, pipe.xdes, u, x_obs=pipe.obst); marks[4].record()
    if order == 2:
        us[..., 0] += pipe.mg
    act = ctrl.compute_low_level(us, env.obs); marks[5].record()
    env.step(act); marks[6].record()
    torch.cuda.synchronize()
    t += env.CTRL_TIMESTEP
    if rep >= 2:
        for i, n in enumerate(names):
            acc[n] += marks[i].elapsed_time(marks[i + 1]) / reps
tot = sum(acc.values())
e0, e1 = ev(), ev()
ro.run(24); torch.cuda.synchronize()
e0.record(); ro.run(24); e1.record(); torch.cuda.synchronize()
fused = e0.elapsed_time(e1) / 24
print(json.dumps({"envs": E, "drones": D, "order": order, "ms_per_control_step": {k: round(v, 4) for k, v in acc.items()}, "sum_ms": round(tot, 4),
                  "fused_ms_per_control_step": round(fused, 4), "qp_status_counts": torch.bincount(trk.status, minlength=3).tolist(),
                  "qp_mean_iters": float(trk.iters.float().mean()), "qp_active_frac": float((trk.iters > 0).float().mean())}))
