"""Small driver for ncu: C5-style swarm; `pre` warm-up control steps in launches of 24 (rollout_loop_kernel), one timed
launch of `n` steps, then one control step as the per-call pair (ctrl_step_kernel, physics_step_kernel).
usage: python tools/prof_rollout.py [envs] [pre] [n] [f32|f64]"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multidronesim_b200 import scenarios  # noqa: E402
E = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 240
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dt = torch.float64 if (len(sys.argv) > 4 and sys.argv[4] == "f64") else torch.float32
sc = scenarios.cbf_swarm(E, 8, order=3, dtype=dt)
ro = sc["rollout"]
ring = torch.empty(max(n, 24), E, 8, 20, device="cuda", dtype=dt)  # as bench.py: every step's observation goes to HBM
for _ in range(pre // 24):
    ro.run(24, obs_log=ring, log_every=1)
torch.cuda.synchronize()
ro.reset_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ro.run(n, obs_log=ring, log_every=1); e1.record()
torch.cuda.synchronize()
print("ms/control-step %.4f" % (e0.elapsed_time(e1) / n), ro.stats_dict())
ro.run(1, stages=4)
torch.cuda.synchronize()
