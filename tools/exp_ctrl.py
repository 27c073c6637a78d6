"""Per-call controller kernel (ctrl_step_kernel, MdsRolloutCfg.stages = 1) and physics kernel (stages = 2) at bench size.
usage: python tools/exp_ctrl.py [envs]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
sc = scenarios.cbf_swarm(E, 8, order=3)
ro = sc["rollout"]
for _ in range(1512 // 24):
    ro.run(24)
torch.cuda.synchronize()
n = 200
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
for k in range(n):
    ev[k][0].record(); ro.run(1, stages=1); ev[k][1].record(); ro.run(1, stages=2); ev[k][2].record()
torch.cuda.synchronize()
c = sum(e[0].elapsed_time(e[1]) for e in ev) / n
p = sum(e[1].elapsed_time(e[2]) for e in ev) / n
print(f"{os.environ.get('MDS_B200_LIB', 'default'):50s} ctrl_step {c * 1e3:.1f} us ({192 * E * 8 / c / 1e6:.0f} GB/s)   physics_step {p * 1e3:.1f} us ({232 * E * 8 / p / 1e6:.0f} GB/s)")
