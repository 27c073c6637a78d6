import os, sys, torch
sys.path.insert(0, '/root/repo')
from multidronesim_b200 import scenarios
E = int(sys.argv[1]); plan = int(sys.argv[2])
sc = scenarios.cbf_swarm(E, 8, order=3)
ro = sc["rollout"]
ring = torch.empty(24, E, 8, 20, device="cuda")
for _ in range(20):
    ro.run(24, obs_log=ring, log_every=1, stages=6)
torch.cuda.synchronize()
ro.reset_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ro.run(24, obs_log=ring, log_every=1, stages=plan); e1.record(); torch.cuda.synchronize()
print("plan", plan, "ms/launch", e0.elapsed_time(e1), ro.stats_dict())
