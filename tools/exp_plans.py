"""Time the rollout launch plans (fused vs two launches per step) on the C5 swarm."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
sc = scenarios.cbf_swarm(E, 8, order=3, dtype=dt)
ro = sc["rollout"]
ro.run(480); torch.cuda.synchronize()
for rep in range(2):
    for name, st in (("fused", 3), ("two-launch", 4), ("loop", 6)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ro.run(240, stages=st); e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 240:.4f} ms per control step of {E * 8} drones", flush=True)
sys.stdout.flush()
