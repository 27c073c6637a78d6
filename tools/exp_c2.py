"""C2 (tracking swarm, geometric controller, N=1) under the launch plans."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
for E in (4096, 65536, 1000000):
    sc = scenarios.tracking_swarm(E)
    ro = sc["rollout"]
    ro.run(240); torch.cuda.synchronize()
    for name, st in (("fused", 3), ("two-launch", 4), ("loop", 6)):
        ro.run(240, stages=st); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ro.run(240, stages=st); ro.run(240, stages=st); e1.record(); torch.cuda.synchronize()
        print(f"C2 E={E} {name}: {e0.elapsed_time(e1) / 480 * 1e3:.2f} us per control step", flush=True)
