"""Cost of materialising the observation in HBM after every control step (obs_log, log_every=1) in the default plan."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
E = 125000
sc = scenarios.cbf_swarm(E, 8, order=3)
ro = sc["rollout"]
log = torch.zeros(24, E, 8, 20, device="cuda")
for _ in range(20): ro.run(24)
torch.cuda.synchronize()
for rep in range(2):
    for name, kw in (("no log", {}), ("log every step", dict(obs_log=log, log_every=1))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ro.run(24, **kw)
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 240:.4f} ms per control step", flush=True)
