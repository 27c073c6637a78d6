"""Work-queue rollout (plan 7) with forced chunk counts.  usage: python tools/exp_chunks.py E chunks [chunks ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
E = int(sys.argv[1])
sc = scenarios.cbf_swarm(E, 8, order=3)
ro = sc["rollout"]
ring = torch.empty(24, E, 8, 20, device="cuda")
for _ in range(1512 // 24):
    ro.run(24, obs_log=ring, log_every=1, stages=6)
for ch in [0] + [int(a) for a in sys.argv[2:]]:
    if ch:
        os.environ["MDS_QUEUE_CHUNKS"] = str(ch)
    plan = 7 if ch else 6
    ro.run(24, obs_log=ring, log_every=1, stages=plan)
    torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ro.run(24, obs_log=ring, log_every=1, stages=plan)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 240 * 1e3)
    print(f"E={E} chunks={ch if ch else 'plan6'}: us/step " + " ".join(f"{t:.2f}" for t in ts))
