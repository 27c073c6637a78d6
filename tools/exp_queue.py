"""Launch plan 6 (block-scheduled loop kernel) vs 7 (work-queue kernel) on the C5 swarm at several shard sizes.
usage: python tools/exp_queue.py [envs ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
sizes = [int(a) for a in sys.argv[1:]] or [15625, 31250, 62500, 125000]
for E in sizes:
    res = {}
    for plan in (6, 7):
        sc = scenarios.cbf_swarm(E, 8, order=3)
        ro = sc["rollout"]
        ring = torch.empty(24, E, 8, 20, device="cuda")
        for _ in range(1512 // 24):
            ro.run(24, obs_log=ring, log_every=1, stages=plan)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            ro.run(24, obs_log=ring, log_every=1, stages=plan)
        e1.record(); torch.cuda.synchronize()
        res[plan] = e0.elapsed_time(e1) / (40 * 24)
        del sc, ro, ring
    print(f"E={E:7d}  plan6 {res[6] * 1e3:8.2f} us/step  {E * 8 / res[6] * 1e3:.3e} /s   plan7 {res[7] * 1e3:8.2f} us/step  {E * 8 / res[7] * 1e3:.3e} /s   ratio {res[6] / res[7]:.3f}")
