"""Loop kernel (plan 6) with forced block sizes.  usage: python tools/exp_threads.py E threads [threads ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
E = int(sys.argv[1])
sc = scenarios.cbf_swarm(E, 8, order=3)
ro = sc["rollout"]
ring = torch.empty(24, E, 8, 20, device="cuda")
for _ in range(1512 // 24):
    ro.run(24, obs_log=ring, log_every=1, stages=6)
for th in [int(a) for a in sys.argv[2:]]:
    os.environ["MDS_LOOP_THREADS"] = str(th)
    ro.run(24, obs_log=ring, log_every=1, stages=6)
    torch.cuda.synchronize()
    ts = []
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ro.run(24, obs_log=ring, log_every=1, stages=6)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 240 * 1e3)
    print(f"E={E} threads={th}: us/step " + " ".join(f"{t:.2f}" for t in ts) + f"   {E * 8 / min(ts) * 1e6:.3e} drone-steps/s")
