"""C5 swarm split into P independent sub-swarms, each advanced on its own CUDA stream (launches of different parts overlap, so
the tail wave of one launch is filled by the next part's blocks).  usage: python tools/exp_streams.py E P [P ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios, dist
E = int(sys.argv[1])
for P in [int(a) for a in sys.argv[2:]]:
    parts = []
    for p in range(P):
        b0, b1 = dist.env_shard(E, p, P)
        sc = scenarios.cbf_swarm(b1 - b0, 8, order=3, env_offset=b0)
        parts.append((sc, torch.empty(24, b1 - b0, 8, 20, device="cuda"), torch.cuda.Stream()))
    def launch():
        for sc, ring, st in parts:
            with torch.cuda.stream(st):
                sc["rollout"].run(24, obs_log=ring, log_every=1)
    for _ in range(1512 // 24):
        launch()
    torch.cuda.synchronize()
    ts = []
    for rep in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        cs = torch.cuda.current_stream()
        e0.record(cs)
        for sc, ring, st in parts:
            st.wait_event(e0)
        for _ in range(20):
            launch()
        for sc, ring, st in parts:
            cs.wait_stream(st)
        e1.record(cs); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 480 * 1e3)
    print(f"E={E} parts={P}: us/step " + " ".join(f"{t:.2f}" for t in ts) + f"   {E * 8 / min(ts) * 1e6:.3e} drone-steps/s")
    del parts
    torch.cuda.empty_cache()
