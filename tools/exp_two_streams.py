"""Experiment: two half-size swarms on two streams (physics of one half overlaps the controller kernel of the other)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = 96
scs = [scenarios.cbf_swarm(E // parts, 8, order=3, env_offset=i * (E // parts)) for i in range(parts)]
streams = [torch.cuda.Stream() for _ in range(parts)]
def run(k, chunk):
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for _ in range(k // chunk):
        for sc, s in zip(scs, streams):
            with torch.cuda.stream(s):
                sc["rollout"].run(chunk)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
for chunk in (1, 4, 24):
    run(480, 24); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(K, chunk); e1.record(); torch.cuda.synchronize()
    print(f"parts {parts} chunk {chunk}: {e0.elapsed_time(e1) / K:.4f} ms per control step of {E * 8} drones")
