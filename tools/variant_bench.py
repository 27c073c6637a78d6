"""Time the C5 rollout (bench workload, obs log every step) for the library selected by MDS_B200_LIB.
usage: python tools/variant_bench.py [envs] [settle_steps] [timed_launches]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 3024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sc = scenarios.cbf_swarm(E, 8, order=3)
ro = sc["rollout"]
ring = torch.empty(24, E, 8, 20, device="cuda", dtype=torch.float32)
for _ in range(settle // 24):
    ro.run(24, obs_log=ring, log_every=1)
torch.cuda.synchronize()
ro.reset_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ro.run(24, obs_log=ring, log_every=1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (reps * 24)
st = ro.stats_dict()
print(f"{os.environ.get('MDS_B200_LIB', 'default'):40s} ms/step {ms:.4f}  drone-steps/s {E * 8 / ms * 1e3:.4e}  iters/solve {st['qp_iters'] / max(1, st['qp_solves']):.3f} "
      f"infeasible {st['qp_infeasible']:.0f} cap {st['qp_iter_cap']:.0f} max_err {st['max_pos_err']:.4f}")
