"""ms per control step and QP activity per 48-step chunk over one lap of the C5 swarm (default launch plan)."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
sc = scenarios.cbf_swarm(E, 8, order=3)
ro = sc["rollout"]
ro.run(24); torch.cuda.synchronize()
rows = []
for chunk in range(3024 // 48):
    s0 = ro.stats.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ro.run(24); ro.run(24); e1.record(); torch.cuda.synchronize()
    d = (ro.stats - s0).tolist()
    rows.append((round(ro.t, 2), round(e0.elapsed_time(e1) / 48, 4), round(d[4] / (E * 48), 3), round(d[5] / max(1, d[4]), 2), round(d[6] / (E * 48), 4), round(d[7] / (E * 48), 5)))
print("t, ms/step, qp_active_frac, iters/solve, infeasible_frac, cap_frac")
for r in rows:
    print(r)
print("mean ms/step", sum(r[1] for r in rows) / len(rows))
