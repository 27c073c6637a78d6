"""Soak: the 1M-drone C5 swarm for many laps; per lap the error statistics, QP outcome counts, quaternion norm drift and
non-finite count.  usage: python tools/soak.py [laps]"""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multidronesim_b200 import scenarios
laps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
E = 125000
sc = scenarios.cbf_swarm(E, 8, order=3)
env, ro = sc["env"], sc["rollout"]
for lap in range(laps):
    ro.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(126):
        ro.run(24)  # 3024 steps = one lemniscate period
    e1.record(); torch.cuda.synchronize()
    st = ro.stats_dict()
    qn = env.quat.norm(dim=-1)
    print(json.dumps({"lap": lap + 1, "t": round(ro.t, 2), "ms_per_step": round(e0.elapsed_time(e1) / 3024, 4), "mean_err": st["sum_pos_err"] / st["drone_steps"],
                      "max_err": st["max_pos_err"], "min_h": st["min_barrier"], "it_per_solve": st["qp_iters"] / max(1, st["qp_solves"]),
                      "infeasible": st["qp_infeasible"], "cap": st["qp_iter_cap"], "quat_norm_min": float(qn.min()), "quat_norm_max": float(qn.max()),
                      "nonfinite": int((~torch.isfinite(env.obs)).sum())}), flush=True)
