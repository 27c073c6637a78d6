import sys, torch, time, json
sys.path.insert(0, '/root/repo')
import multidronesim_b200 as mds
from multidronesim_b200 import _lib, scenarios
E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
sc = scenarios.cbf_swarm(E, 8, order=3)
env, ro, ctrl = sc["env"], sc["rollout"], sc["ctrl"]
ro.run(96); torch.cuda.synchronize()
snap = (env.state_dict(), ctrl.low_level._a.clone(), ctrl.low_level._b.clone(), ro.t)
def restore():
    env.load_state_dict(snap[0]); ctrl.low_level._a.copy_(snap[1]); ctrl.low_level._b.copy_(snap[2]); ro.t = snap[3]
def timeit(f, K, reps=3):
    ts = []
    for _ in range(reps + 1):
        restore(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(K); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / K)
    return min(ts[1:])
def run_nostats(K):
    s = ro.stats; ro.stats = None
    try: ro.run(K)
    finally: ro.stats = s
print("48 steps from snapshot: with stats %.4f ms/step, without %.4f ms/step" % (timeit(ro.run, 48), timeit(run_nostats, 48)))
# activity profile over one lemniscate period
restore(); ro.reset_stats(); torch.cuda.synchronize()
prof = []
for chunk in range(0, 3024, 48):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0 = ro.stats.clone()
    e0.record(); ro.run(48); e1.record(); torch.cuda.synchronize()
    d = (ro.stats - s0).tolist()
    prof.append((round(e0.elapsed_time(e1) / 48, 3), round(d[4] / (E * 48), 3), round(d[5] / max(1, d[4]), 2), round(d[6] / (E * 48), 3)))
print("per-48-step chunks (ms/step, qp_active_frac, iters/solve, infeasible_frac):")
print(prof)
print("mean ms/step over the period:", sum(p[0] for p in prof) / len(prof))
print(json.dumps(ro.stats_dict()))
