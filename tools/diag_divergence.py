"""Diagnostic: how the C5 closed loop behaves over time (per-env tracking error distribution, non-finite states)."""
import sys, json, torch
sys.path.insert(0, '/root/repo')
from multidronesim_b200 import scenarios
E = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1440
sc = scenarios.cbf_swarm(E, 8, order=3, dtype=dt)
env, ro, trajs = sc["env"], sc["rollout"], sc["trajs"]
out = []
for chunk in range(steps // 48):
    ro.run(48)
    ref = trajs.eval(ro.t).reshape(E, 8, 11)[..., 0:3]
    obs = env.obs
    err = (obs[..., 0:3] - ref).norm(dim=-1)
    bad = ~torch.isfinite(obs).all(dim=-1)
    emax = torch.nan_to_num(err, nan=1e30).amax(dim=1)
    row = dict(t=round(ro.t, 3), nonfinite=int(bad.sum()), gt1=int((emax > 1).sum()), gt10=int((emax > 10).sum()), gt1e3=int((emax > 1e3).sum()),
               max=float(emax.max()), med=float(emax.median()), stats=ro.stats_dict())
    out.append(row)
    if row["gt1e3"] and not any(r.get("dump") for r in out):
        e = int(torch.argmax(emax))
        row["dump"] = dict(env=e, obs=obs[e].double().cpu().tolist(), init=sc["init"][e].tolist())
    print(json.dumps({k: v for k, v in row.items() if k not in ("dump",)}), flush=True)
json.dump(out, open("gpurun_out/diag_divergence_%s.json" % ("f64" if dt == torch.float64 else "f32"), "w"))
