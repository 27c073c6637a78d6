"""Turn the raw artefacts of tools/gpu_measure.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
usage: python tools/summarise_profiles.py <tag>"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
so = os.path.join(ROOT, "multidronesim_b200", "csrc", "libmds_b200.so")
lines = []

# ---- bench lines
for name in (f"{tag}_bench.json", f"{tag}_bench_reference.json"):
    src = os.path.join(G, name)
    if os.path.isfile(src) and os.path.getsize(src):
        shutil.copy(src, os.path.join(P, name))
b = json.load(open(os.path.join(P, f"{tag}_bench.json")))
lines += [f"# {tag}: measured on 1x B200 (python bench.py, defaults)", "",
          f"value {b['value']:.4g} {b['unit']}  ({b['ms_per_step'] / b['config']['control_steps_per_step']:.4f} ms per control step of "
          f"{b['config']['envs_per_gpu'] * b['config']['drones_per_env']} drones)",
          f"e2e   {b['e2e']['value']:.4g} {b['unit']}  ({b['e2e']['ms_per_control_step']:.3f} ms per control step; H2D {b['e2e']['h2d_bytes_per_step'] / 1e6:.0f} MB, D2H {b['e2e']['d2h_bytes_per_step'] / 1e6:.0f} MB per step)",
          f"cpu   {b['cpu_baseline']['value']:.4g} {b['unit']} on {b['cpu_baseline']['cores']} cores ({b['cpu_baseline']['kind']})" if b.get("cpu_baseline") else "cpu   -",
          f"clocks {b['clocks']}", ""]
for key in ("roofline", "roofline_ctrl", "roofline_physics"):
    r = b[key]
    lines.append(f"{key}: {r['kernel']}: {r['achieved']:.4g} {r['unit']} of {r['peak']:.4g} = {r['frac']:.3f} ({r['bound']}); "
                 f"launch {r['launch_ms'] * 1e3:.1f} us, ncu DRAM traffic per launch {r['traffic']}")
lines.append(f"rollout_stats: {b['rollout_stats']}")
lines.append("")

# ---- launch list
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.isfile(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        us = v / 1000 if r[ui].startswith("n") else (v if r[ui].startswith("u") else v * 1000)
        k = r[ki].split("(")[0].replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    ll = [f"ncu launch list of `python bench.py --no-cpu-baseline` ({sum(v[0] for v in agg.values())} launches; cold-cache, serialised: compare shares)"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ll.append(f"  {k[:72]:<72} n={v[0]:6d} total {v[1] / 1e3:10.2f} ms  mean {v[1] / v[0]:8.1f} us  share {100 * v[1] / tot:5.1f}%")
    loop = [v for k, v in agg.items() if k.startswith("rollout_loop_kernel")]
    if loop:
        ll.append(f"  timed region of bench.py = {b['steps']} consecutive rollout_loop_kernel launches and nothing else: share of the step 1.00 "
                  f"(bench roofline.share_of_step {b['roofline']['share_of_step']:.2f}); ncu mean {loop[0][1] / loop[0][0]:.0f} us per launch (cold, serialised) vs "
                  f"{b['roofline']['launch_ms'] * 1e3:.0f} us from the bench's CUDA events; the other launches are settle / warm-up, the per-kernel "
                  "replay (ctrl_step, physics_step), the e2e per-call pipeline and the FMA-peak microbenchmark")
    c = sum(v[1] for k, v in agg.items() if k.startswith("ctrl_step_kernel"))
    p = sum(v[1] for k, v in agg.items() if k.startswith("physics_step_kernel"))
    ll.append(f"  ctrl_step : physics_step total time ratio under ncu = {c / max(p, 1e-9):.2f} (bench replay events: "
              f"{b['roofline_ctrl']['launch_ms'] / b['roofline_physics']['launch_ms']:.2f})")
    open(os.path.join(P, f"{tag}_launches_summary.txt"), "w").write("\n".join(ll) + "\n")
    lines += ll + [""]

# ---- ncu full capture
rep = os.path.join(G, f"{tag}_prof.ncu-rep")
if os.path.isfile(rep):
    raw = os.path.join(G, f"{tag}_prof_raw.csv")
    open(raw, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_ncu_kernels.txt"), "w").write(summ)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), rep, "125000", "f32"], check=True)
    for kern, pat, fn in (("rollout_loop", "rollout_loop_kernelIfLi3ELb1ELi8", "loop_lines"), ("ctrl_step", "ctrl_step_kernelIfLi3ELb1ELi8", "ctrl_lines"),
                          ("physics_step", "physics_step_kernelIfLi8", "physics_lines")):
        env = dict(os.environ, NCU_KERNEL=kern)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, so, pat, "40"], capture_output=True, text=True, env=env).stdout
        open(os.path.join(P, f"{tag}_{fn}.txt"), "w").write(out)
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
