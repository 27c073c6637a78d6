"""Join an ncu SASS source page with nvdisasm line info: instructions executed and stall samples per source line.
usage: python tools/ncu_lines.py report.ncu-rep lib.so kernel_substring [top_n]
Prints (a) totals per OUTERMOST line (statement of the __global__ function) and (b) per innermost file:line."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, so, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# The library holds one cubin per translation unit and cuobjdump names them after the SOURCE file (six of them are
# mds_rollout_tu.sm_100a.cubin), so extracting from the .so overwrites them: disassemble the objects next to it instead.
import glob
objs = [so] if so.endswith(".o") else sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(so)), "build", "*.o")))
sass = []
for obj in objs:
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        sass += subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
want_k = os.environ.get("NCU_KERNEL", "")
for a, b in zip(starts[:-1], starts[1:]):  # a report may hold several kernels: take the first whose name matches
    if want_k in rows[a][1]:
        rows = rows[a:b]
        break
kname = rows[0][1]
hdr = rows[1]
body = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
base = int(body[0]["Address"], 16)
# mangled-name match: take the section whose instruction count equals the report's
sections, cur = [], None
for ln in sass:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur = dict(name=m.group(1), ins={}, pend=[])
        sections.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur["pend"].append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m:
        if cur["pend"]:
            cur["last"] = cur["pend"]
            cur["pend"] = []
        cur["ins"][int(m.group(1), 16)] = cur.get("last", [("?", 0)])
cands = [s for s in sections if pat in s["name"] and len(s["ins"]) == len(body)]
if not cands:
    cands = [s for s in sections if pat in s["name"]]
    print("warning: no section with matching instruction count; candidates:", [(s["name"][:60], len(s["ins"])) for s in cands], "report has", len(body))
sec = cands[0]
inner, outer = collections.Counter(), collections.Counter()
s_inner, s_outer = collections.Counter(), collections.Counter()
tot_i = tot_s = 0
for d in body:
    off = int(d["Address"], 16) - base
    chain = sec["ins"].get(off, [("?", 0)])
    n = int(d["Instructions Executed"] or 0)
    s = int(d["# Samples"] or 0)
    inner[chain[0]] += n; s_inner[chain[0]] += s
    outer[chain[-1]] += n; s_outer[chain[-1]] += s
    tot_i += n; tot_s += s
print(f"kernel {kname[:80]}\ninstructions {len(body)} static, {tot_i} warp-level executed, {tot_s} samples")
print("-- per outermost (kernel-level) line: inst%  samples%")
for k, v in outer.most_common(top):
    print(f"  {k[0]}:{k[1]:<5} {100 * v / tot_i:6.2f}%  {100 * s_outer[k] / max(1, tot_s):6.2f}%")
print("-- per innermost line")
for k, v in inner.most_common(top):
    print(f"  {k[0]}:{k[1]:<5} {100 * v / tot_i:6.2f}%  {100 * s_inner[k] / max(1, tot_s):6.2f}%")
if len(sys.argv) > 5:  # drill into one outermost line: attribution one inlining level further in
    want = int(sys.argv[5])
    lvl, s_lvl = collections.Counter(), collections.Counter()
    for d in body:
        chain = sec["ins"].get(int(d["Address"], 16) - base, [("?", 0)])
        if chain[-1][1] != want:
            continue
        k = chain[-2] if len(chain) > 1 else chain[-1]
        lvl[k] += int(d["Instructions Executed"] or 0); s_lvl[k] += int(d["# Samples"] or 0)
    print(f"-- inside line {want}, one level in: inst%  samples%")
    for k, v in lvl.most_common(top):
        print(f"  {k[0]}:{k[1]:<5} {100 * v / tot_i:6.2f}%  {100 * s_lvl[k] / max(1, tot_s):6.2f}%")
