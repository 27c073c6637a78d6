"""Inlining tree of an ncu capture: executed warp-instructions (and stall samples) per call path, down to `depth` frames.
usage: python tools/ncu_tree.py report.ncu-rep lib.so kernel_substring [depth] [min_pct]
Frames come from nvdisasm's "inlined at" chains (tools/ncu_lines.py does the flat views)."""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, so, pat = sys.argv[1], sys.argv[2], sys.argv[3]
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 3
minp = float(sys.argv[5]) if len(sys.argv) > 5 else 0.4
objs = [so] if so.endswith(".o") else sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(so)), "build", "*.o")))
sass = []
for obj in objs:
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        sass += subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(starts[:-1], starts[1:]):
    if os.environ.get("NCU_KERNEL", "") in rows[a][1]:
        rows = rows[a:b]; break
hdr = rows[1]
body = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
base = int(body[0]["Address"], 16)
sections, cur = [], None
for ln in sass:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur = dict(name=m.group(1), ins={}, pend=[]); sections.append(cur); continue
    if cur is None: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur["pend"].append((os.path.basename(m.group(1)).replace("mds_", "").replace(".cuh", ""), int(m.group(2)))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m:
        if cur["pend"]: cur["last"], cur["pend"] = cur["pend"], []
        cur["ins"][int(m.group(1), 16)] = (cur.get("last", [("?", 0)]), m.group(2))
sec = ([s for s in sections if pat in s["name"] and len(s["ins"]) == len(body)] or [s for s in sections if pat in s["name"]])[0]
tree = collections.Counter(); samp = collections.Counter(); fp = collections.Counter(); tot = 0
for d in body:
    chain, text = sec["ins"].get(int(d["Address"], 16) - base, ([("?", 0)], ""))
    n = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0); tot += n
    t = text.split(); op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
    isfp = op in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2", "DFMA", "DMUL", "DADD")
    path = tuple(reversed(chain))[:depth]
    for k in range(1, len(path) + 1):
        tree[path[:k]] += n; samp[path[:k]] += s
        if isfp: fp[path[:k]] += n
tots = sum(int(d["# Samples"] or 0) for d in body)
print(f"{tot} warp-instructions, {tots} samples; columns: inst% samples% fp-share")
for k in sorted(tree, key=lambda k: tuple((-tree[k[:i + 1]], k[i]) for i in range(len(k)))):
    if 100 * tree[k] / tot >= minp:
        print(f"{'  ' * (len(k) - 1)}{k[-1][0]}:{k[-1][1]:<5} {100 * tree[k] / tot:6.2f}% {100 * samp[k] / max(1, tots):6.2f}%  fp {100 * fp[k] / max(1, tree[k]):3.0f}%")
