"""Static SASS census of one kernel (no GPU needed): instruction count, opcode mix and instructions per source file / line,
from `nvdisasm -gi` of the cubin inside an object or shared library built with -lineinfo.
usage: python tools/sass_static.py <file.o|.so> <kernel-name-substring> [top_n]
Static counts include cold paths (slow-path solvers, table walks); use it to compare variants of straight-line stages."""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    sass = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    cur, secs = None, []
    for ln in sass:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur = dict(name=m.group(1), ins=[], pend=[], last=[("?", 0)])
            secs.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur["pend"].append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m:
            if cur["pend"]:
                cur["last"], cur["pend"] = cur["pend"], []
            cur["ins"].append((m.group(2), cur["last"]))
    for s in secs:
        if pat not in s["name"] or not s["ins"]:
            continue
        ops, files, lines = collections.Counter(), collections.Counter(), collections.Counter()
        for text, chain in s["ins"]:
            t = text.split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += 1
            files[chain[0][0]] += 1   # innermost frame
            lines[chain[0]] += 1
        n = len(s["ins"])
        fp = sum(ops[k] for k in ("FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD"))
        print(f"== {s['name'][:110]}\n   {n} instructions, FP {fp} ({100 * fp / n:.0f} %): " +
              ", ".join(f"{k} {v}" for k, v in ops.most_common(22)))
        print("   per file: " + ", ".join(f"{k} {v}" for k, v in files.most_common()))
        print("   top lines: " + ", ".join(f"{k[0]}:{k[1]} {v}" for k, v in lines.most_common(top)))
