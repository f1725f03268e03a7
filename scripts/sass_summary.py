#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libcovb200.so (cuobjdump -sass; works without a GPU):
packed fp32 (FFMA2 / FMUL2 / FADD2), TMA bulk copies (UBLKCP) and their mbarriers (SYNCS), warp reductions (REDUX /
CREDUX), reductions to memory (RED / ATOM), MUFU, and the instruction count.  usage: sass_summary.py [out.csv]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "trajectory_optimization_b200", "libcovb200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.csv")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
groups = [("FFMA2", r"^FFMA2"), ("FMUL2", r"^FMUL2"), ("FADD2", r"^FADD2"), ("FFMA", r"^FFMA$|^FFMA\."), ("MUFU", r"^MUFU"),
          ("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("REDUX", r"^REDUX|^CREDUX"), ("RED", r"^RED"), ("ATOM", r"^ATOM"),
          ("SHFL", r"^SHFL"), ("VOTE", r"^VOTE"), ("BAR", r"^BAR"), ("LDS", r"^LDS"), ("LDG", r"^LDG"), ("STG", r"^STG"),
          ("DFMA", r"^DFMA|^DMUL|^DADD")]
rows, cur, k = [], None, -1
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        cur = collections.Counter()
        short = re.sub(r"\(.*", "", names[k].replace("(anonymous namespace)::", "").replace("void ", ""))
        rows.append((short, cur))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for g, pat in groups:
            if re.match(pat, op):
                cur[g] += 1
with open(out_path, "w") as f:
    f.write("kernel,total," + ",".join(g for g, _ in groups) + "\n")
    for name, c in sorted(rows, key=lambda r: r[0]):
        f.write(f"\"{name}\",{c['total']}," + ",".join(str(c[g]) for g, _ in groups) + "\n")
print("wrote", out_path, len(rows), "kernels")
