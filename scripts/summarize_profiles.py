#!/usr/bin/env python
"""Condense gpurun_out/ ncu artefacts into small tracked files under profiles/ (run in the build container).
usage: summarize_profiles.py <tag> [launches.csv] [prof.ncu-rep]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launch_csv = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
rep = sys.argv[3] if len(sys.argv) > 3 else None
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

if os.path.exists(launch_csv):
    rows = list(csv.reader(open(launch_csv)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = [(int(r[idc]), r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, k, v in launches:
        name = k.split("(")[0].replace("void ", "").strip()
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches_by_kernel.csv"), "w") as f:
        f.write("kernel,launches,total_ns,share_of_all_profiled_launches\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{name}\",{c},{t:.0f},{t / total:.4f}\n")
    ours = [(i, k.split("(")[0].replace("void ", "").strip(), v) for i, k, v in launches if "cov_" in k or "hull_" in k]
    with open(os.path.join(out_dir, f"{tag}_launch_list_own_kernels.csv"), "w") as f:
        f.write("id,kernel,gpu__time_duration_ns\n")
        for i, k, v in ours:
            f.write(f"{i},\"{k}\",{v:.0f}\n")
    print("launches:", len(launches), "own:", len(ours))

if rep and os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
            "smsp__warps_eligible.avg.per_cycle_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
    out = []
    for r in rows[2:]:
        d = {}
        for k in want:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
                except ValueError:
                    pass
        d["warp_stalls_per_issue"] = {k: round(v, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        out.append(d)
    json.dump(out, open(os.path.join(out_dir, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
    print("ncu kernels:", len(out))
