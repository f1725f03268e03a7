#!/usr/bin/env python
"""BASELINE config 2: hidden-point removal (Katz flip + hull) on a 1M-point shell cloud, one camera.
Times the CUDA stage with CUDA events and the reference arithmetic (numpy flip + scipy/Qhull) on the host."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coverage_oracle as orc  # noqa: E402
from trajectory_optimization_b200 import ops, tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "shell"
gen = np.random.default_rng(1)
if kind == "shell":
    d = gen.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * gen.uniform(2, 8, (n, 1))).astype(np.float32)
else:
    pts = (gen.random((n, 3)) * np.array([20, 20, 4]) + np.array([-10, -10, 2])).astype(np.float32)
dev = torch.device("cuda:0")
P = torch.from_numpy(pts).to(dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms_flip, (flipped, _) = timed(lambda: ops.spherical_flip(P, 2))
ms_hull, (mask, origin, n_unc) = timed(lambda: ops.hpr_hull_mask(flipped))
ms_all, (vis, vmask) = timed(lambda: tools.hidden_pts_removal(P, dev, 2))
t0 = time.perf_counter()
ref_idx, _ = orc.hidden_pts_removal(pts, 2)
cpu_s = time.perf_counter() - t0
idx = torch.nonzero(vmask).reshape(-1).cpu().numpy()
print(json.dumps({"workload": f"c2: HPR, {n} points, {kind} cloud, 1 camera", "flip_ms": ms_flip, "hull_ms": ms_hull,
                  "hidden_pts_removal_ms": ms_all, "points_per_s": n / (ms_all * 1e-3), "visible": int(len(idx)),
                  "origin_is_vertex": origin, "uncertified": n_unc, "index_set_equal_to_qhull": bool(np.array_equal(idx, ref_idx)),
                  "cpu_reference_s": cpu_s, "flip_GBps": n * 24 / (ms_flip * 1e-3) / 1e9}))
