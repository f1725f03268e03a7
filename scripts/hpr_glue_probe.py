#!/usr/bin/env python
"""Development probe: where hidden_pts_removal's time goes beyond the flip and hull kernels (wall clock per statement,
synchronised).  usage: hpr_glue_probe.py [n] [shell|halfspace]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_optimization_b200 import ops, tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "halfspace"
gen = np.random.default_rng(1)
if kind == "shell":
    d = gen.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * gen.uniform(2, 8, (n, 1))).astype(np.float32)
else:
    pts = (gen.random((n, 3)) * np.array([20, 20, 4]) + np.array([-10, -10, 2])).astype(np.float32)
dev = torch.device("cuda:0")
P = torch.from_numpy(pts).to(dev)


def lap(label, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
    return out


for rep in range(4):
    acc = {}
    flipped, _ = lap("flip", lambda: ops.spherical_flip(P, 2), acc)
    mask, origin, unc = lap("hull", lambda: ops.hpr_hull_mask(flipped), acc)
    idx = lap("nonzero", lambda: torch.nonzero(mask, as_tuple=False).reshape(-1), acc)
    if not origin:
        idx = idx[:-1]
    vm = lap("zeros", lambda: torch.zeros(P.size()[0], device=dev), acc)
    lap("scatter", lambda: vm.__setitem__(idx, 1), acc)
    lap("gather", lambda: P[idx, :], acc)
    lap("whole", lambda: tools.hidden_pts_removal(P, dev, 2), acc)
    print(kind, rep, {k: round(v, 3) for k, v in acc.items()}, flush=True)
