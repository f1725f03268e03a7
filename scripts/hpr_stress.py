#!/usr/bin/env python
"""Visible-point index sets of the CUDA hidden-point removal against the reference arithmetic (numpy flip + scipy/Qhull)
on clouds of different shapes, sizes and densities (generic position: random coordinates).  Prints one line per case and
exits non-zero on the first mismatch.  usage: hpr_stress.py [seeds]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coverage_oracle as orc  # noqa: E402
from trajectory_optimization_b200 import tools  # noqa: E402

dev = torch.device("cuda:0")
seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2


def clouds(gen, n):
    d = gen.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    yield "shell", d * gen.uniform(2, 8, (n, 1))
    yield "halfspace box", gen.random((n, 3)) * np.array([20, 20, 4]) + np.array([-10, -10, 2])
    yield "gaussian blob off centre", gen.standard_normal((n, 3)) * np.array([1.0, 2.0, 0.5]) + np.array([6.0, 0.0, 1.0])
    yield "two clusters", np.concatenate([gen.standard_normal((n // 2, 3)) * 0.7 + np.array([4.0, 3.0, 0.0]),
                                          gen.standard_normal((n - n // 2, 3)) * 1.5 + np.array([-5.0, -2.0, 1.0])])
    yield "thick wall", gen.random((n, 3)) * np.array([0.3, 30, 10]) + np.array([5.0, -15, -5])
    yield "camera inside a box", (gen.random((n, 3)) - 0.5) * np.array([12, 9, 5])
    yield "ground plane with noise", np.stack([gen.uniform(-20, 20, n), gen.uniform(-20, 20, n), -1.5 + 0.05 * gen.standard_normal(n)], 1)


bad = 0
for seed in range(seeds):
    for n in (50, 1000, 20_000, 200_000):
        gen = np.random.default_rng(100 * seed + n)
        for name, pts in clouds(gen, n):
            pts = pts.astype(np.float32)
            ref_idx, _ = orc.hidden_pts_removal(pts, 2)
            vis, mask = tools.hidden_pts_removal(torch.from_numpy(pts).to(dev), dev, 2)
            idx = torch.nonzero(mask).reshape(-1).cpu().numpy()
            same = np.array_equal(idx, ref_idx)
            bad += not same
            print(f"seed {seed} n {n:7d} {name:26s} visible {len(ref_idx):6d} {'equal' if same else 'DIFFERENT: ' + str(len(np.setxor1d(idx, ref_idx)))}",
                  flush=True)
sys.exit(1 if bad else 0)
