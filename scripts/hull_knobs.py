#!/usr/bin/env python
"""Development probe for the HPR hull stage: sweep the near-phase radius / evaluation budget after which a point is
handed to the warp-per-point kernel.  Needs a library built with -DCOV_HULL_KNOBS (COV_B200_LIB=...); the vertex set
must not depend on either knob.  usage: hull_knobs.py [n_points]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_optimization_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")


def cloud(kind):
    gen = np.random.default_rng(1)
    if kind == "shell":
        d = gen.standard_normal((n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        return (d * gen.uniform(2, 8, (n, 1))).astype(np.float32)
    return (gen.random((n, 3)) * np.array([20, 20, 4]) + np.array([-10, -10, 2])).astype(np.float32)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for kind in ("shell", "halfspace"):
    P = torch.from_numpy(cloud(kind)).to(dev)
    flipped, _ = ops.spherical_flip(P, 2)
    base = None
    for r_near, budget, r_mid in [(1, 1000, 16), (1, 1000, 16), (1, 2000, 16), (1, 700, 16), (1, 500, 16), (1, 400, 16), (1, 300, 16),
                                  (1, 200, 16), (2, 1000, 16), (2, 2000, 16), (1, 1000, 24), (6, 0, 16)]:
        os.environ["COV_HULL_R_NEAR"] = str(r_near)
        os.environ["COV_HULL_BUDGET"] = str(budget)
        os.environ["COV_HULL_R_MID"] = str(r_mid)
        ms, (mask, origin, n_unc, info) = timed(lambda: ops.hpr_hull_mask(flipped, return_info=True))
        if base is None:
            base = mask.clone()
        print(f"{kind:9s} r_near {r_near} budget {budget:5d} r_mid {r_mid:2d}: {ms:7.3f} ms  vertices {int(mask.sum())} uncertified {n_unc} "
              f"same set {bool(torch.equal(mask, base))} handed to the warp stage {info[3]}, to the block stage {info[2]}", flush=True)
