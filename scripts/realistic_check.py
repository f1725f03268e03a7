#!/usr/bin/env python
"""Pruned vs dense on a cloud with the structure of real data: the reference's sample cloud (walls, ground; 40 k points)
replicated with centimetre jitter to ~10 M points, the reference's sample path with a 5-camera rig.  Checks that the
pruned pipeline reproduces the dense rewards bit for bit and reports both step times."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_optimization_b200 import _lib, multicam, ops, tools  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
sample = np.load(os.path.join(ROOT, "tests", "golden", "sample_inputs.npz"))
gen = np.random.default_rng(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 250
base = sample["pts"].astype(np.float32)
pts_np = (np.repeat(base[None], reps, 0) + gen.normal(0, 0.03, (reps,) + base.shape).astype(np.float32)).reshape(-1, 3)
gen.shuffle(pts_np)                      # arrival order carries no spatial structure
pts = torch.from_numpy(pts_np).to(dev)
n = pts.shape[0]
wp = sample["poses"].astype(np.float32)
d = np.diff(wp, axis=0, append=wp[-1:] + (wp[-1:] - wp[-2:-1]))
body = torch.from_numpy(np.concatenate([wp, np.arctan2(d[:, 1], d[:, 0])[:, None]], 1).astype(np.float32)).to(dev)
t, q = multicam.camera_poses_fused(body, multicam.rig_tensor(multicam.ring_rig(5), dev))
K, iw, ih = tools.load_intrinsics(dev)
spts, perm = ops.spatial_sort(pts)
boxes = ops.tile_boxes(spts)


def run():
    P, Q = t.clone().requires_grad_(True), q.clone().requires_grad_(True)
    rewards, mean = ops.coverage_traj(spts, P, Q, K, iw, ih, reward_index=perm, boxes=boxes)
    gp, gq = torch.autograd.grad(mean, [P, Q])
    return rewards, mean, gp, gq


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


stats_dev = torch.zeros(8, dtype=torch.int64, device="cuda")
with ops.evaluation(stats=stats_dev):
    ms_p, (r1, m1, gp1, gq1) = timed(run)
stats = stats_dev.tolist()
with ops.evaluation(dense=True):
    ms_d, (r0, m0, gp0, gq0) = timed(run, 2)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())  # noqa: E731
print(f"{n} points x {t.shape[0]} poses (sample cloud x{reps}, jittered, shuffled): pruned step {ms_p:.3f} ms, dense step {ms_d:.3f} ms "
      f"({ms_d / ms_p:.1f}x); rewards equal {torch.equal(r1, r0)}, mean equal {torch.equal(m1, m0)}, grad rel diff "
      f"{rel(gp1, gp0):.1e} / {rel(gq1, gq0):.1e}; pass B evaluated {stats[1] / max(stats[0], 1):.4f} of the (warp, pose) pairs, "
      f"points with a gated pose {float((r0 != 0.5).float().mean()):.3f}")
