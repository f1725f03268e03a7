#!/usr/bin/env python
"""Profiling driver for config 5 (candidate sweep): one sweep on the ordered 50M-point cloud (run under ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import ops, tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
K, iw, ih = tools.load_intrinsics(dev)
pts, _ = ops.spatial_sort(bench.make_cloud_shard(n, 0, 1, dev))
boxes = ops.tile_boxes(pts)
P, Q = (t.to(dev) for t in bench.c5_trajectories())
for _ in range(reps):
    res = ops.sweep_rewards(pts, P, Q, K, iw, ih, boxes=boxes, presorted=True)
torch.cuda.synchronize()
print("means", float(res.min()), float(res.max()))
