#!/usr/bin/env python
"""Development probe: does keeping a reference to a tensor allocated inside a captured step change the graph?
Captures the bench step twice (with and without a retained `rewards`), times the replays and dumps the node lists."""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import _lib, multicam, ops, tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_600_000
dev = torch.device("cuda:0")
L = _lib.lib()
pts, perm = ops.spatial_sort(bench.make_cloud_shard(n, 0, 1, dev))
boxes = ops.tile_boxes(pts)
K, iw, ih = tools.load_intrinsics(dev)
rig7 = multicam.rig_tensor(multicam.ring_rig(5), dev)
body = bench.body_waypoints().to(dev).requires_grad_(True)
ws = torch.empty(L.cov_traj_workspace_bytes(n, 320), dtype=torch.uint8, device=dev)
keep = {}


def step(retain):
    body.grad = None
    t, q = multicam.camera_poses_fused(body, rig7)
    rewards, mean = ops.coverage_traj(pts, t, q, K, iw, ih, n_total=n, reward_index=perm, boxes=boxes, workspace=ws)
    loss = 1.0 / (mean + 1e-6)
    loss.backward()
    if retain:
        keep["rewards"] = rewards
    return loss


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for retain in (False, True):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step(retain)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step(retain)
    print(f"retain={retain}: replay {timeit(g.replay):.3f} ms, eager {timeit(lambda: step(retain)):.3f} ms", flush=True)
    keep.clear()
