#!/usr/bin/env python
"""Development probe: the bench.py step (rig front end + coverage_traj + backward) eagerly and from a CUDA graph.
usage: step_probe.py [n_points] [reps]   (run under `ncu --graph-profiling node` to list the kernels of a replay)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import _lib, multicam, ops, tools  # noqa: E402
from trajectory_optimization_b200.graphs import GraphedCall  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
L = _lib.lib()
pts, perm = ops.spatial_sort(bench.make_cloud_shard(n, 0, 1, dev))
boxes = ops.tile_boxes(pts)
K, iw, ih = tools.load_intrinsics(dev)
rig7 = multicam.rig_tensor(multicam.ring_rig(5), dev)
body = bench.body_waypoints().to(dev).requires_grad_(True)
W = 320
ws = torch.empty(L.cov_traj_workspace_bytes(n, W), dtype=torch.uint8, device=dev)


def step():
    body.grad = None
    t, q = multicam.camera_poses_fused(body, rig7)
    rewards, mean = ops.coverage_traj(pts, t, q, K, iw, ih, n_total=n, reward_index=perm, boxes=boxes, workspace=ws)
    loss = 1.0 / (mean + 1e-6)
    loss.backward()
    return loss


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for _ in range(3):
    step()
print(f"eager {timeit(step):.3f} ms", flush=True)
g = GraphedCall(lambda: (step(), body.grad), warmup=2)
print(f"graph {timeit(g):.3f} ms", flush=True)
print(f"eager again {timeit(step):.3f} ms", flush=True)
