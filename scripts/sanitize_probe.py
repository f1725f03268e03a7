#!/usr/bin/env python
"""Small self-checking calls of the kernels added in the second half of round 2 (pair sort, scan, Morton ordering, voxel
filter, the staged hull): quick to run on their own, and small enough for compute-sanitizer (memcheck / racecheck /
synccheck) where a pool allows it (this round's pool does not).  usage: sanitize_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_optimization_b200 import ops, tools  # noqa: E402

dev = torch.device("cuda:0")
gen = np.random.default_rng(0)
for n in (100, 4096, 20_001):
    keys = gen.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    k, v = ops.sort_pairs(torch.from_numpy(keys.view(np.int32)).to(dev), torch.arange(n, dtype=torch.int32, device=dev))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(v.cpu().numpy(), order.astype(np.int32)), n
pts = (gen.random((30_011, 3)) * np.array([40, 40, 5])).astype(np.float32)
out, perm = ops.spatial_sort(torch.from_numpy(pts).to(dev))
assert torch.equal(out, torch.from_numpy(pts).to(dev)[perm.long()])
vox = tools.voxel_grid_filter(torch.from_numpy(pts - np.array([20, 20, 2.5], np.float32)).to(dev), 0.5, "z", -2.5, 2.5)
print("voxels", tuple(vox.shape))
for kind in ("halfspace", "shell"):
    n = 6000
    if kind == "shell":
        d = gen.standard_normal((n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = (d * gen.uniform(2, 8, (n, 1))).astype(np.float32)
    else:
        p = (gen.random((n, 3)) * np.array([20, 20, 4]) + np.array([-10, -10, 2])).astype(np.float32)
    vis, mask = tools.hidden_pts_removal(torch.from_numpy(p).to(dev), dev, 2)
    print(kind, "visible", int(mask.sum()))
torch.cuda.synchronize()
print("done")
