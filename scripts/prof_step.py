#!/usr/bin/env python
"""Profiling driver: the two pruned trajectory passes on the Morton-sorted bench cloud, a few launches each
(run under `ncu -k regex:cov_traj --launch-skip 2 --launch-count 2`).  usage: prof_step.py [n_points] [reps]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import _lib, multicam, ops, tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
L = _lib.lib()
opts = _lib.traj_opts(dense=os.environ.get("COV_PRUNING", "1") == "0")  # COV_PRUNING=0: profile the dense kernels
pts, perm = ops.spatial_sort(bench.make_cloud_shard(n, 0, 1, dev))
boxes = ops.tile_boxes(pts)
K, iw, ih = tools.load_intrinsics(dev)
t, q = multicam.camera_poses_from_body(bench.body_waypoints().to(dev), multicam.ring_rig(5))
P, Q = t.reshape(-1, 3).contiguous(), q.reshape(-1, 4).contiguous()
W = P.shape[0]
cam = _lib.camera(iw, ih, 1.0, 5.0, 1e-6)
minmax = torch.empty(2 * W, device=dev)
acc = torch.empty(W * _lib.ACC_STRIDE + 1, dtype=torch.float64, device=dev)
rewards = torch.empty(n, device=dev)
wsb = L.cov_traj_workspace_bytes(n, W)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(reps):
    _lib.check(L.cov_traj_minmax(pts.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(), ctypes.byref(cam),
                                 boxes.data_ptr(), minmax.data_ptr(), ctypes.byref(opts), ws.data_ptr(), wsb, stream), "minmax")
    _lib.check(L.cov_traj_fused(pts.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(), ctypes.byref(cam),
                                boxes.data_ptr(), minmax.data_ptr(), None, perm.data_ptr(), rewards.data_ptr(), acc.data_ptr(),
                                ctypes.byref(opts), ws.data_ptr(), wsb, stream), "fused")
torch.cuda.synchronize()
print("mean reward", float(acc[-1]) / n)
