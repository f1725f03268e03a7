#!/usr/bin/env python
"""Development probe: the two pruned passes launched eagerly vs replayed from CUDA graphs (separately and together)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import _lib, multicam, ops, tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda:0")
L = _lib.lib()
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
opts = _lib.traj_opts()


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pass_a():
    _lib.check(L.cov_traj_minmax(pts.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(), ctypes.byref(cam),
                                 boxes.data_ptr(), minmax.data_ptr(), ctypes.byref(opts), ws.data_ptr(), wsb, stream()), "a")


def pass_b():
    _lib.check(L.cov_traj_fused(pts.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(), ctypes.byref(cam),
                                boxes.data_ptr(), minmax.data_ptr(), None, perm.data_ptr(), rewards.data_ptr(), acc.data_ptr(),
                                ctypes.byref(opts), ws.data_ptr(), wsb, stream()), "b")


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


def graphed(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


pass_a()
pass_b()
for name, fn in (("pass A", pass_a), ("pass B", pass_b), ("A + B", lambda: (pass_a(), pass_b()))):
    e = timeit(fn)
    g = timeit(graphed(fn))
    print(f"{name}: eager {e:.3f} ms, graph {g:.3f} ms", flush=True)
