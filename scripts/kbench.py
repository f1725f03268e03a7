#!/usr/bin/env python
"""Development micro-benchmark: times cov_traj_minmax / cov_traj_fused on the bench.py workload at a chosen cloud
size — dense, pruned on the unsorted cloud, pruned on the Morton-sorted cloud — and checks that the three give
bit-identical normalisers and rewards.  Not part of the product or the reported numbers.

usage: kbench.py [n_points]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import _lib, multicam, ops, tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_600_000
dev = torch.device("cuda:0")
L = _lib.lib()
pts_raw = bench.make_cloud_shard(n, 0, 1, dev)
K, iw, ih = tools.load_intrinsics(dev)
rig = multicam.ring_rig(5)
t, q = multicam.camera_poses_from_body(bench.body_waypoints().to(dev), rig)
P, Q = t.reshape(-1, 3).contiguous(), q.reshape(-1, 4).contiguous()
W = P.shape[0]
cam = _lib.camera(iw, ih, 1.0, 5.0, 1e-6)
minmax = torch.empty(2 * W, device=dev)
acc = torch.empty(W * _lib.ACC_STRIDE + 1, dtype=torch.float64, device=dev)
rewards = torch.empty(n, device=dev)
wsb = L.cov_traj_workspace_bytes(n, W)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
REPS = int(os.environ.get("KBENCH_REPS", "5"))

state = {"pts": pts_raw, "perm": None, "boxes": None, "dense": False}
stats_dev = torch.zeros(8, dtype=torch.int64, device=dev)


def opts():
    return ctypes.byref(_lib.traj_opts(dense=state["dense"], stats=stats_dev))


def pass_a():
    p = state["pts"]
    _lib.check(L.cov_traj_minmax(p.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(), ctypes.byref(cam),
                                 state["boxes"].data_ptr() if state["boxes"] is not None else None, minmax.data_ptr(),
                                 opts(), ws.data_ptr(), wsb, stream), "minmax")


def pass_b():
    p, perm = state["pts"], state["perm"]
    _lib.check(L.cov_traj_fused(p.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(), ctypes.byref(cam),
                                state["boxes"].data_ptr() if state["boxes"] is not None else None,
                                minmax.data_ptr(), None, None if perm is None else perm.data_ptr(), rewards.data_ptr(),
                                acc.data_ptr(), opts(), ws.data_ptr(), wsb, stream), "fused")


def timeit(fn, reps=None):
    reps = REPS if reps is None else reps
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(label):
    stats_dev.zero_()
    ms_a = timeit(pass_a)
    mm = minmax.clone()
    ms_b = timeit(pass_b)
    s = stats_dev.tolist()
    fr = lambda a, b: a / max(b, 1)  # noqa: E731
    print(f"{label:14s}: pass A {ms_a:8.3f} ms, pass B {ms_b:8.3f} ms -> {n * W / (ms_a + ms_b) / 1e6:9.1f} G evals/s "
          f"(dense-equivalent) | B: tile-listed {fr(s[6], s[0]):.4f} prefiltered {fr(s[4], s[0]):.4f} full {fr(s[1], s[0]):.4f}"
          f" | A: tile-listed {fr(s[7], s[2]):.4f} prefiltered {fr(s[5], s[2]):.4f} full {fr(s[3], s[2]):.4f}", flush=True)
    return mm, rewards.clone(), acc.clone()


state["dense"] = True
mm_d, rew_d, acc_d = run("dense")
state["dense"] = False
mm_u, rew_u, acc_u = run("pruned/unsorted")
print(f"   == dense: minmax {torch.equal(mm_u, mm_d)} rewards {torch.equal(rew_u, rew_d)} acc {torch.equal(acc_u, acc_d)} "
      f"acc_rel {float(((acc_u - acc_d).abs().max() / acc_d.abs().max()).item()):.2e}", flush=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.spatial_sort(pts_raw)
e0.record()
sorted_pts, perm = ops.spatial_sort(pts_raw)
e1.record()
torch.cuda.synchronize()
print(f"spatial sort of {n} points: {e0.elapsed_time(e1):.3f} ms; sorted == gather: "
      f"{torch.equal(sorted_pts, pts_raw[perm.long()])}; perm is a permutation: "
      f"{bool((torch.sort(perm.long()).values == torch.arange(n, device=dev)).all().item())}", flush=True)
state["pts"], state["perm"], state["boxes"] = sorted_pts, perm, ops.tile_boxes(sorted_pts)
mm_s, rew_s, acc_s = run("pruned/sorted")
print(f"   == dense: minmax {torch.equal(mm_s, mm_d)} rewards {torch.equal(rew_s, rew_d)} "
      f"acc_rel {float(((acc_s - acc_d).abs().max() / acc_d.abs().max()).item()):.2e} "
      f"sum_r_rel {abs(float(acc_s[-1] - acc_d[-1])) / float(acc_d[-1]):.2e}", flush=True)
mm_s2, rew_s2, acc_s2 = run("pruned/sorted")
print(f"   run-to-run: minmax {torch.equal(mm_s2, mm_s)} rewards {torch.equal(rew_s2, rew_s)} acc {torch.equal(acc_s2, acc_s)}", flush=True)
state["dense"] = True
mm_ds, rew_ds, acc_ds = run("dense/sorted")
print(f"   sorted pruned == sorted dense: minmax {torch.equal(mm_s, mm_ds)} rewards {torch.equal(rew_s, rew_ds)} "
      f"acc_rel {float(((acc_s - acc_ds).abs().max() / acc_ds.abs().max()).item()):.2e}", flush=True)
state["dense"] = False
