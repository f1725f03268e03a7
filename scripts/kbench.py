#!/usr/bin/env python
"""Development micro-benchmark: time cov_traj_minmax / cov_traj_fused kernel variants (COV_DEV_* switches)
on the bench.py workload at a reduced cloud size.  Not part of the product or the reported numbers."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import _lib, multicam, tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_600_000
dev = torch.device("cuda:0")
L = _lib.lib()
pts = bench.make_cloud_shard(n, 0, 1, dev)
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


def pass_a():
    _lib.check(L.cov_traj_minmax(pts.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(), ctypes.byref(cam),
                                 minmax.data_ptr(), stream), "minmax")


def pass_b():
    _lib.check(L.cov_traj_fused(pts.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(), ctypes.byref(cam),
                                minmax.data_ptr(), None, rewards.data_ptr(), acc.data_ptr(), ws.data_ptr(), wsb, stream),
               "fused")


REPS = int(os.environ.get("KBENCH_REPS", "5"))


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


import ctypes as C
stats = (C.c_ulonglong * 4)()
L.cov_set_pruning(1)
L.cov_stats(1, None)
ms_a = timeit(pass_a)
mm_p = minmax.clone()
ms_b = timeit(pass_b)
acc_p, rew_p = acc.clone(), rewards.clone()
L.cov_stats(1, stats)
print(f"pruned: pass A {ms_a:.3f} ms, pass B {ms_b:.3f} ms  -> {n * W / (ms_a + ms_b) / 1e6:.1f} G evals/s (dense-equivalent); "
      f"full-evaluated warp-iterations: pass B {stats[1] / max(stats[0], 1):.3f}, pass A {stats[3] / max(stats[2], 1):.3f}", flush=True)
L.cov_set_pruning(0)
ms_a = timeit(pass_a)
same_mm = torch.equal(minmax, mm_p)
ms_b = timeit(pass_b)
print(f"dense : pass A {ms_a:.3f} ms, pass B {ms_b:.3f} ms  -> {n * W / (ms_a + ms_b) / 1e6:.1f} G evals/s; "
      f"pruned==dense: minmax {same_mm} rewards {torch.equal(rewards, rew_p)} acc {torch.equal(acc, acc_p)}", flush=True)
if len(sys.argv) <= 2:
    sys.exit(0)
ref_mm = ref_acc = ref_rew = None
for v in range(6):
    os.environ["COV_DEV_MM"] = str(v)
    ms = timeit(pass_a)
    if ref_mm is None:
        ref_mm = minmax.clone()
    print(f"minmax variant {v}: {ms:8.3f} ms  {n * W / ms / 1e6:8.1f} G evals/s  same={torch.equal(minmax, ref_mm)}", flush=True)
os.environ["COV_DEV_MM"] = "0"
pass_a()
for v in range(4):
    os.environ["COV_DEV_F"] = str(v)
    ms = timeit(pass_b)
    if ref_acc is None:
        ref_acc, ref_rew = acc.clone(), rewards.clone()
    err = float(((acc - ref_acc).abs().max() / ref_acc.abs().max()).item())
    print(f"fused  variant {v}: {ms:8.3f} ms  {n * W / ms / 1e6:8.1f} G evals/s  rewards_same={torch.equal(rewards, ref_rew)} acc_rel={err:.2e}",
          flush=True)
