#!/usr/bin/env python
"""Secondary measurements for the other BASELINE configs (one GPU), written as JSON lines:
  c1  single-camera pose optimisation, 100k points: opt-step latency (zero_grad + fwd + bwd + Adam), plus the
      ModelPose kernel at 1e8 points against the HBM roofline;
  c3  5-camera trajectory evaluation, 20 waypoints, 10M points (fwd+bwd evals/s);
  c5  candidate sweep sample: 64 of the 1024 trajectories x 32 waypoints on a 50M-point cloud (fwd evals/s).
c2 (HPR) is scripts/hpr_bench.py.  Every GPU figure is followed by the same work through oracle/torch_port.py (the
torch port of the reference) on the host cores AND on the GPU in eager mode (what a user of the reference gets today
on this machine).  Not the driver's bench; numbers are quoted in DESIGN.md/profiles."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import model, multicam, ops, tools  # noqa: E402

dev = torch.device("cuda:0")
K, iw, ih = tools.load_intrinsics(dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6650.0))


def events(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def box(n, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    lo, hi = torch.tensor(bench.BOX_LO, device=dev), torch.tensor(bench.BOX_HI, device=dev)
    return torch.rand(n, 3, generator=g, device=dev) * (hi - lo) + lo


# ---- c1: pose optimisation step latency at 100k points ----
pts = box(100_000, 0)
m = model.ModelPose(pts, torch.tensor([[6.0, 2.0, 0.0]]), torch.tensor([[0.92, 0.0, 0.0, 0.39]]), K, iw, ih, device=dev)
opt = torch.optim.Adam([{"params": [m.trans], "lr": 0.02}, {"params": [m.quat], "lr": 0.02}])


def pose_step():
    opt.zero_grad()
    loss = m()
    loss.backward()
    opt.step()


ms = events(pose_step, 200)
t0 = time.perf_counter()
for _ in range(200):
    pose_step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 200 * 1e3
print(json.dumps({"config": "c1: ModelPose opt step (zero_grad+fwd+bwd+Adam), 100k points", "gpu_ms_per_step": ms,
                  "wall_ms_per_step": wall, "evals_per_s": 1e5 / (wall * 1e-3)}), flush=True)
from trajectory_optimization_b200.graphs import GraphedStep  # noqa: E402
mg = model.ModelPose(pts, torch.tensor([[6.0, 2.0, 0.0]]), torch.tensor([[0.92, 0.0, 0.0, 0.39]]), K, iw, ih, device=dev)
optg = torch.optim.Adam([{"params": [mg.trans], "lr": 0.02}, {"params": [mg.quat], "lr": 0.02}], capturable=True)
gstep = GraphedStep(mg, optg)
ms = events(gstep.step, 500)
t0 = time.perf_counter()
for _ in range(500):
    gstep.step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 500 * 1e3
print(json.dumps({"config": "c1: ModelPose opt step as one CUDA graph (graphs.GraphedStep), 100k points", "gpu_ms_per_step": ms,
                  "wall_ms_per_step": wall}), flush=True)
# the reference's trajectory optimisation step (src/trajectory_optimization.py:106-116; its comment: "~125 msec") on the
# sample inputs the reference ships (40 k points, 27 waypoints), eager and as one CUDA graph
sample = np.load(os.path.join(ROOT, "tests", "golden", "sample_inputs.npz"))
spts, sposes = torch.from_numpy(sample["pts"]).float(), torch.from_numpy(sample["poses"]).float()
squats = torch.tensor([[1.0, 0.0, 0.0, 0.0]]).repeat(sposes.shape[0], 1)
for graphed, fused in ((False, False), (True, False), (False, True), (True, True)):
    mt = model.ModelTraj(spts, sposes, squats, K, iw, ih, device=dev, fused_regularizers=fused)
    optt = torch.optim.Adam([{"params": [mt.poses], "lr": 0.1}, {"params": [mt.quats], "lr": 0.02}], capturable=graphed)
    if graphed:
        gs = GraphedStep(mt, optt)
        fn = gs.step
    else:
        def fn():
            optt.zero_grad()
            loss = mt()
            loss.backward()
            optt.step()
    ms = events(fn, 200)
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 200 * 1e3
    print(json.dumps({"config": "ModelTraj opt step (zero_grad+fwd+bwd+Adam) on the reference's sample cloud and path, "
                                + ("one CUDA graph" if graphed else "eager") + (", fused regularisers" if fused else ""),
                      "gpu_ms_per_step": ms, "wall_ms_per_step": wall}), flush=True)
# ModelPose kernel alone at 1e8 points vs HBM
big = box(100_000_000, 1)
T = torch.tensor([[6.0, 2.0, 0.0]], device=dev, requires_grad=True)
Q = torch.tensor([[0.92, 0.0, 0.0, 0.39]], device=dev, requires_grad=True)
ms = events(lambda: ops.coverage_pose(big, T, Q, K, iw, ih), 10)
print(json.dumps({"config": "ModelPose fused fwd+bwd, 1e8 points", "ms": ms, "evals_per_s": 1e8 / (ms * 1e-3),
                  "hbm_GBps_algorithmic_16B_per_point": 1.6e9 / (ms * 1e-3) / 1e9, "hbm_frac_of_measured_peak": 1.6 / (ms * 1e-3) / HBM}),
      flush=True)
del big

# ---- c3: 20 waypoints x 5 cams, 10M points, fwd+bwd ----
pts = box(10_000_000, 2)
rig = multicam.ring_rig(5)
body = bench.body_waypoints(20, 12.0).to(dev).requires_grad_(True)


rig7 = multicam.rig_tensor(rig, dev)
spts3, perm3 = ops.spatial_sort(pts)
boxes3 = ops.tile_boxes(spts3)


def traj_step():
    body.grad = None
    t, q = multicam.camera_poses_fused(body, rig7)
    rewards, mean = ops.coverage_traj(spts3, t, q, K, iw, ih, reward_index=perm3, boxes=boxes3)
    (1.0 / (mean + 1e-6)).backward()


ms = events(traj_step, 20)
print(json.dumps({"config": "c3: 20 waypoints x 5 cams, 10M points, fwd+bwd", "ms_per_step": ms,
                  "evals_per_s": 1e7 * 100 / (ms * 1e-3)}), flush=True)

# ---- c5: 1024 trajectories x 32 waypoints (single camera per waypoint), 50M points, forward only ----
pts = box(50_000_000, 3)
gen = np.random.default_rng(2)
base = bench.body_waypoints(32, 20.0).numpy()
Tn = 1024
poses = np.repeat(base[None, :, :3], Tn, 0) + gen.normal(0, 1.0, (Tn, 1, 3)) * np.array([1, 1, 0])
yaw = base[None, :, 3] + gen.normal(0, 0.2, (Tn, 32))
quats = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], -1)
P, Qs = torch.tensor(poses, dtype=torch.float32, device=dev), torch.tensor(quats, dtype=torch.float32, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
spts, _ = ops.spatial_sort(pts)
sboxes = ops.tile_boxes(spts)
torch.cuda.synchronize()
ms_order = (time.perf_counter() - t0) * 1e3
ms = events(lambda: ops.sweep_rewards(spts, P, Qs, K, iw, ih, boxes=sboxes, presorted=True), 3)
pairs = 5e7 * Tn * 32
res = ops.sweep_rewards(spts, P, Qs, K, iw, ih, boxes=sboxes, presorted=True)
print(json.dumps({"config": "c5: 1024 trajectories x 32 waypoints, 50M points, forward only (2 passes), Morton-ordered cloud",
                  "ms": ms, "order_cloud_ms": ms_order, "pairs_per_s": pairs / (ms * 1e-3),
                  "mean_reward_range": [float(res.min()), float(res.max())]}), flush=True)
# dense reference on a sample of 32 trajectories (the full dense sweep takes ~8 s)
from trajectory_optimization_b200 import _lib  # noqa: E402
with ops.evaluation(dense=True):
    ms_d = events(lambda: ops.sweep_rewards(spts, P[:32], Qs[:32], K, iw, ih, boxes=sboxes, presorted=True), 1)
    res_d = ops.sweep_rewards(spts, P[:32], Qs[:32], K, iw, ih, boxes=sboxes, presorted=True)
print(json.dumps({"config": "c5 dense sample: 32 of the 1024 trajectories, pruning off", "ms": ms_d,
                  "pairs_per_s": 5e7 * 32 * 32 / (ms_d * 1e-3), "extrapolated_full_c5_s": ms_d * 1e-3 * 32,
                  "max_rel_diff_vs_pruned": float(((res[:32] - res_d).abs() / res_d.abs()).max())}), flush=True)


# ---- the reference's own arithmetic (oracle/torch_port.py) on the host cores and on the GPU in eager mode ----
from oracle import torch_port  # noqa: E402

threads = os.cpu_count() or 1
torch.set_num_threads(threads)


def port_pose_step(points, device, reps):
    T = torch.tensor([[6.0, 2.0, 0.0]], device=device, requires_grad=True)
    Qp = torch.tensor([[0.92, 0.0, 0.0, 0.39]], device=device, requires_grad=True)
    o = torch.optim.Adam([{"params": [T], "lr": 0.02}, {"params": [Qp], "lr": 0.02}])
    Kd = K.to(device)

    def one():
        o.zero_grad()
        loss, _ = torch_port.pose_loss(points, T, Qp, Kd, iw, ih)
        loss.backward()
        o.step()

    for _ in range(2):
        one()
    if device.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        one()
    if device.type == "cuda":
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def port_traj_step(points, P, Q, device, reps):
    Kd = K.to(device)
    P, Q, points = P.to(device), Q.to(device), points.to(device)
    for _ in range(1):
        torch_port.traj_step(points, P, Q, Kd, iw, ih)
    if device.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        torch_port.traj_step(points, P, Q, Kd, iw, ih)
    if device.type == "cuda":
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


cpu = torch.device("cpu")
pts1 = box(100_000, 0)
ms_cpu = port_pose_step(pts1.cpu(), cpu, 10)
ms_gpu = port_pose_step(pts1, dev, 50)
print(json.dumps({"config": "c1 reference arithmetic (oracle/torch_port.py): ModelPose opt step, 100k points",
                  "cpu_ms_per_step": ms_cpu, "cpu_threads": threads, "torch_cuda_eager_ms_per_step": ms_gpu}), flush=True)
sel = slice(0, None, 2)   # wps_step = 2 on the sample path (src/model.py:214-217)
sample_pts = torch.from_numpy(sample["pts"]).float()   # (`spts` was reused for the c5 cloud above)
ms_cpu = port_traj_step(sample_pts, sposes[sel], squats[sel], cpu, 3)
ms_gpu = port_traj_step(sample_pts, sposes[sel], squats[sel], dev, 10)
print(json.dumps({"config": "reference arithmetic: trajectory visibility fwd+bwd on the reference's sample cloud (40k points, 14 of 27 "
                            "waypoints)", "cpu_ms_per_step": ms_cpu, "cpu_threads": threads,
                  "torch_cuda_eager_ms_per_step": ms_gpu}), flush=True)
with torch.no_grad():
    t3, q3 = multicam.camera_poses_from_body(bench.body_waypoints(20, 12.0), multicam.ring_rig(5))
P3, Q3 = t3.reshape(-1, 3).contiguous(), q3.reshape(-1, 4).contiguous()
n3 = 100_000   # torch autograd keeps ~240 B per (point, pose): 1e7 x 100 would need 240 GB
pts3 = box(n3, 2)
ms_cpu = port_traj_step(pts3.cpu(), P3, Q3, cpu, 2)
ms_gpu = port_traj_step(pts3, P3, Q3, dev, 5)
print(json.dumps({"config": "c3 reference arithmetic on a 100k-point sample (100 poses; 1e7 points do not fit torch autograd)",
                  "cpu_ms_per_step": ms_cpu, "cpu_evals_per_s": n3 * 100 / (ms_cpu * 1e-3), "cpu_threads": threads,
                  "torch_cuda_eager_ms_per_step": ms_gpu, "torch_cuda_eager_evals_per_s": n3 * 100 / (ms_gpu * 1e-3),
                  "extrapolated_cpu_s_per_c3_step": ms_cpu * 1e-3 * 1e7 / n3}), flush=True)
