#!/usr/bin/env python
"""Development probe: ModelTraj optimisation step (zero_grad + forward + backward + Adam) on a large cloud, eager vs
graphs.GraphedStep.  `model.rewards` keeps the per-point tensor of the last step alive — does that cost replays anything?
usage: graphstep_probe.py [n_points]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from trajectory_optimization_b200 import model, multicam, tools  # noqa: E402
from trajectory_optimization_b200.graphs import GraphedStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda:0")
pts = bench.make_cloud_shard(n, 0, 1, dev)
K, iw, ih = tools.load_intrinsics(dev)
with torch.no_grad():
    t, q = multicam.camera_poses_from_body(bench.body_waypoints().to(dev), multicam.ring_rig(5))
P, Q = t.reshape(-1, 3).contiguous(), q.reshape(-1, 4).contiguous()


def timeit(fn, reps=20):
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


for fused in (False, True):
    m = model.ModelTraj(pts, P, Q, K, iw, ih, device=dev, fused_regularizers=fused)
    opt = torch.optim.Adam([{"params": [m.poses], "lr": 0.01}, {"params": [m.quats], "lr": 0.0}], capturable=True)

    def eager():
        opt.zero_grad()
        loss = m(vis_wps_dist=0.0)
        loss.backward()
        opt.step()

    ms_e = timeit(eager)
    gs = GraphedStep(m, opt, forward_kwargs={"vis_wps_dist": 0.0})
    ms_g = timeit(gs.step)
    print(f"ModelTraj {n} points x {P.shape[0]} poses, fused_regularizers={fused}: eager {ms_e:.3f} ms, GraphedStep {ms_g:.3f} ms",
          flush=True)
    del gs, m, opt
