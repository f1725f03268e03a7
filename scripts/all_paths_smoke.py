#!/usr/bin/env python
"""One small call of every entry point of the library (pruned objective + gradient incl. the upstream path, ModelPose,
sweep, codec, voxel filter, HPR, frustum cull): a quick "does everything still run" check on a GPU box, with sizes
chosen so that every kernel of the pruned pipeline launches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_optimization_b200 import model, ops, pointcloud_utils as pcu, tools  # noqa: E402

dev = torch.device("cuda:0")
gen = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150_001
pts = torch.from_numpy((gen.random((n, 3), dtype=np.float32) * np.array([40, 40, 5], np.float32)
                        + np.array([-10, -10, -1], np.float32))).to(dev)
W = 40
xs = np.linspace(0, 16, W)
poses = torch.tensor(np.stack([xs, 0.5 * xs + 0.3 * np.sin(xs), np.zeros(W)], 1), dtype=torch.float32)
quats = torch.tensor(gen.standard_normal((W, 4)) * 0.3 + np.array([1.0, 0, 0, 0]), dtype=torch.float32)
K, iw, ih = tools.load_intrinsics(dev)
m = model.ModelTraj(pts, poses, quats, K, iw, ih, device=dev, fused_regularizers=True)
loss = m(vis_wps_dist=0.0)
loss.backward(retain_graph=True)
(m.rewards * torch.linspace(0.5, 1.5, n, device=dev)).sum().backward()      # upstream path
mp = model.ModelPose(pts, torch.tensor([[6.0, 2.0, 0.0]]), torch.tensor([[0.9, 0.1, 0.0, 0.3]]), K, iw, ih, device=dev)
mp().backward()
T = 6
means = ops.sweep_rewards(pts, poses[None, :8].repeat(T, 1, 1) + torch.randn(T, 1, 3) * 0.5, quats[None, :8].repeat(T, 1, 1), K, iw, ih)
payload, dense = pcu.xyz_to_payload(pts)
back = pcu.payload_to_xyz(payload, n, 12, 0, 4, 8)
vox = tools.voxel_grid_filter(pts, 0.25)
sh = torch.randn(20_000, 3, device=dev)
sh = sh / sh.norm(dim=1, keepdim=True) * (2 + 6 * torch.rand(20_000, 1, device=dev))
vis, mask = tools.hidden_pts_removal(sh, dev)
culled, dm, fm = tools.get_cam_frustum_pts(pts.t(), ih, iw, K)
torch.cuda.synchronize()
print("ok", float(loss), float(means.mean()), back.shape[0], vox.shape[0], vis.shape[0], culled.shape[0])
