#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL): the point-sharded ModelTraj / ModelPose
give the same loss and pose gradients on every rank as the unsharded model on one GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_optimization_b200 import model, tools  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    gen = np.random.default_rng(0)
    n, W = 2_000_003, 24
    pts = (gen.random((n, 3), dtype=np.float32) * np.array([40, 40, 5], np.float32) + np.array([-10, -10, -1], np.float32))
    xs = np.linspace(0, 12, W)
    poses = torch.tensor(np.stack([xs, 0.5 * xs + 0.3 * np.sin(xs), np.zeros(W)], 1), dtype=torch.float32)
    quats = torch.tensor(gen.standard_normal((W, 4)) * 0.4 + np.array([1.0, 0, 0, 0]), dtype=torch.float32)
    K, iw, ih = tools.load_intrinsics(dev)
    bounds = np.linspace(0, n, world + 1).astype(np.int64)
    shard = torch.from_numpy(pts[bounds[rank]:bounds[rank + 1]])
    m = model.ModelTraj(shard, poses, quats, K, iw, ih, device=dev, group=dist.group.WORLD, n_total=n)
    loss = m(vis_wps_dist=0.0)
    loss.backward()
    full = model.ModelTraj(torch.from_numpy(pts), poses, quats, K, iw, ih, device=dev)
    loss_f = full(vis_wps_dist=0.0)
    loss_f.backward()

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())

    errs = dict(loss=rel(loss.detach(), loss_f.detach()), g_poses=rel(m.poses.grad, full.poses.grad),
                g_quats=rel(m.quats.grad, full.quats.grad),
                rewards_equal=bool(torch.equal(m.rewards, full.rewards[bounds[rank]:bounds[rank + 1]])))
    mp = model.ModelPose(shard, torch.tensor([[6.0, 2.0, 0.0]]), torch.tensor([[0.9, 0.1, 0.0, 0.3]]), K, iw, ih, device=dev,
                         group=dist.group.WORLD)
    lp = mp()
    lp.backward()
    fp = model.ModelPose(torch.from_numpy(pts), torch.tensor([[6.0, 2.0, 0.0]]), torch.tensor([[0.9, 0.1, 0.0, 0.3]]), K, iw, ih,
                         device=dev)
    lf = fp()
    lf.backward()
    errs.update(pose_loss=rel(lp.detach(), lf.detach()), pose_gt=rel(mp.trans.grad, fp.trans.grad),
                pose_gq=rel(mp.quat.grad, fp.quat.grad))
    gathered = [None] * world
    dist.all_gather_object(gathered, errs)
    if rank == 0:
        print("dist_check world", world, gathered)
        for e in gathered:
            assert e["rewards_equal"] and all(v < 1e-5 for k, v in e.items() if k != "rewards_equal"), e
        print("dist_check OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
