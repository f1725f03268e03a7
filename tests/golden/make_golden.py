#!/usr/bin/env python
"""Regenerate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE on CPU.

Runs only in the build container (needs /root/reference, which is absent on the
GPU box).  The reference's `src/model.py` / `src/tools.py` are imported as-is
through the stubs in oracle/shims (ROS modules, pytorch3d.transforms,
numpy.float alias — SURVEY.md App. B); their outputs on seeded inputs are the
pin for the oracle (tests/test_oracle_golden.py) and for the CUDA path
(tests/test_gpu_*.py).

    python tests/golden/make_golden.py            # rewrites the fixtures in place

Each fixture stores inputs, the reference's fp32 outputs/gradients and an fp64
"truth" computed with the reference's own helper functions on double tensors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), os.path.join(REF, "src")]
np.float = float  # src/pointcloud_utils.py:180
if not hasattr(np, "_fromstring_text"):  # src/pointcloud_utils.py:71 uses the binary mode of np.fromstring, which
    np._fromstring_text = np.fromstring  # numpy >= 2.3 removed; it was frombuffer + copy
    np.fromstring = lambda data, dtype=float, **kw: np.frombuffer(data, dtype=dtype).copy()

import torch  # noqa: E402
import model as ref_model  # noqa: E402  (the reference)
import tools as ref_tools  # noqa: E402  (the reference)

torch.set_num_threads(max(1, os.cpu_count() or 1))
CPU = torch.device("cpu")
K, IMG_W, IMG_H = ref_tools.load_intrinsics(CPU)


def rand_quats(gen, n, spread=1.0, scale_jitter=True):
    q = torch.randn(n, 4, generator=gen) * spread + torch.tensor([1.0, 0, 0, 0])
    if not scale_jitter:
        q = q / q.norm(dim=1, keepdim=True)
    return q.float()


def run_pose(points, trans, quat, min_d=1.0, max_d=5.0, weight=None):
    m = ref_model.ModelPose(points, trans, quat, K, IMG_W, IMG_H, min_d, max_d, CPU)
    if weight is None:
        loss = m()
    else:  # same arithmetic as the hpr=True branch (src/model.py:112-117) with a given mask
        pts = ref_model.to_camera_frame(m.points, m.quat, m.trans)
        mask = ref_model.get_dist_mask(pts, min_d, max_d) * ref_model.get_fov_mask(pts, m.img_height, m.img_width, m.K, eps=m.eps)
        m.observations = weight * mask
        loss = m.criterion(m.observations)
    loss.backward()
    out = dict(loss=loss.item(), obs=m.observations.detach().numpy(), g_trans=m.trans.grad.numpy().copy(),
               g_quat=m.quat.grad.numpy().copy())
    # fp64 truth with the reference's helper functions
    t64 = trans.double().clone().requires_grad_(True)
    q64 = quat.double().clone().requires_grad_(True)
    pts = ref_model.to_camera_frame(points.double(), q64, t64)
    mask = ref_model.get_dist_mask(pts, min_d, max_d) * ref_model.get_fov_mask(pts, IMG_H, IMG_W, K.double(), eps=1e-6)
    if weight is not None:
        mask = weight.double() * mask
    l64 = 1.0 / (mask.sum() + 1e-6)
    l64.backward()
    out.update(loss64=l64.item(), obs64=mask.detach().numpy(), g_trans64=t64.grad.numpy().copy(),
               g_quat64=q64.grad.numpy().copy())
    return out


def traj64(points, poses, quats, sel, min_d, max_d):
    p64 = poses.double().clone().requires_grad_(True)
    q64 = quats.double().clone().requires_grad_(True)
    lo_sum = 0.0
    for i in sel:
        pts = ref_model.to_camera_frame(points.double(), q64[i].unsqueeze(0), p64[i].unsqueeze(0))
        p = ref_model.get_dist_mask(pts, min_d, max_d) * ref_model.get_fov_mask(pts, IMG_H, IMG_W, K.double(), eps=1e-6)
        p = p - p.min()
        p = p / p.max()
        p = torch.clip(p, 0.5, 1.0 - 1e-6)
        lo_sum = lo_sum + torch.log(p / (1.0 - p))
    r = 1.0 / (1.0 + torch.exp(-lo_sum))
    vis = 1.0 / (torch.mean(r) + 1e-6)
    vis.backward()
    return dict(rewards64=r.detach().numpy(), vis64=vis.item(), gv_poses64=p64.grad.numpy().copy(),
                gv_quats64=q64.grad.numpy().copy())


def run_traj(points, poses, quats, min_d=1.0, max_d=5.0, vis_wps_dist=0.5, sw=14.0, lw=0.02):
    m = ref_model.ModelTraj(points, poses, quats, K, IMG_W, IMG_H, min_d, max_d, sw, lw, CPU)
    loss = m(vis_wps_dist=vis_wps_dist)
    # split gradient: visibility term alone, then the full objective
    gv = torch.autograd.grad(m.loss["vis"], [m.poses, m.quats], retain_graph=True)
    loss.backward()
    mean_d = (m.poses0[1:] - m.poses0[:-1]).norm(dim=1).mean()
    step = int(vis_wps_dist / mean_d) + 1
    out = dict(loss=loss.item(), vis=float(m.loss["vis"]), l2=float(m.loss["l2"]), smooth=float(m.loss["smooth"]),
               length=float(m.loss["length"]), rewards=m.rewards.detach().numpy(), wps_step=step,
               gv_poses=gv[0].numpy().copy(), gv_quats=gv[1].numpy().copy(),
               g_poses=m.poses.grad.numpy().copy(), g_quats=m.quats.grad.numpy().copy())
    out.update(traj64(points, poses, quats, list(range(0, len(poses), step)), min_d, max_d))
    return out


def save(name, inputs, outputs):
    d = {f"in_{k}": np.asarray(v) for k, v in inputs.items()}
    d.update({f"out_{k}": np.asarray(v) for k, v in outputs.items()})
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def box_cloud(gen, n, lo=(-10, -10, -1), hi=(30, 30, 4)):
    lo, hi = torch.tensor(lo, dtype=torch.float32), torch.tensor(hi, dtype=torch.float32)
    return (torch.rand(n, 3, generator=gen) * (hi - lo) + lo).float()


def codec_fixtures():
    """PointCloud2 decode (src/pointcloud_utils.py:58-80,180-198) on synthetic messages of three layouts."""
    import pointcloud_utils as ref_pc  # the reference
    from sensor_msgs.msg import PointCloud2, PointField
    gnp = np.random.default_rng(7)

    def message(n, step, fields, nan_frac):
        # random payload bytes, then the x/y/z fields overwritten with values (some NaN / inf)
        raw = gnp.integers(0, 256, size=(n, step), dtype=np.uint8)
        vals = gnp.normal(0, 10, (n, 3))
        bad = gnp.random((n, 3)) < nan_frac
        vals[bad] = gnp.choice([np.nan, np.inf, -np.inf], size=int(bad.sum()))
        for k, (name, off, dt) in enumerate(f for f in fields if f[0] in "xyz"):
            b = vals[:, k].astype(np.float32 if dt == 7 else np.float64).view(np.uint8).reshape(n, 4 if dt == 7 else 8)
            raw[:, off:off + b.shape[1]] = b
        for name, off, dt in fields:   # other float fields must hold valid floats for np.isfinite on the record
            if name not in "xyz" and dt in (7, 8):
                b = gnp.normal(0, 1, n).astype(np.float32 if dt == 7 else np.float64).view(np.uint8).reshape(n, 4 if dt == 7 else 8)
                raw[:, off:off + b.shape[1]] = b
        msg = PointCloud2()
        msg.height, msg.width, msg.point_step, msg.row_step = 1, n, step, n * step
        msg.fields = [PointField(nm, off, dt, 1) for nm, off, dt in fields]
        msg.data = raw.tobytes()
        return msg, raw

    cases = {
        "pc2_xyz12": (3000, 12, [("x", 0, 7), ("y", 4, 7), ("z", 8, 7)], 0.02),
        "pc2_xyzi_padded32": (2049, 32, [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 7), ("ring", 20, 4)], 0.05),
        "pc2_unaligned_f64": (1000, 29, [("tag", 0, 2), ("x", 1, 8), ("y", 9, 8), ("z", 17, 8), ("i", 25, 7)], 0.1),
        "pc2_dense": (300, 16, [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("i", 12, 7)], 0.0),
        "pc2_empty": (0, 16, [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("i", 12, 7)], 0.0),
    }
    for name, (n, step, fields, frac) in cases.items():
        msg, raw = message(n, step, fields, frac)
        out = ref_pc.pointcloud2_to_xyz_array(msg) if n else np.zeros((0, 3))
        out_all = ref_pc.pointcloud2_to_xyz_array(msg, remove_nans=False) if n else np.zeros((0, 3))
        save(name, dict(data=raw.reshape(-1), n=n, point_step=step,
                        field_names=np.array([f[0] for f in fields]), field_offsets=np.array([f[1] for f in fields]),
                        field_types=np.array([f[2] for f in fields])),
             dict(xyz=np.asarray(out, np.float64).reshape(-1, 3), xyz_all=np.asarray(out_all, np.float64).reshape(-1, 3)))


def opt_loop_fixtures():
    """The optimisation loops of src/pose_optimization.py:82-137 and src/trajectory_optimization.py:83-116 (model +
    Adam with per-parameter learning rates, zero_grad / forward / backward / step), 10 steps on the sample inputs."""
    sample = np.load(os.path.join(HERE, "sample_inputs.npz"))
    pts = torch.from_numpy(sample["pts"]).float()
    steps, lr_pose, lr_quat = 10, 0.1, 0.02
    m = ref_model.ModelPose(pts, torch.tensor([[6.0, 2.0, 0.0]]), torch.tensor([[1.0, 0.0, 0.0, 0.0]]), K, IMG_W, IMG_H,
                            1.0, 5.0, CPU)
    opt = torch.optim.Adam([{"params": [m.trans], "lr": lr_pose}, {"params": [m.quat], "lr": lr_quat}])
    hist = []
    for _ in range(steps):
        opt.zero_grad()
        loss = m()
        loss.backward()
        opt.step()
        hist.append(loss.item())
    save("opt_pose_sample", dict(trans0=np.array([[6.0, 2.0, 0.0]], np.float32), quat0=np.array([[1.0, 0, 0, 0]], np.float32),
                                 steps=steps, lr_pose=lr_pose, lr_quat=lr_quat),
         dict(loss_history=np.array(hist), trans=m.trans.detach().numpy().copy(), quat=m.quat.detach().numpy().copy()))
    poses = torch.from_numpy(sample["poses"]).float()
    quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]]).repeat(poses.shape[0], 1)
    mt = ref_model.ModelTraj(pts, poses, quats, K, IMG_W, IMG_H, device=CPU)
    opt = torch.optim.Adam([{"params": [mt.poses], "lr": lr_pose}, {"params": [mt.quats], "lr": lr_quat}])
    hist, vis = [], []
    for _ in range(steps):
        opt.zero_grad()
        loss = mt()
        loss.backward()
        opt.step()
        hist.append(loss.item())
        vis.append(float(mt.loss["vis"]))
    save("opt_traj_sample", dict(steps=steps, lr_pose=lr_pose, lr_quat=lr_quat),
         dict(loss_history=np.array(hist), vis_history=np.array(vis), poses=mt.poses.detach().numpy().copy(),
              quats=mt.quats.detach().numpy().copy(), mean_reward=float(mt.rewards.mean())))


def main():
    if "--only-codec" in sys.argv:
        return codec_fixtures()
    if "--only-opt" in sys.argv:
        return opt_loop_fixtures()
    sample = np.load(os.path.join(REF, "data/points/point_cloud_10.npz"))["pts"].astype(np.float32)
    path = np.load(os.path.join(REF, "data/paths/path_poses_10.npz"))["poses"].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "sample_inputs.npz"), pts=sample, poses=path)
    pts = torch.from_numpy(sample)

    # ---- ModelPose on the shipped sample (src/pose_optimization_sample.py:59,72) ----
    t0 = torch.tensor([[6.0, 2.0, 0.0]])
    q0 = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    save("pose_sample", dict(trans=t0, quat=q0, min_d=1.0, max_d=5.0), run_pose(pts, t0, q0))

    # ---- ModelPose, seeded synthetic clouds, random unnormalised quaternions ----
    g = torch.Generator().manual_seed(1234)
    for i, n in enumerate([1, 37, 5000]):
        cloud = box_cloud(g, n, (-4, -4, -1), (8, 8, 4))
        t = (torch.rand(1, 3, generator=g) * 4 - 1).float()
        q = rand_quats(g, 1, 0.7)
        wgt = (torch.rand(n, generator=g) > 0.4).float() if i == 2 else None
        save(f"pose_synth{i}", dict(points=cloud, trans=t, quat=q, min_d=1.0, max_d=5.0,
                                    **({} if wgt is None else dict(weight=wgt))), run_pose(cloud, t, q, weight=wgt))
    cloud = box_cloud(g, 3000, (-3, -3, -1), (6, 6, 3))
    t = torch.tensor([[0.5, -0.25, 0.3]])
    q = rand_quats(g, 1, 0.5)
    save("pose_synth_clip", dict(points=cloud, trans=t, quat=q, min_d=0.5, max_d=8.0), run_pose(cloud, t, q, 0.5, 8.0))

    # ---- ModelTraj on the shipped sample (src/trajectory_optimization_sample.py:46-49) ----
    poses = torch.from_numpy(path)
    quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]]).repeat(len(poses), 1)
    save("traj_sample", dict(poses=poses, quats=quats, min_d=1.0, max_d=5.0, vis_wps_dist=0.5, sw=14.0, lw=0.02),
         run_traj(pts, poses, quats))
    # same, every waypoint evaluated, launch-file weights, random orientations
    quats_r = rand_quats(g, len(poses), 0.6)
    save("traj_sample_all", dict(poses=poses, quats=quats_r, min_d=1.0, max_d=5.0, vis_wps_dist=0.0, sw=28.0, lw=0.02),
         run_traj(pts, poses, quats_r, vis_wps_dist=0.0, sw=28.0))

    # ---- ModelTraj, synthetic: wide box (min underflows to 0, many ties) ----
    cloud = box_cloud(g, 20000)
    W = 12
    xs = torch.linspace(0, 9, W)
    poses = torch.stack([xs, 0.5 * xs + 0.3 * torch.sin(xs), torch.zeros(W)], 1).float()
    quats = rand_quats(g, W, 0.8)
    save("traj_box", dict(points=cloud, poses=poses, quats=quats, min_d=1.0, max_d=5.0, vis_wps_dist=0.5, sw=14.0, lw=0.02),
         run_traj(cloud, poses, quats))
    # ---- ModelTraj, compact cloud: min > 0, unique arg-min/arg-max carry gradient ----
    cloud = box_cloud(g, 1500, (1.0, 1.0, 1.5), (4.0, 4.0, 4.5))
    W = 5
    poses = (torch.rand(W, 3, generator=g) * 0.6 - 0.3).float()
    quats = rand_quats(g, W, 0.15)
    save("traj_compact", dict(points=cloud, poses=poses, quats=quats, min_d=1.0, max_d=5.0, vis_wps_dist=0.0, sw=14.0, lw=0.02),
         run_traj(cloud, poses, quats, vis_wps_dist=0.0))
    # ---- ModelTraj, tiny ragged sizes ----
    cloud = box_cloud(g, 33, (1.0, 1.0, 1.5), (4.0, 4.0, 4.5))
    poses = (torch.rand(3, 3, generator=g) * 0.6 - 0.3).float()
    quats = rand_quats(g, 3, 0.15)
    save("traj_tiny", dict(points=cloud, poses=poses, quats=quats, min_d=1.0, max_d=5.0, vis_wps_dist=0.0, sw=14.0, lw=0.02),
         run_traj(cloud, poses, quats, vis_wps_dist=0.0))

    # ---- binary frustum cull (src/tools.py:176-187) in the camera frame ----
    cloud = box_cloud(g, 30000, (-12, -12, -2), (12, 12, 14))
    culled, dm, fm = ref_tools.get_cam_frustum_pts(cloud.T.clone(), IMG_H, IMG_W, K, 1.0, 10.0)
    fb = ref_model.get_fov_mask(cloud, IMG_H, IMG_W, K, binary=True)
    save("cull", dict(points=cloud, min_d=1.0, max_d=10.0),
         dict(culled=culled.numpy(), dist_mask=dm.numpy(), fov_mask=fm.numpy(), fov_binary_model=fb.numpy()))

    # ---- masks / frame helpers on a few points (src/model.py:13-57) ----
    cloud = box_cloud(g, 257, (-4, -4, -1), (8, 8, 6))
    q = rand_quats(g, 1, 0.7)
    t = torch.tensor([[0.3, -1.2, 0.4]])
    cam = ref_model.to_camera_frame(cloud, q, t)
    save("helpers", dict(points=cloud, trans=t, quat=q),
         dict(cam=cam.numpy(), dist=ref_model.get_dist_mask(cam, 1.0, 5.0).numpy(),
              fov=ref_model.get_fov_mask(cam, IMG_H, IMG_W, K).numpy()))

    # ---- HPR (src/tools.py:38-85): shell cloud around the camera, and a half-space cloud ----
    gnp = np.random.default_rng(1)
    for name, n in [("hpr_shell", 20000), ("hpr_shell_small", 500)]:
        d = gnp.standard_normal((n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        cloud = torch.from_numpy((d * gnp.uniform(2, 8, (n, 1))).astype(np.float32))
        flipped = ref_tools.sphericalFlip(cloud, CPU, 2)
        vis, mask = ref_tools.hidden_pts_removal(cloud, CPU)
        idx = np.flatnonzero(mask.numpy())
        save(name, dict(points=cloud, R_param=2), dict(flipped=flipped.numpy(), idx=idx, visible=vis.numpy()))
    cloud = box_cloud(g, 20000, (-10, -10, 2), (10, 10, 6))
    flipped = ref_tools.sphericalFlip(cloud, CPU, 2)
    vis, mask = ref_tools.hidden_pts_removal(cloud, CPU)
    save("hpr_halfspace", dict(points=cloud, R_param=2),
         dict(flipped=flipped.numpy(), idx=np.flatnonzero(mask.numpy()), visible=vis.numpy()))
    # sample cloud seen from the sample camera pose (camera frame, as src/pc_processor.py:168-178)
    cam = ref_model.to_camera_frame(pts, q0, t0)
    vis, mask = ref_tools.hidden_pts_removal(cam, CPU)
    save("hpr_sample", dict(points=cam, R_param=2), dict(idx=np.flatnonzero(mask.numpy())))
    codec_fixtures()
    opt_loop_fixtures()


if __name__ == "__main__":
    main()
