"""N>1 path on CPU: two gloo ranks, each holding a shard of the cloud, run the product's orchestration
(trajectory_optimization_b200/ops.py: pass A -> all-reduce MIN/MAX -> pass B -> all-reduce SUM -> epilogue,
autograd chain) with the five C-ABI calls replaced by an oracle-backed stand-in (tests only), and must
reproduce the unsharded oracle's loss and pose gradients on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests._standins import OracleBackend  # noqa: E402


def _case():
    from tests.conftest import load_golden
    g = load_golden("traj_compact")  # min > 0: the arg-min/arg-max tie accumulators cross the all-reduce too
    return g["in_points"], g["in_poses"], g["in_quats"]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import coverage_oracle as orc
        from trajectory_optimization_b200 import ops
        ops._BACKEND = OracleBackend()
        pts, poses, quats = _case()
        n = len(pts)
        lo, hi = rank * n // world, (rank + 1) * n // world
        K = torch.from_numpy(orc.K_DEFAULT.copy())
        P = torch.from_numpy(poses.copy()).requires_grad_(True)
        Q = torch.from_numpy(quats.copy()).requires_grad_(True)
        rewards, mean = ops.coverage_traj(torch.from_numpy(pts[lo:hi].copy()), P, Q, K, orc.IMG_WIDTH, orc.IMG_HEIGHT,
                                          n_total=n, group=dist.group.WORLD)
        vis = 1.0 / (mean + 1e-6)
        vis.backward()
        T = torch.tensor([[0.1, -0.2, 0.05]], requires_grad=True)
        Qp = torch.tensor([[0.9, 0.1, -0.1, 0.2]], requires_grad=True)
        obs, total = ops.coverage_pose(torch.from_numpy(pts[lo:hi].copy()), T, Qp, K, orc.IMG_WIDTH, orc.IMG_HEIGHT,
                                       group=dist.group.WORLD)
        (1.0 / (total + 1e-6)).backward()
        q.put((rank, vis.item(), P.grad.numpy(), Q.grad.numpy(), rewards.detach().numpy(), total.item(), T.grad.numpy(),
               Qp.grad.numpy()))
    except Exception as e:  # surface the failure instead of letting the parent time out
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_sharded_objective_matches_unsharded_oracle():
    from oracle import coverage_oracle as orc
    from tests.conftest import rel_err
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    assert all(len(r) == 8 for r in results), results
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pts, poses, quats = _case()
    K, W, H = orc.load_intrinsics()
    ref = orc.traj_objective(pts, poses, quats, K, W, H, dtype=np.float64)
    refp = orc.pose_objective(pts, [[0.1, -0.2, 0.05]], [[0.9, 0.1, -0.1, 0.2]], K, W, H, dtype=np.float64)
    n = len(pts)
    for rank, vis, gp, gq, rew, total, gt, gqp in results:
        assert rel_err(vis, ref["vis"]) < 1e-6
        assert rel_err(gp, ref["g_poses"]) < 1e-5 and rel_err(gq, ref["g_quats"]) < 1e-5
        assert rel_err(rew, ref["rewards"][rank * n // world:(rank + 1) * n // world]) < 1e-6
        assert rel_err(total, refp["sum"]) < 1e-6
        g_loss = -float(refp["loss"]) ** 2  # d(1/(S+eps))/dS applied by torch to d S/d pose from the op
        assert rel_err(gt.ravel() * 1.0, refp["g_trans"]) < 1e-5 and rel_err(gqp.ravel(), refp["g_quat"]) < 1e-5
        assert g_loss < 0
    assert np.allclose(results[0][2], results[1][2]) and np.allclose(results[0][3], results[1][3])  # replicas agree


def _sweep_case():
    gen = np.random.default_rng(5)
    pts = (gen.random((1200, 3)) * np.array([16, 16, 4]) + np.array([-4, -4, -1])).astype(np.float32)
    T, Pn = 5, 3   # 5 trajectories over 2 ranks: an uneven split
    base = np.stack([np.linspace(0, 6, Pn), np.linspace(0, 3, Pn), np.zeros(Pn)], 1)
    poses = (base[None] + gen.normal(0, 0.8, (T, 1, 3)) * np.array([1, 1, 0])).astype(np.float32)
    quats = (gen.normal(0, 0.3, (T, Pn, 4)) + np.array([1.0, 0, 0, 0])).astype(np.float32)
    return pts, poses, quats


def _sweep_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import coverage_oracle as orc
        from trajectory_optimization_b200 import ops
        ops._BACKEND = OracleBackend()
        pts, poses, quats = _sweep_case()
        n = len(pts)
        K = torch.from_numpy(orc.K_DEFAULT.copy())
        P, Q = torch.from_numpy(poses), torch.from_numpy(quats)
        by_traj = ops.sweep_rewards(torch.from_numpy(pts), P, Q, K, orc.IMG_WIDTH, orc.IMG_HEIGHT, group=dist.group.WORLD,
                                    shard="trajectories", presorted=True)
        lo, hi = rank * n // world, (rank + 1) * n // world
        by_pts = ops.sweep_rewards(torch.from_numpy(pts[lo:hi].copy()), P, Q, K, orc.IMG_WIDTH, orc.IMG_HEIGHT, n_total=n,
                                   group=dist.group.WORLD, shard="points", presorted=True)
        q.put((rank, by_traj.numpy(), by_pts.numpy()))
    except Exception as e:
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_sweep_trajectory_and_point_sharding_match_the_oracle():
    """BASELINE config 5 on N > 1 ranks: trajectories sharded (whole cloud per rank, no data-path
    collective, one all-gather) and points sharded (MAX + SUM all-reduces) both give the unsharded per-trajectory means."""
    from oracle import coverage_oracle as orc
    from tests.conftest import rel_err
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    assert all(len(r) == 3 for r in results), results
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pts, poses, quats = _sweep_case()
    K, W, H = orc.load_intrinsics()
    ref = np.array([orc.traj_objective(pts, poses[t], quats[t], K, W, H, dtype=np.float64, want_grad=False)["mean"]
                    for t in range(len(poses))])
    for rank, by_traj, by_pts in results:
        assert by_traj.shape == ref.shape and by_pts.shape == ref.shape
        assert rel_err(by_traj, ref) < 1e-9 and rel_err(by_pts, ref) < 1e-9


def _rig_case():
    gen = np.random.default_rng(23)
    d = gen.normal(size=(4000, 3))
    pts = (d / np.linalg.norm(d, axis=1, keepdims=True) * gen.uniform(1.5, 9.0, (4000, 1))).astype(np.float32)
    yaws = np.deg2rad([0, 72, -72, 144, -144])   # 5 cameras over 2 ranks: an uneven split (3 + 2)
    quats = np.stack([np.cos(yaws / 2), 0 * yaws, 0 * yaws, np.sin(yaws / 2)], 1).astype(np.float32)
    # rotate the optical axis (+z) into the horizontal plane first: q = q_yaw * q_tilt, q_tilt = rotation by -90 deg about x
    s = np.float32(np.sqrt(0.5))
    tilt = np.array([s, -s, 0, 0], np.float32)

    def qmul(a, b):
        w1, x1, y1, z1 = a
        w2, x2, y2, z2 = b
        return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                         w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], np.float32)

    quats = np.stack([qmul(q, tilt) for q in quats])
    trans = np.stack([[0.1 * i, -0.05 * i, 0.02 * i] for i in range(5)]).astype(np.float32)
    return pts, trans, quats


def _install_cpu_visibility_standins(ops, orc):
    """Oracle-backed stand-ins for the three C-ABI wrappers the per-camera pipeline calls (tests only)."""
    def frustum_cull(points_nx3, intrins, img_width, img_height, min_dist=1.0, max_dist=10.0):
        p = points_nx3.detach().cpu().numpy().astype(np.float32)
        _, dm, fm = orc.frustum_cull(p.T, img_height, img_width, intrins.detach().cpu().numpy(), min_dist, max_dist)
        return torch.from_numpy(np.flatnonzero(dm & fm).astype(np.int64)), torch.from_numpy(dm), torch.from_numpy(fm)

    def spherical_flip(points, param):
        f, radius, _ = orc.spherical_flip(points.detach().cpu().numpy(), param)
        return torch.from_numpy(f), torch.tensor(float(radius))

    def hpr_hull_mask(flipped, return_info=False):
        from scipy.spatial import ConvexHull
        f = flipped.detach().cpu().numpy()
        hull = ConvexHull(np.concatenate([f, np.zeros((1, 3), np.float32)], 0))
        v = np.sort(hull.vertices)
        mask = np.zeros(len(f), np.uint8)
        mask[v[v < len(f)]] = 1
        return torch.from_numpy(mask), bool(v[-1] == len(f)), 0

    ops.frustum_cull, ops.spherical_flip, ops.hpr_hull_mask = frustum_cull, spherical_flip, hpr_hull_mask


def _rig_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import coverage_oracle as orc
        from trajectory_optimization_b200 import ops, tools
        _install_cpu_visibility_standins(ops, orc)
        pts, trans, quats = _rig_case()
        K = torch.from_numpy(orc.K_DEFAULT.copy())
        res = tools.multi_camera_visibility(torch.from_numpy(pts), torch.from_numpy(trans), torch.from_numpy(quats), K,
                                            orc.IMG_HEIGHT, orc.IMG_WIDTH, 1.0, 10.0, device=torch.device("cpu"),
                                            group=dist.group.WORLD)
        q.put((rank, [(r["camera"], r["frustum_idx"].numpy(), r["visible_idx"].numpy()) for r in res]))
    except Exception as e:
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_multi_camera_visibility_replicates_cameras():
    """Per-camera visibility over N > 1 ranks is REPLICATED work (HPR does not shard): rank r takes cameras
    r, r + world, ...; together the ranks cover every camera once, and each camera's index sets are what one process
    computes for it (reference src/pc_processor.py:158-182 per camera)."""
    from oracle import coverage_oracle as orc
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rig_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    assert all(isinstance(r[1], list) for r in results), results
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [c for c, _, _ in results[0][1]] == [0, 2, 4] and [c for c, _, _ in results[1][1]] == [1, 3]
    pts, trans, quats = _rig_case()
    K, W, H = orc.load_intrinsics()
    seen = 0
    for _, cams in results:
        for c, fr_idx, vis_idx in cams:
            cam = orc.to_camera_frame(pts, quats[c], trans[c], dtype=np.float32)
            cam = np.asarray(cam, np.float32).reshape(-1, 3)
            culled, dm, fm = orc.frustum_cull(cam.T, H, W, K, 1.0, 10.0)
            ref_idx = np.flatnonzero(dm & fm)
            assert len(ref_idx) > 50 and np.array_equal(fr_idx, ref_idx)
            vis, _ = orc.hidden_pts_removal(culled, 2)
            assert np.array_equal(vis_idx, ref_idx[vis]) and 0 < len(vis) < len(ref_idx)
            seen += 1
    assert seen == 5
