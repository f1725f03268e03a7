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

ACC = 22


class OracleBackend:
    """Same methods and buffer layouts as ops.CudaBackend (include/coverage_b200.h), computed by the numpy oracle."""

    def prepare(self, t, device=None, what="tensor"):
        return t.detach().float().contiguous()

    @staticmethod
    def _cam(cam):
        return dict(img_width=cam.img_width, img_height=cam.img_height, min_dist=cam.min_dist, max_dist=cam.max_dist,
                    eps=cam.eps)

    def _vis(self, pts, t, q, Kd, cam):
        from oracle import coverage_oracle as orc
        c = self._cam(cam)
        m, y, g_c, R, qn, nrm = orc.visibility(pts.numpy(), t.numpy(), q.numpy(), Kd.numpy().reshape(3, 3), c["img_width"],
                                               c["img_height"], c["min_dist"], c["max_dist"], c["eps"], np.float64, True)
        return m, y, g_c @ R.T  # m, lever arm y = x - t, world-frame gradient dm/dx

    def pose_fused(self, pts, t, q, Kd, cam, w, obs):
        m, y, gy = self._vis(pts, t, q, Kd, cam)
        if w is not None:
            m, gy = m * w.numpy(), gy * w.numpy()[:, None]
        if obs is not None:
            obs.copy_(torch.from_numpy(m.astype(np.float32)))
        acc = np.zeros(8)
        acc[0], acc[1:4], acc[4:7] = m.sum(), gy.sum(0), np.cross(gy, y).sum(0)
        return torch.from_numpy(acc)

    @staticmethod
    def _quat_grad(T, q):
        q = q.double().numpy()
        n = max(np.linalg.norm(q), 1e-12)
        w, x, y, z = q / n
        return 2.0 / n * np.array([-T[0] * x - T[1] * y - T[2] * z, T[0] * w + T[1] * z - T[2] * y,
                                   T[1] * w - T[0] * z + T[2] * x, T[2] * w + T[0] * y - T[1] * x])

    def pose_epilogue(self, acc, t, q):
        a = acc.numpy()
        return torch.tensor(np.concatenate([[a[0]], -a[1:4], self._quat_grad(a[4:7], q)]), dtype=torch.float32)

    def traj_workspace(self, pts, W):
        return None

    def traj_minmax(self, pts, P, Q, Kd, cam, boxes=None, ws=None):
        mm = [self._vis(pts, P[w], Q[w], Kd, cam)[0] for w in range(len(P))]
        # fp64 so that the stand-in's second pass can find its arg-min/arg-max by exact comparison
        return torch.tensor([m.min() for m in mm] + [m.max() for m in mm], dtype=torch.float64)

    def traj_fused(self, pts, P, Q, Kd, cam, minmax, upstream, rewards, reward_index=None, boxes=None, ws=None):
        W, hi = len(P), float(np.float32(1.0 - cam.eps))
        acc, L, keep = np.zeros(W * ACC + 1), np.zeros(len(pts)), []
        for w in range(W):
            m, y, gy = self._vis(pts, P[w], Q[w], Kd, cam)
            a, b = float(minmax[w]), float(minmax[W + w]) - float(minmax[w])
            p = (m - a) / b
            qc = np.clip(p, 0.5, hi)
            L += np.log(qc / (1 - qc))
            keep.append((m, y, gy, p, qc, a, b))
        r = 1 / (1 + np.exp(-L))
        rewards.copy_(torch.from_numpy(r.astype(np.float32)))
        G = r * (1 - r) * (1.0 if upstream is None else upstream.double().numpy())
        for w, (m, y, gy, p, qc, a, b) in enumerate(keep):
            gate = (p >= 0.5) & (p <= hi)
            e = np.where(gate, G / (qc * (1 - qc)), 0.0)
            row = acc[w * ACC:(w + 1) * ACC]
            row[0:3], row[3:6] = ((e / b)[:, None] * gy).sum(0), ((e / b)[:, None] * np.cross(gy, y)).sum(0)
            row[6], row[7] = e.sum(), (e * p).sum()
            for off, sel in ((8, (m - a) == b), (15, (m == a) & (a > 0))):
                row[off:off + 3], row[off + 3:off + 6], row[off + 6] = gy[sel].sum(0), np.cross(gy[sel], y[sel]).sum(0), sel.sum()
        acc[W * ACC] = r.sum()
        return torch.from_numpy(acc)

    def traj_epilogue(self, acc, minmax, Q, n_total, upstream_mode):
        a, W = acc.numpy(), len(Q)
        out = np.zeros(1 + 7 * W)
        out[0] = a[W * ACC] / n_total
        c0 = 1.0 if upstream_mode else 1.0 / n_total
        for w in range(W):
            r = a[w * ACC:(w + 1) * ACC]
            b = float(minmax[W + w]) - float(minmax[w])
            dLdb = -r[7] / b
            dLda = -r[6] / b - dLdb
            F, T = r[0:3].copy(), r[3:6].copy()
            if r[14] > 0:
                F += dLdb / r[14] * r[8:11]
                T += dLdb / r[14] * r[11:14]
            if r[21] > 0:
                F += dLda / r[21] * r[15:18]
                T += dLda / r[21] * r[18:21]
            out[1 + 3 * w:4 + 3 * w] = -c0 * F
            out[1 + 3 * W + 4 * w:5 + 3 * W + 4 * w] = c0 * self._quat_grad(T, Q[w])
        return torch.tensor(out, dtype=torch.float32)


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
