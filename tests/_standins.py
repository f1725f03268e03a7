"""CPU stand-ins for the product's two device back ends — TEST INFRASTRUCTURE ONLY.

`OracleBackend` has the methods and buffer layouts of `ops.CudaBackend` (include/coverage_b200.h), computed by the
numpy oracle; `NumpyCodec` has the methods of `pointcloud_utils.CudaCodec`, restating the reference's host codec
(src/pointcloud_utils.py:58-80,180-198,290-338).  They let the GPU-less container exercise everything AROUND the C ABI:
the autograd plumbing, the collectives over gloo (tests/test_dist_gloo.py) and the unmodified reference nodes
(tests/test_reference_nodes.py).  Nothing under trajectory_optimization_b200/ imports this file."""
import numpy as np
import torch

ACC = 22


class OracleBackend:
    """Same methods and buffer layouts as ops.CudaBackend (include/coverage_b200.h), computed by the numpy oracle."""

    def prepare(self, t, device=None, what="tensor"):
        return t.detach().float().contiguous()

    def order_cloud(self, pts, spatial_sort_cloud=True):
        return pts, None, None   # no ordering on the CPU stand-in: (cloud, permutation, boxes)

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

    def traj_workspace_bytes(self, pts, W):
        return 0

    def prefill_applies(self, pts, dense=None):
        return False

    def traj_minmax(self, pts, P, Q, Kd, cam, boxes=None, ws=None, dense=None, prefill=None):
        mm = [self._vis(pts, P[w], Q[w], Kd, cam)[0] for w in range(len(P))]
        # fp64 so that the stand-in's second pass can find its arg-min/arg-max by exact comparison
        return torch.tensor([m.min() for m in mm] + [m.max() for m in mm], dtype=torch.float64)

    def traj_fused(self, pts, P, Q, Kd, cam, minmax, upstream, rewards, reward_index=None, boxes=None, ws=None,
                   dense=None, prefilled=False):
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

    def sweep_minmax(self, pts, P, Q, Kd, cam, boxes=None):
        return self.traj_minmax(pts, P, Q, Kd, cam)

    def sweep_sums(self, pts, P, Q, T, Pn, Kd, cam, boxes, minmax):
        W, hi = len(P), float(np.float32(1.0 - cam.eps))
        sums = np.zeros(T)
        for t in range(T):
            L = np.zeros(len(pts))
            for w in range(t * Pn, (t + 1) * Pn):
                m = self._vis(pts, P[w], Q[w], Kd, cam)[0]
                a, b = float(minmax[w]), float(minmax[W + w]) - float(minmax[w])
                qc = np.clip((m - a) / b, 0.5, hi)
                L += np.log(qc / (1 - qc))
            sums[t] = (1 / (1 + np.exp(-L))).sum()
        return torch.from_numpy(sums)

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


class NumpyCodec:
    """pointcloud_utils.CudaCodec on the host with numpy (same results as the reference's record-array codec)."""

    def default_device(self):
        return torch.device("cpu")

    def stage_bytes(self, data, nbytes, device):
        return torch.from_numpy(np.frombuffer(bytes(data), dtype=np.uint8, count=nbytes).copy())

    def to_device(self, points, device):
        t = points if isinstance(points, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(points))
        return t.detach().cpu()

    def pc2_to_xyz(self, data_dev, n, point_step, off_x, off_y, off_z, datatype, remove_nans):
        raw = data_dev.numpy().reshape(n, point_step) if n else np.zeros((0, max(point_step, 1)), np.uint8)
        dt = np.float32 if datatype == 7 else np.float64
        w = np.dtype(dt).itemsize
        cols = [raw[:, o:o + w].copy().view(dt).reshape(-1) for o in (off_x, off_y, off_z)]
        xyz = np.stack(cols, 1)
        if remove_nans:
            xyz = xyz[np.isfinite(xyz).all(1)]
        return torch.from_numpy(xyz.astype(np.float32))

    def xyz_to_pc2(self, pts, extra):
        a = pts.numpy().astype(np.float32)
        if extra is not None:
            a = np.concatenate([a, extra.numpy().astype(np.float32).reshape(-1, 1)], 1)
        return torch.from_numpy(np.frombuffer(a.tobytes(), dtype=np.uint8).copy()), bool(np.isfinite(a).all())
