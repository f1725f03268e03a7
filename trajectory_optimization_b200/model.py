"""Drop-in for the reference's `src/model.py`: same names, signatures, attributes and asserts,
with the N-sized arithmetic replaced by the fused sm_100a kernels in libcovb200.so.

Reference citations are relative to the reference repository root.  The callers
(`src/pose_optimization.py:82-137`, `src/trajectory_optimization.py:83-127`) construct the
models, put `.trans/.quat` (or `.poses/.quats`) in Adam parameter groups, call `model()`,
`loss.backward()`, and read `.observations` / `.rewards` / `.loss[...]` — all of that works
unchanged.  The helper functions operate on CUDA tensors only (no CPU fallback).
"""
from copy import deepcopy
from time import time

import torch
import torch.nn as nn

from . import ops
from .tools import hidden_pts_removal, load_intrinsics  # noqa: F401  (re-exported like the reference does)


# ------------------------------------------------------------------------------------------
# helper functions (src/model.py:13-57)
# ------------------------------------------------------------------------------------------
def _unit_pose(like):
    t = torch.zeros(1, 3, device=like.device)
    q = torch.tensor([[1.0, 0.0, 0.0, 0.0]], device=like.device)
    return t, q


def to_camera_frame(verts, quat, trans):
    """src/model.py:50-57.  R(q)^T (v - t) with q = normalize(quat); small differentiable torch
    helper kept for API parity (the fused ops do this in-kernel)."""
    assert verts.dim() == trans.dim()
    assert quat.size() == torch.Size([1, 4])
    q = torch.nn.functional.normalize(quat)
    w, x, y, z = q[0, 0], q[0, 1], q[0, 2], q[0, 3]
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)]),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)]),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]),
    ]).to(verts.dtype)
    return (verts - trans) @ R


def get_dist_mask(points, min_dist=1.0, max_dist=5.0):
    """src/model.py:13-24: Gaussian about the camera-frame point (mu, mu, mu)."""
    assert isinstance(points, torch.Tensor)
    assert points.size()[1] == 3
    mean = (min_dist + max_dist) / 2.0
    std = (max_dist - min_dist) / 2.0
    sq = ((points - mean) ** 2).sum(dim=1)
    return torch.exp(-0.5 * sq / (std * std))


def get_fov_mask(points, img_height, img_width, intrins, eps=1e-6, binary=False):
    """src/model.py:27-47.  `binary=True` runs the bit-exact CUDA cull test; the smooth branch is a
    small differentiable torch helper kept for API parity (the fused ops do this in-kernel)."""
    assert isinstance(points, torch.Tensor)
    assert points.size()[1] == 3
    assert isinstance(intrins, torch.Tensor)
    assert intrins.size() == torch.Size([3, 3])
    if binary:
        _, _, fov = ops.frustum_cull(points, intrins, img_width, img_height, min_dist=float("-inf"),
                                     max_dist=float("inf"))
        return fov
    h = points @ intrins.t().to(points.dtype)
    z = h[:, 2]
    u = h[:, 0] / (z + eps)
    v = h[:, 1] / (z + eps)
    return (torch.sigmoid(z) * torch.exp(-0.5 * ((u - img_width / 2.0) / img_width) ** 2)
            * torch.exp(-0.5 * ((v - img_height / 2.0) / img_height) ** 2))


# ------------------------------------------------------------------------------------------
# single pose (src/model.py:65-127)
# ------------------------------------------------------------------------------------------
class ModelPose(nn.Module):
    def __init__(self, points, trans0, q0, intrins, img_width, img_height, min_dist=1.0, max_dist=5.0,
                 device=torch.device("cuda:0"), group=None):
        super().__init__()
        assert trans0.size() == torch.Size([1, 3])
        assert q0.size() == torch.Size([1, 4])
        assert intrins.size() == torch.Size([3, 3])

        self.device = device
        self.points = points.to(self.device)
        self.rewards = None
        self.observations = None
        self.lo_sum = 0.0
        self.trans = nn.Parameter(torch.as_tensor(trans0, dtype=torch.float32).to(self.device))
        self.quat = nn.Parameter(torch.as_tensor(q0, dtype=torch.float32).to(self.device))
        self.K = torch.as_tensor(intrins, dtype=torch.float32).to(self.device)
        self.img_width, self.img_height = float(img_width), float(img_height)
        self.eps = 1e-6
        self.pc_clip_limits = [min_dist, max_dist]
        self.group = group            # torch.distributed group when `points` is this rank's shard
        self._total = None
        self._pts32 = None
        self.to(self.device)

    def _cloud(self):
        # fp32, contiguous, aligned copy made once per cloud (the reference keeps whatever dtype it was given); the
        # cache follows `self.points` by identity AND by torch's version counter, so an in-place edit is seen too
        key = (id(self.points), self.points._version)
        if self._pts32 is None or self._pts32_key != key:
            self._pts32 = ops._BACKEND.prepare(self.points, what="points")
            self._pts32_key = key
        return self._pts32

    def set_points(self, points):
        """Replace the cloud (same as assigning `model.points`; the fp32 working copy is rebuilt on the next forward)."""
        self.points = points.to(self.device)
        self._pts32 = None

    def detach_state(self):
        """Drop the references this model keeps into the last autograd graph (`_total`) so that the graph — and the
        gradient-accumulation nodes bound to the stream it ran on — can be freed (graphs.GraphedStep calls this before
        it warms up on its capture stream)."""
        self._total = None
        if self.observations is not None:
            self.observations = self.observations.detach()

    def forward(self, debug=False, hpr=False):
        t0 = time()
        weight = None
        if hpr:  # src/model.py:112-115 — HPR of the untransformed cloud, mask multiplies the observations
            weight = hidden_pts_removal(self.points.detach(), device=self.device)[1]
        obs, total = ops.coverage_pose(self._cloud(), self.trans, self.quat, self.K, self.img_width, self.img_height,
                                       self.pc_clip_limits[0], self.pc_clip_limits[1], self.eps, weight=weight,
                                       group=self.group)
        self.observations = obs
        self._total = total
        if debug:
            if obs.is_cuda:
                torch.cuda.synchronize(obs.device)
            print(f"\nFused transformation + visibility estimation took: {1000 * (time() - t0)} msec")
            print(f"Point cloud size {self.points.size()}")
        return self.criterion(self.observations)

    def criterion(self, observations):
        # 1 / (sum + eps); the sum of the tensor produced by forward() was already reduced in-kernel
        if observations is self.observations and self._total is not None:
            total = self._total
        else:
            total = torch.sum(observations)
        return 1.0 / (total + self.eps)


# ------------------------------------------------------------------------------------------
# trajectory (src/model.py:135-260)
# ------------------------------------------------------------------------------------------
def length_calc(traj):
    """src/model.py:135-139 — sum of segment lengths (vectorised, same arithmetic per segment)."""
    if len(traj) < 2:
        return 0.0
    return torch.linalg.norm(traj[1:] - traj[:-1], dim=-1).sum()


def mean_angle_calc(traj_wps, eps=1e-6):
    """src/model.py:142-155 — mean interior angle at the waypoints."""
    n_wps = len(traj_wps)
    traj = torch.as_tensor(traj_wps).reshape(n_wps, -1)
    if n_wps <= 2:
        return 0.0 / (n_wps - 2)  # ZeroDivisionError for 2 waypoints, 0.0/-1 for one: as the reference
    ab = traj[:-2] - traj[1:-1]
    ac = traj[2:] - traj[1:-1]
    cosang = (ab * ac).sum(dim=1) / (torch.linalg.norm(ab, dim=1) * torch.linalg.norm(ac, dim=1) + eps)
    return torch.arccos(cosang).sum() / (n_wps - 2)


class ModelTraj(nn.Module):
    def __init__(self, points, wps_poses, wps_quats, intrins, img_width, img_height, min_dist=1.0, max_dist=5.0,
                 smoothness_weight=14.0, traj_length_weight=0.02, device=torch.device("cuda"), group=None,
                 n_total=None, spatial_sort=True, fused_regularizers=False):
        super().__init__()
        assert wps_poses.dim() == wps_quats.dim()
        assert wps_poses.size()[1] == 3
        assert wps_quats.size()[1] == 4

        self.device = device
        self.points = torch.as_tensor(points, dtype=torch.float32).to(self.device)
        self.rewards = None
        self.observations = None
        self.lo_sum = 0.0
        self.poses0 = torch.as_tensor(wps_poses, dtype=torch.float32).to(self.device)
        self.quats0 = torch.as_tensor(wps_quats, dtype=torch.float32).to(self.device)
        self.poses = nn.Parameter(deepcopy(self.poses0))
        self.quats = nn.Parameter(deepcopy(self.quats0))
        self.K = torch.as_tensor(intrins, dtype=torch.float32).to(self.device)
        self.img_width, self.img_height = float(img_width), float(img_height)
        self.eps = 1e-6
        self.pc_clip_limits = [min_dist, max_dist]
        self.loss = {"vis": float("inf"), "length": float("inf"), "l2": float("inf"), "smooth": float("inf")}
        self.smoothness_weight = smoothness_weight
        self.traj_length_weight = traj_length_weight
        self.group = group        # torch.distributed group when `points` is this rank's shard
        self.n_total = n_total    # global point count in that case
        # The cloud is constant over the optimisation (src/trajectory_optimization.py:83-127), so a Morton-ordered
        # copy is made once; the kernels prune whole tiles of it per pose (bit-identical results) and write
        # `rewards` back in the order of `self.points`.
        self.spatial_sort = spatial_sort
        # The l2 / smooth / length terms default to the torch-op restatement, which repeats the reference's fp32
        # arithmetic operation by operation.  `fused_regularizers=True` computes them and their gradients in one launch
        # in fp64 (cov_traj_regularizers): ~40 % less step latency on small clouds, but on nearly straight paths the
        # smooth-term gradient then differs from the reference's by up to ~1e-4 relative — arccos' is ill-conditioned
        # near -1 and the reference evaluates it in fp32.
        self.fused_regularizers = fused_regularizers
        self._mean = None
        self._step_cache = {}
        self._pts32 = None
        self._perm = None
        self._boxes = None
        self._ws = None
        self.to(self.device)

    def _cloud(self):
        # The ordered copy, its permutation and the tile boxes belong to ONE state of `self.points`: the cache is keyed
        # on identity and on torch's version counter, so both `model.points = new` and an in-place `model.points.copy_()`
        # trigger a re-sort (the exact pruning relies on the boxes bounding the points the kernels read).
        key = (id(self.points), self.points._version)
        if self._pts32 is None or self._pts32_key != key:
            pts = ops._BACKEND.prepare(self.points, what="points")
            self._pts32, self._perm, self._boxes = ops._BACKEND.order_cloud(pts, self.spatial_sort)
            self._pts32_key = key
        return self._pts32

    def set_points(self, points):
        """Replace the cloud; the Morton-ordered copy, permutation and boxes are rebuilt on the next forward.  Inside a
        captured CUDA graph use `refresh_points_()` instead (same shapes, buffers rewritten in place)."""
        self.points = torch.as_tensor(points, dtype=torch.float32).to(self.device)
        self._pts32 = None

    def detach_state(self):
        """Drop the references this model keeps into the last autograd graph (`loss[...]`, `_mean`) so that the graph —
        and the gradient-accumulation nodes bound to the stream it ran on — can be freed (graphs.GraphedStep calls this
        before it warms up on its capture stream)."""
        self._mean = None
        for k, v in list(self.loss.items()):
            if isinstance(v, torch.Tensor):
                self.loss[k] = v.detach()
        if self.rewards is not None:
            self.rewards = self.rewards.detach()

    @torch.no_grad()
    def refresh_points_(self, points):
        """In-place cloud update for CUDA-graph replays (graphs.GraphedStep): `points` must have the shape of the
        current cloud; `self.points`, the ordered copy, the permutation and the boxes keep their addresses and are
        rewritten, so a graph captured earlier evaluates the NEW cloud on its next replay."""
        if self._pts32 is None:
            self._cloud()
        if tuple(points.shape) != tuple(self.points.shape):
            raise ValueError("refresh_points_ needs a cloud of the same shape; use set_points() and re-capture")
        self.points.copy_(points)
        pts = ops._BACKEND.prepare(self.points, what="points")
        new_pts, new_perm, new_boxes = ops._BACKEND.order_cloud(pts, self.spatial_sort)
        if self._pts32.data_ptr() != self.points.data_ptr():
            self._pts32.copy_(new_pts)
        if self._perm is not None:
            self._perm.copy_(new_perm)
        if self._boxes is not None and new_boxes is not None:
            self._boxes.copy_(new_boxes)
        self._pts32_key = (id(self.points), self.points._version)

    def _workspace(self, cloud):
        # one workspace per (cloud, pose count), kept across steps: nothing in it outlives a call
        if not cloud.is_cuda:
            return None
        need = ops._BACKEND.traj_workspace_bytes(cloud, max(1, self.poses.shape[0]))
        if self._ws is None or self._ws.numel() < need or self._ws.device != cloud.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=cloud.device)
        return self._ws

    def _wps_step(self, vis_wps_dist):
        # src/model.py:214-215.  poses0 never changes, so the host sync happens once per distance
        key = float(vis_wps_dist)
        if key not in self._step_cache:
            mean_wps_dist = (self.poses0[1:, :] - self.poses0[:-1, :]).norm(dim=1).mean()
            self._step_cache[key] = int(vis_wps_dist / mean_wps_dist) + 1
        return self._step_cache[key]

    def forward(self, vis_wps_dist=0.5, debug=False):
        t0 = time()
        wps_step = self._wps_step(vis_wps_dist)
        # waypoints range(0, N_wps, wps_step) (src/model.py:217); the others get no visibility gradient
        cloud = self._cloud()
        rewards, mean = ops.coverage_traj(cloud, self.poses[::wps_step], self.quats[::wps_step], self.K,
                                          self.img_width, self.img_height, self.pc_clip_limits[0],
                                          self.pc_clip_limits[1], self.eps, n_total=self.n_total, group=self.group,
                                          reward_index=self._perm, boxes=self._boxes,
                                          dense=None if self.spatial_sort else True, workspace=self._workspace(cloud))
        self.rewards = rewards
        self._mean = mean
        if debug:
            if rewards.is_cuda:
                torch.cuda.synchronize(rewards.device)
            print(f"Trajectory evaluation took {1000 * (time() - t0)} msec")
        t1 = time()
        loss = self.criterion(self.rewards)
        if debug:
            print(f"Loss calculation took {1000 * (time() - t1)} msec")
        return loss

    def criterion(self, rewards):
        # src/model.py:244-260
        mean = self._mean if (rewards is self.rewards and self._mean is not None) else torch.mean(rewards)
        self.loss["vis"] = 1.0 / (mean + self.eps)
        if self.fused_regularizers and self.poses.is_cuda and self.poses.shape[0] >= 3:
            # the three O(W) regularisers and their gradients in one launch (same formulas, fp64 inside)
            reg = ops.traj_regularizers(self.poses, self.poses0, self.smoothness_weight, self.traj_length_weight, self.eps)
            self.loss["l2"], self.loss["smooth"], self.loss["length"] = reg[0], reg[1], reg[2]
            return self.loss["vis"] + reg.sum()
        return self._criterion_terms_torch()

    def _criterion_terms_torch(self):
        # the same terms as plain torch expressions (fewer than 3 waypoints keep the reference's corner cases; also
        # the restatement the fused kernel is tested against)
        self.loss["l2"] = torch.linalg.norm(self.poses[0] - self.poses0[0])
        self.loss["smooth"] = self.smoothness_weight / (mean_angle_calc(self.poses, self.eps) + self.eps)
        self.loss["length"] = self.traj_length_weight * torch.abs(length_calc(self.poses) - length_calc(self.poses0))
        return self.loss["vis"] + self.loss["l2"] + self.loss["length"] + self.loss["smooth"]
