"""Batched multi-camera front end: body waypoints (x, y, z, yaw) + fixed camera extrinsics ->
per-camera poses (t, q) for the fused coverage ops, differentiable in torch (tiny O(W) tensors).

The reference optimises camera poses directly (translation + quaternion, src/model.py:66-90); its
multi-camera setting comes from tf extrinsics of a 5-6 camera rig (src/pc_processor.py:33-39,161-165).
Here a rig is a list of (R_body_cam, t_body_cam); the autograd chain from the per-camera gradients
the kernels return back to the 4 body parameters is handled by torch on (W, 4)-sized tensors.
"""
import math

import torch

from . import ops

# body (x fwd, y left, z up) <- optical (x right, y down, z fwd)
R_BODY_OPTICAL = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


def ring_rig(n_cams=5, lever=(0.0, 0.0, 0.0)):
    """n cameras looking outwards at yaw offsets 0, +-360/n, +-2*360/n ... (deg) about body z."""
    rig = []
    for k in range(n_cams):
        step = (k + 1) // 2 * (1 if k % 2 else -1) if k else 0
        a = step * 2.0 * math.pi / n_cams
        Rz = torch.tensor([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])
        rig.append((Rz @ R_BODY_OPTICAL, torch.tensor(lever, dtype=torch.float32)))
    return rig


def matrix_to_quat_wxyz(R):
    """Batched rotation matrix -> unit quaternion (w, x, y, z), branch-free on the largest diagonal term."""
    m00, m11, m22 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    q_abs = torch.sqrt(torch.clamp(torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22,
                                                1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1), min=1e-12))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1),
        torch.stack([R[..., 2, 1] - R[..., 1, 2], q_abs[..., 1] ** 2, R[..., 1, 0] + R[..., 0, 1], R[..., 0, 2] + R[..., 2, 0]], -1),
        torch.stack([R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] + R[..., 0, 1], q_abs[..., 2] ** 2, R[..., 2, 1] + R[..., 1, 2]], -1),
        torch.stack([R[..., 1, 0] - R[..., 0, 1], R[..., 2, 0] + R[..., 0, 2], R[..., 2, 1] + R[..., 1, 2], q_abs[..., 3] ** 2], -1),
    ], -2) / (2.0 * q_abs[..., None])
    best = q_abs.argmax(-1)
    return torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)


def camera_poses_from_body(body_xyzyaw, rig):
    """body_xyzyaw (..., 4) -> (trans (..., C, 3), quat (..., C, 4)); camera->world rotation is
    Rz(yaw) @ R_body_cam, camera centre is body + Rz(yaw) @ t_body_cam."""
    xyz, yaw = body_xyzyaw[..., :3], body_xyzyaw[..., 3]
    c, s, o, l = torch.cos(yaw), torch.sin(yaw), torch.zeros_like(yaw), torch.ones_like(yaw)
    Rz = torch.stack([torch.stack([c, -s, o], -1), torch.stack([s, c, o], -1), torch.stack([o, o, l], -1)], -2)
    Rbc = torch.stack([r for r, _ in rig]).to(body_xyzyaw)      # (C,3,3)
    tbc = torch.stack([t for _, t in rig]).to(body_xyzyaw)      # (C,3)
    Rwc = Rz[..., None, :, :] @ Rbc                              # (...,C,3,3)
    twc = xyz[..., None, :] + (Rz[..., None, :, :] @ tbc[..., None]).squeeze(-1)
    return twc, matrix_to_quat_wxyz(Rwc)


def rig_tensor(rig, device):
    """(C,7) fp32 device tensor [q_body_cam (w,x,y,z), t_body_cam] for the fused front end."""
    R = torch.stack([r for r, _ in rig]).double()
    q = matrix_to_quat_wxyz(R)
    q = q / q.norm(dim=-1, keepdim=True)
    t = torch.stack([torch.as_tensor(t) for _, t in rig]).double()
    return torch.cat([q, t], -1).float().contiguous().to(device)


class RigPosesFn(torch.autograd.Function):
    """camera_poses_from_body as two O(W) CUDA launches (cov_rig_poses / cov_rig_poses_backward)."""

    @staticmethod
    def forward(ctx, body, rig7):
        if not body.is_cuda:
            raise RuntimeError("RigPosesFn is CUDA-only (use camera_poses_from_body for CPU tensors)")
        b = body.detach().contiguous().float()
        B, C = b.shape[0], rig7.shape[0]
        poses = torch.empty(B * C, 3, dtype=torch.float32, device=b.device)
        quats = torch.empty(B * C, 4, dtype=torch.float32, device=b.device)
        if rig7.device != b.device:
            raise RuntimeError(f"rig tensor is on {rig7.device}, body parameters on {b.device}")
        ops._call("cov_rig_poses", b, b.data_ptr(), B, rig7.data_ptr(), C, poses.data_ptr(), quats.data_ptr())
        ctx.save_for_backward(b, rig7)
        ctx.set_materialize_grads(False)
        return poses, quats

    @staticmethod
    def backward(ctx, g_poses, g_quats):
        b, rig7 = ctx.saved_tensors
        if g_poses is None and g_quats is None:
            return None, None
        gp = None if g_poses is None else g_poses.contiguous().float()
        gq = None if g_quats is None else g_quats.contiguous().float()
        out = torch.empty_like(b)
        ops._call("cov_rig_poses_backward", b, b.data_ptr(), b.shape[0], rig7.data_ptr(), rig7.shape[0],
                  0 if gp is None else gp.data_ptr(), 0 if gq is None else gq.data_ptr(), 1.0, out.data_ptr())
        return out, None


def camera_poses_fused(body_xyzyaw, rig7):
    """body (B,4) CUDA tensor, rig7 from `rig_tensor` -> (poses (B*C,3), quats (B*C,4)), differentiable in body.
    Same map as `camera_poses_from_body` up to the sign of each quaternion (which no consumer depends on)."""
    return RigPosesFn.apply(body_xyzyaw, rig7)
