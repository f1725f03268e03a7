"""Property tests: random poses, intrinsics and clip limits.

CPU part: the numpy oracle against the torch port of the reference (autograd) — two independent restatements of
src/model.py:13-57,98-127,200-246 must agree for ANY camera, not only the fixtures'.
GPU part: the CUDA kernels against the fp64 oracle under the same random draws (1e-4 relative, per-row 5e-4 on gradients)."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import coverage_oracle as orc, torch_port
from tests.conftest import rel_err, row_rel_err

finite = dict(allow_nan=False, allow_infinity=False)
camera = st.fixed_dictionaries(dict(
    fx=st.floats(300.0, 1200.0, **finite), fy=st.floats(300.0, 1200.0, **finite),
    cx=st.floats(200.0, 1000.0, **finite), cy=st.floats(200.0, 1200.0, **finite),
    width=st.floats(400.0, 2000.0, **finite), height=st.floats(400.0, 2000.0, **finite),
    # clip limits on a 1/8 grid: the reference builds mean and std as fp32 tensors from Python floats (src/model.py:20-21),
    # so only values whose half-sum and half-difference are exact in fp32 make its fp64 path comparable to 1e-9
    min_d=st.integers(2, 16).map(lambda k: k / 8.0), span=st.integers(4, 32).map(lambda k: k / 4.0)))
pose = st.fixed_dictionaries(dict(
    t=st.tuples(st.floats(-4.0, 12.0, **finite), st.floats(-4.0, 12.0, **finite), st.floats(-1.0, 2.0, **finite)),
    q=st.tuples(*[st.floats(-1.0, 1.0, **finite)] * 4).filter(lambda q: sum(v * v for v in q) > 0.05),
    scale=st.floats(0.3, 3.0, **finite)))


def _cloud(seed, n):
    g = np.random.default_rng(seed)
    return (g.random((n, 3)) * np.array([16, 16, 4]) + np.array([-4, -4, -1])).astype(np.float32)


def _K(c):
    return np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]], np.float32)


@settings(max_examples=25, deadline=None)
@given(cam=camera, p=pose, seed=st.integers(0, 2 ** 16))
def test_oracle_pose_objective_agrees_with_torch_autograd(cam, p, seed):
    pts = _cloud(seed, 600)
    K = _K(cam)
    t = np.array([p["t"]], np.float64)
    q = np.array([p["q"]], np.float64) * p["scale"]
    ref = orc.pose_objective(pts, t, q, K, cam["width"], cam["height"], cam["min_d"], cam["min_d"] + cam["span"], dtype=np.float64)
    T = torch.tensor(t, dtype=torch.float64, requires_grad=True)
    Q = torch.tensor(q, dtype=torch.float64, requires_grad=True)
    loss, obs = torch_port.pose_loss(torch.from_numpy(pts).double(), T, Q, torch.from_numpy(K).double(), cam["width"],
                                     cam["height"], cam["min_d"], cam["min_d"] + cam["span"])
    loss.backward()
    assert rel_err(obs.detach().numpy(), ref["obs"]) < 1e-9
    if float(ref["sum"]) > 1e-30:   # a camera that sees nothing has no gradient to compare
        assert rel_err(loss.item(), ref["loss"]) < 1e-9
        scale = max(np.abs(ref["g_trans"]).max(), np.abs(ref["g_quat"]).max(), 1e-300)
        assert np.abs(T.grad.numpy().ravel() - ref["g_trans"]).max() / scale < 1e-7
        assert np.abs(Q.grad.numpy().ravel() - ref["g_quat"]).max() / scale < 1e-7


@pytest.mark.gpu
@settings(max_examples=30, deadline=None)
@given(cam=camera, p=pose, seed=st.integers(0, 2 ** 16))
def test_cuda_pose_objective_matches_oracle_for_any_camera(cam, p, seed):
    from trajectory_optimization_b200 import ops
    dev = torch.device("cuda:0")
    pts = _cloud(seed, 3000)
    K = _K(cam)
    t = np.array([p["t"]], np.float32)
    q = (np.array([p["q"]], np.float64) * p["scale"]).astype(np.float32)
    mx = cam["min_d"] + cam["span"]
    ref = orc.pose_objective(pts, t, q, K, cam["width"], cam["height"], cam["min_d"], mx, dtype=np.float64)
    T = torch.from_numpy(t).to(dev).requires_grad_(True)
    Q = torch.from_numpy(q).to(dev).requires_grad_(True)
    obs, total = ops.coverage_pose(torch.from_numpy(pts).to(dev), T, Q, torch.from_numpy(K).to(dev), cam["width"],
                                   cam["height"], cam["min_d"], mx)
    total.backward()
    assert rel_err(obs.detach().cpu().numpy(), ref["obs"]) < 1e-4
    if float(ref["sum"]) > 1e-20:
        assert rel_err(total.item(), ref["sum"]) < 1e-4
        # d(sum)/d(pose) = -d(loss)/d(pose) / loss^2
        g_t = -ref["g_trans"] / float(ref["loss"]) ** 2
        g_q = -ref["g_quat"] / float(ref["loss"]) ** 2
        scale = max(np.abs(g_t).max(), np.abs(g_q).max())
        assert np.abs(T.grad.cpu().numpy().ravel() - g_t).max() / scale < 1e-4
        assert np.abs(Q.grad.cpu().numpy().ravel() - g_q).max() / scale < 1e-4


@pytest.mark.gpu
@settings(max_examples=20, deadline=None)
@given(cam=camera, poses=st.lists(pose, min_size=2, max_size=5), seed=st.integers(0, 2 ** 16), n=st.sampled_from([257, 4099, 70_001]))
def test_cuda_traj_objective_matches_oracle_for_any_camera(cam, poses, seed, n):
    """n = 70 001 takes the pruned pipeline (>= 65 536 points), the others the dense kernels."""
    from trajectory_optimization_b200 import ops
    dev = torch.device("cuda:0")
    pts = _cloud(seed, n)
    K = _K(cam)
    P = np.array([p["t"] for p in poses], np.float32)
    Qn = np.array([np.array(p["q"]) * p["scale"] for p in poses], np.float32)
    mx = cam["min_d"] + cam["span"]
    ref = orc.traj_objective(pts, P, Qn, K, cam["width"], cam["height"], cam["min_d"], mx, dtype=np.float64)
    if not np.isfinite(ref["vis"]):   # a pose that sees nothing: NaN on both sides (covered by a dedicated test)
        return
    Pt = torch.from_numpy(P).to(dev).requires_grad_(True)
    Qt = torch.from_numpy(Qn).to(dev).requires_grad_(True)
    rewards, mean = ops.coverage_traj(torch.from_numpy(pts).to(dev), Pt, Qt, torch.from_numpy(K).to(dev), cam["width"],
                                      cam["height"], cam["min_d"], mx)
    (1.0 / (mean + 1e-6)).backward()
    assert rel_err(rewards.detach().cpu().numpy(), ref["rewards"]) < 1e-4
    assert rel_err(1.0 / (mean.item() + 1e-6), ref["vis"]) < 1e-4
    # gradients: the fp32 maxima can tie differently from the fp64 ones only on degenerate draws; compare when the
    # fp32 oracle agrees with the fp64 one (well-conditioned normalisation)
    ref32 = orc.traj_objective(pts, P, Qn, K, cam["width"], cam["height"], cam["min_d"], mx, dtype=np.float32)
    if rel_err(ref32["g_poses"], ref["g_poses"]) < 1e-5 and np.abs(ref["g_poses"]).max() > 1e-12:
        assert rel_err(Pt.grad.cpu().numpy(), ref["g_poses"]) < 1e-4
        assert rel_err(Qt.grad.cpu().numpy(), ref["g_quats"]) < 1e-4
        assert row_rel_err(Pt.grad.cpu().numpy(), ref["g_poses"], floor=1e-3) < 5e-4
