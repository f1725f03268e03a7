"""Parity of the CUDA path (through the Python drop-in surface -> C ABI -> sm_100a kernels) with
(a) the fixtures produced by running the reference, (b) the CPU oracle on seeded inputs.

Tolerance: north_star asks for rewards and pose gradients within an FP32 relative tolerance of
1e-4 of the reference torch implementation.  Norm = max|a-b| / max|b| per array (conftest.rel_err).
The fp32 reference itself sits ~1e-6 from the fp64 truth, so we also require the CUDA result to
be within 1e-4 of the fp64 truth."""
import numpy as np
import pytest
import torch

from oracle import coverage_oracle as orc
from tests.conftest import load_golden, rel_err, row_rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
TOL_ROW = 2e-4   # per-pose gradient rows, each against its own magnitude (conftest.row_rel_err)
K_np, IMG_W, IMG_H = orc.load_intrinsics()


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def mod():
    from trajectory_optimization_b200 import model, tools, ops, _lib
    _lib.lib()  # fail loudly if the extension is not built
    return model, tools, ops


def _pts(g, sample_inputs):
    return g["in_points"] if "in_points" in g else sample_inputs["pts"]


POSE_CASES = ["pose_sample", "pose_synth0", "pose_synth1", "pose_synth2", "pose_synth_clip"]
TRAJ_CASES = ["traj_sample", "traj_sample_all", "traj_box", "traj_compact", "traj_tiny"]


@pytest.mark.parametrize("name", POSE_CASES)
def test_model_pose_matches_reference(name, dev, mod, sample_inputs):
    model, tools, ops = mod
    g = load_golden(name)
    K, W, H = tools.load_intrinsics(dev)
    pts = torch.from_numpy(_pts(g, sample_inputs))
    m = model.ModelPose(pts, torch.from_numpy(g["in_trans"]), torch.from_numpy(g["in_quat"]), K, W, H,
                        float(g["in_min_d"]), float(g["in_max_d"]), dev)
    if "in_weight" in g:
        obs, total = ops.coverage_pose(m.points, m.trans, m.quat, m.K, W, H, float(g["in_min_d"]),
                                       float(g["in_max_d"]), weight=torch.from_numpy(g["in_weight"]).to(dev))
        m.observations = obs
        loss = 1.0 / (total + m.eps)
    else:
        loss = m()
    loss.backward()
    for sfx in ("", "64"):
        assert rel_err(loss.item(), g["out_loss" + sfx]) < TOL
        assert rel_err(m.observations.detach().cpu().numpy(), g["out_obs" + sfx]) < TOL
        assert rel_err(m.trans.grad.detach().cpu().numpy(), g["out_g_trans" + sfx]) < TOL
        assert rel_err(m.quat.grad.detach().cpu().numpy(), g["out_g_quat" + sfx]) < TOL
    assert m.observations.shape == (len(pts),) and m.trans.grad.shape == (1, 3) and m.quat.grad.shape == (1, 4)


@pytest.mark.parametrize("name", TRAJ_CASES)
def test_model_traj_matches_reference(name, dev, mod, sample_inputs):
    model, tools, ops = mod
    g = load_golden(name)
    K, W, H = tools.load_intrinsics(dev)
    pts = torch.from_numpy(_pts(g, sample_inputs))
    m = model.ModelTraj(pts, torch.from_numpy(g["in_poses"]), torch.from_numpy(g["in_quats"]), K, W, H,
                        float(g["in_min_d"]), float(g["in_max_d"]), float(g["in_sw"]), float(g["in_lw"]), dev)
    loss = m(vis_wps_dist=float(g["in_vis_wps_dist"]))
    gv = torch.autograd.grad(m.loss["vis"], [m.poses, m.quats], retain_graph=True)
    loss.backward()
    assert rel_err(loss.item(), g["out_loss"]) < TOL
    for k in ("vis", "smooth"):
        assert rel_err(float(m.loss[k]), g["out_" + k]) < TOL
    assert float(m.loss["l2"]) == 0.0 and float(m.loss["length"]) == 0.0
    for sfx in ("", "64"):
        assert rel_err(float(m.loss["vis"]), g["out_vis" + sfx]) < TOL
        assert rel_err(m.rewards.detach().cpu().numpy(), g["out_rewards" + sfx]) < TOL
        assert rel_err(gv[0].detach().cpu().numpy(), g["out_gv_poses" + sfx]) < TOL
        assert rel_err(gv[1].detach().cpu().numpy(), g["out_gv_quats" + sfx]) < TOL
        assert row_rel_err(gv[0].detach().cpu().numpy(), g["out_gv_poses" + sfx]) < TOL_ROW
        assert row_rel_err(gv[1].detach().cpu().numpy(), g["out_gv_quats" + sfx]) < TOL_ROW
    assert rel_err(m.poses.grad.detach().cpu().numpy(), g["out_g_poses"]) < TOL
    assert rel_err(m.quats.grad.detach().cpu().numpy(), g["out_g_quats"]) < TOL
    step = int(g["out_wps_step"])
    if step > 1:  # skipped waypoints carry no visibility gradient (src/model.py:217)
        assert float(gv[1][1::step].abs().max()) == 0.0


def _box(gen, n, lo=(-10, -10, -1), hi=(30, 30, 4)):
    lo, hi = np.array(lo, np.float32), np.array(hi, np.float32)
    return (gen.random((n, 3), dtype=np.float32) * (hi - lo) + lo).astype(np.float32)


def _s_curve(W, L=12.0):
    xs = np.linspace(0, L, W)
    poses = np.stack([xs, 0.5 * xs + 0.3 * np.sin(xs), np.zeros(W)], 1).astype(np.float32)
    yaw = np.arctan2(0.5 + 0.3 * np.cos(xs), 1.0)
    return poses, yaw


@pytest.mark.parametrize("n,W,ragged", [(200_000, 40, False), (100_003, 7, True), (1_000_000, 24, False)])
def test_traj_matches_oracle_on_seeded_clouds(n, W, ragged, dev, mod):
    """CUDA vs the fp64 oracle at sizes the oracle finishes in seconds (exercises PPT=4/2/1 tiles)."""
    model, tools, ops = mod
    gen = np.random.default_rng(7 + n)
    pts = _box(gen, n)
    poses, _ = _s_curve(W)
    quats = (gen.standard_normal((W, 4)) * 0.5 + np.array([1.0, 0, 0, 0])).astype(np.float32)
    P = torch.from_numpy(poses).to(dev).requires_grad_(True)
    Q = torch.from_numpy(quats).to(dev).requires_grad_(True)
    K, Wd, Hd = tools.load_intrinsics(dev)
    rewards, mean = ops.coverage_traj(torch.from_numpy(pts).to(dev), P, Q, K, Wd, Hd)
    vis = 1.0 / (mean + 1e-6)
    vis.backward()
    ref = orc.traj_objective(pts, poses, quats, K_np, IMG_W, IMG_H, dtype=np.float64)
    assert rel_err(vis.item(), ref["vis"]) < TOL
    assert rel_err(rewards.detach().cpu().numpy(), ref["rewards"]) < TOL
    assert rel_err(P.grad.detach().cpu().numpy(), ref["g_poses"]) < TOL
    assert rel_err(Q.grad.detach().cpu().numpy(), ref["g_quats"]) < TOL
    assert row_rel_err(P.grad.detach().cpu().numpy(), ref["g_poses"]) < TOL_ROW
    assert row_rel_err(Q.grad.detach().cpu().numpy(), ref["g_quats"]) < TOL_ROW


def test_config3_size_against_the_oracle_on_a_point_subsample(dev, mod):
    """BASELINE config 3 (1e7 points x 100 poses) against the oracle: the oracle cannot hold 1e9 pairs, but with the
    GLOBAL normalisers passed in (`minmax=`) its per-point rewards and its gradient accumulators are sums over points,
    so a random subsample of 50 000 points is checked exactly: rewards on the subsample, and the subsample's share of the
    accumulators against a CUDA evaluation of the same subsample with the same normalisers."""
    model, tools, ops = mod
    from trajectory_optimization_b200 import _lib
    n, W = 10_000_000, 100
    g = torch.Generator(device=dev).manual_seed(11)
    lo = torch.tensor([-10.0, -10.0, -1.0], device=dev)
    hi = torch.tensor([30.0, 30.0, 4.0], device=dev)
    pts = torch.rand(n, 3, generator=g, device=dev) * (hi - lo) + lo
    poses_np, _ = _s_curve(W, L=30.0)
    gen = np.random.default_rng(5)
    quats_np = (gen.standard_normal((W, 4)) * 0.3 + np.array([1.0, 0, 0, 0])).astype(np.float32)
    P = torch.from_numpy(poses_np).to(dev).requires_grad_(True)
    Q = torch.from_numpy(quats_np).to(dev).requires_grad_(True)
    K, Wd, Hd = tools.load_intrinsics(dev)
    m = model.ModelTraj(pts, P.detach(), Q.detach(), K, Wd, Hd, device=dev)   # Morton-ordered, pruned: the product path
    m(vis_wps_dist=0.0)
    m.loss["vis"].backward()
    rewards = m.rewards.detach()
    # global normalisers from the raw C ABI (pass A on the whole cloud)
    cam = _lib.camera(Wd, Hd, 1.0, 5.0, 1e-6)
    minmax = ops._BACKEND.traj_minmax(ops._dev_f32(pts), P.detach(), Q.detach(), K.reshape(9), cam)
    mm = minmax.cpu().numpy().astype(np.float64)
    idx = torch.randperm(n, generator=torch.Generator().manual_seed(3))[:50_000].sort().values
    sub = pts[idx.to(dev)].cpu().numpy()
    ref = orc.traj_partials(sub, poses_np, quats_np, K_np, IMG_W, IMG_H, minmax=(mm[:W], mm[W:]), dtype=np.float64)
    assert rel_err(rewards[idx.to(dev)].cpu().numpy(), ref["rewards"]) < TOL
    assert float(rewards.min()) >= 0.5 and float(rewards.max()) <= 1.0
    # mean reward of the whole cloud vs the subsample's (statistical, 5e4 samples): sanity only
    assert abs(float(rewards.mean()) - float(ref["rewards"].mean())) < 5e-3
    # the subsample's share of the gradient: CUDA pass B + epilogue on the subsample with the GLOBAL normalisers vs the
    # oracle's accumulators for the same points (the global arg-max points are not in the subsample: no tie terms)
    sub_dev = torch.from_numpy(sub).to(dev)
    n_sub = len(sub)
    acc = ops._BACKEND.traj_fused(sub_dev, P.detach(), Q.detach(), K.reshape(9), cam, minmax, None,
                                  torch.empty(n_sub, device=dev))
    out = ops._BACKEND.traj_epilogue(acc, minmax, Q.detach(), n_sub, 0).cpu().numpy().astype(np.float64)
    refg = orc.traj_grads_from_partials(ref, poses_np, quats_np, n_sub)
    assert rel_err(out[0], refg["mean"]) < 1e-6
    gp = -refg["vis"] ** 2 * out[1:1 + 3 * W].reshape(W, 3)
    gq = -refg["vis"] ** 2 * out[1 + 3 * W:].reshape(W, 4)
    assert rel_err(gp, refg["g_poses"]) < TOL and rel_err(gq, refg["g_quats"]) < TOL
    assert row_rel_err(gp, refg["g_poses"]) < TOL_ROW and row_rel_err(gq, refg["g_quats"]) < TOL_ROW


def test_pose_matches_oracle_large_and_ragged(dev, mod):
    model, tools, ops = mod
    gen = np.random.default_rng(3)
    for n in (1, 2, 3, 5, 1027, 1_000_001):
        pts = _box(gen, n, (-4, -4, -1), (8, 8, 4))
        t = np.array([[0.4, -0.3, 0.2]], np.float32)
        q = np.array([[0.9, 0.1, -0.2, 0.3]], np.float32)
        T = torch.from_numpy(t).to(dev).requires_grad_(True)
        Q = torch.from_numpy(q).to(dev).requires_grad_(True)
        K, Wd, Hd = tools.load_intrinsics(dev)
        obs, total = ops.coverage_pose(torch.from_numpy(pts).to(dev), T, Q, K, Wd, Hd)
        (1.0 / (total + 1e-6)).backward()
        ref = orc.pose_objective(pts, t, q, K_np, IMG_W, IMG_H, dtype=np.float64)
        assert rel_err(obs.detach().cpu().numpy(), ref["obs"]) < TOL
        assert rel_err(total.item(), ref["sum"]) < TOL
        assert rel_err(T.grad.detach().cpu().numpy().ravel(), ref["g_trans"]) < TOL
        assert rel_err(Q.grad.detach().cpu().numpy().ravel(), ref["g_quat"]) < TOL


def test_general_backward_through_per_point_outputs(dev, mod):
    """Differentiating through observations / rewards (not the fused scalar) takes the upstream-weighted path."""
    model, tools, ops = mod
    g = load_golden("traj_box")
    K, W, H = tools.load_intrinsics(dev)
    pts = torch.from_numpy(g["in_points"]).to(dev)
    P = torch.from_numpy(g["in_poses"][::2].copy()).to(dev).requires_grad_(True)
    Q = torch.from_numpy(g["in_quats"][::2].copy()).to(dev).requires_grad_(True)
    rewards, mean = ops.coverage_traj(pts, P, Q, K, W, H)
    g_fused = torch.autograd.grad(mean, [P, Q], retain_graph=True)
    g_gen = torch.autograd.grad(rewards.mean(), [P, Q])
    assert rel_err(g_gen[0].detach().cpu().numpy(), g_fused[0].detach().cpu().numpy()) < 1e-5
    assert rel_err(g_gen[1].detach().cpu().numpy(), g_fused[1].detach().cpu().numpy()) < 1e-5
    g2 = load_golden("pose_synth2")
    T = torch.from_numpy(g2["in_trans"]).to(dev).requires_grad_(True)
    Qp = torch.from_numpy(g2["in_quat"]).to(dev).requires_grad_(True)
    obs, total = ops.coverage_pose(torch.from_numpy(g2["in_points"]).to(dev), T, Qp, K, W, H)
    a = torch.autograd.grad(total, [T, Qp], retain_graph=True)
    b = torch.autograd.grad(obs.sum(), [T, Qp])
    assert rel_err(b[0].detach().cpu().numpy(), a[0].detach().cpu().numpy()) < 1e-5 and rel_err(b[1].detach().cpu().numpy(), a[1].detach().cpu().numpy()) < 1e-5


def test_sharded_accumulators_equal_whole_cloud(dev, mod):
    """Size-independent property at a BASELINE-scale cloud (config 3: 1e7 points, 100 poses): evaluating two
    point shards with shared (MIN/MAX-reduced) normalisers and summing the accumulators — exactly what the
    multi-GPU path does — reproduces the single-shot result; rewards lie in [0.5, 1]."""
    import ctypes
    from trajectory_optimization_b200 import _lib
    model, tools, ops = mod
    L = _lib.lib()
    n, W = 10_000_000, 100
    gen = torch.Generator(device="cpu").manual_seed(0)
    lo, hi = torch.tensor([-10.0, -10, -1]), torch.tensor([30.0, 30, 4])
    pts = (torch.rand(n, 3, generator=gen) * (hi - lo) + lo).to(dev)
    poses_np, yaw = _s_curve(20, 19.0)
    offs = np.deg2rad([0, 72, -72, 144, -144])
    poses = np.repeat(poses_np, 5, axis=0)
    ang = (yaw[:, None] + offs[None, :]).reshape(-1)
    quats = np.stack([np.cos(ang / 2), np.zeros_like(ang), np.zeros_like(ang), np.sin(ang / 2)], 1).astype(np.float32)
    P = torch.from_numpy(poses).to(dev)
    Q = torch.from_numpy(quats).to(dev)
    K, Wd, Hd = tools.load_intrinsics(dev)
    cam = _lib.camera(Wd, Hd, 1.0, 5.0, 1e-6)
    Pg, Qg = P.clone().requires_grad_(True), Q.clone().requires_grad_(True)
    rewards, mean = ops.coverage_traj(pts, Pg, Qg, K, Wd, Hd)
    gp, gq = torch.autograd.grad(mean, [Pg, Qg])
    assert float(rewards.min()) >= 0.5 and float(rewards.max()) <= 1.0
    # two shards through the C ABI, reduced by hand
    cut = 3_333_332  # multiple of 4 keeps the second shard 16-byte aligned
    shards = [pts[:cut], pts[cut:]]
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    mm = []
    for s in shards:
        t = torch.empty(2 * W, device=dev)
        wsb = L.cov_traj_workspace_bytes(s.shape[0], W)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(L.cov_traj_minmax(s.data_ptr(), s.shape[0], P.data_ptr(), Q.data_ptr(), W, K.data_ptr(),
                                     ctypes.byref(cam), None, t.data_ptr(), None, ws.data_ptr(), wsb, stream), "minmax")
        mm.append(t)
    minmax = torch.cat([torch.minimum(mm[0][:W], mm[1][:W]), torch.maximum(mm[0][W:], mm[1][W:])])
    acc = torch.zeros(W * _lib.ACC_STRIDE + 1, dtype=torch.float64, device=dev)
    rew = []
    for s in shards:
        a = torch.empty_like(acc)
        r = torch.empty(s.shape[0], device=dev)
        wsb = L.cov_traj_workspace_bytes(s.shape[0], W)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(L.cov_traj_fused(s.data_ptr(), s.shape[0], P.data_ptr(), Q.data_ptr(), W, K.data_ptr(),
                                    ctypes.byref(cam), None, minmax.data_ptr(), None, None, r.data_ptr(), a.data_ptr(),
                                    None, ws.data_ptr(), wsb, stream), "fused")
        acc += a
        rew.append(r)
    out = torch.empty(1 + 7 * W, device=dev)
    _lib.check(L.cov_traj_epilogue(acc.data_ptr(), minmax.data_ptr(), Q.data_ptr(), W, n, 0, out.data_ptr(), stream),
               "epilogue")
    assert torch.equal(torch.cat(rew), rewards)  # per-point results do not depend on the sharding
    assert rel_err(out[0].item(), mean.item()) < 1e-6
    assert rel_err(out[1:1 + 3 * W].detach().cpu().numpy(), gp.reshape(-1).detach().cpu().numpy()) < 1e-5
    assert rel_err(out[1 + 3 * W:].detach().cpu().numpy(), gq.reshape(-1).detach().cpu().numpy()) < 1e-5
    # run to run: per-point outputs are bitwise reproducible; the accumulators are sums of the same fp32 block partials
    # added with fp64 atomics, so only the order of ~1e-16-relative roundings can differ
    rewards2, mean2 = ops.coverage_traj(pts, Pg, Qg, K, Wd, Hd)
    gp2, gq2 = torch.autograd.grad(mean2, [Pg, Qg])
    assert torch.equal(rewards2, rewards) and rel_err(mean2.item(), mean.item()) < 1e-7
    assert rel_err(gp2.cpu().numpy(), gp.cpu().numpy()) < 1e-6 and rel_err(gq2.cpu().numpy(), gq.cpu().numpy()) < 1e-6


def test_sweep_matches_per_trajectory_forward(dev, mod):
    model, tools, ops = mod
    gen = np.random.default_rng(11)
    pts = torch.from_numpy(_box(gen, 300_000)).to(dev)
    T, Pn = 37, 8
    base, yaw = _s_curve(Pn, 10.0)
    poses = base[None] + gen.normal(0, 1.0, (T, 1, 3)).astype(np.float32) * np.array([1, 1, 0], np.float32)
    ang = yaw[None, :] + gen.normal(0, 0.3, (T, Pn))
    quats = np.stack([np.cos(ang / 2), 0 * ang, 0 * ang, np.sin(ang / 2)], -1).astype(np.float32)
    K, Wd, Hd = tools.load_intrinsics(dev)
    from trajectory_optimization_b200 import _lib
    means = ops.sweep_rewards(pts, torch.from_numpy(poses.astype(np.float32)), torch.from_numpy(quats), K, Wd, Hd)
    try:  # the dense sweep (pruning off) and the pruned pipeline agree; so does the pruned one on the cloud as given
        ops._MODE.dense = not 0
        means_dense = ops.sweep_rewards(pts, torch.from_numpy(poses.astype(np.float32)), torch.from_numpy(quats), K, Wd, Hd)
    finally:
        ops._MODE.dense = not 1
    means_unsorted = ops.sweep_rewards(pts, torch.from_numpy(poses.astype(np.float32)), torch.from_numpy(quats), K, Wd, Hd,
                                       presorted=True)
    assert rel_err(means.cpu().numpy(), means_dense.cpu().numpy()) < 1e-9
    assert rel_err(means_unsorted.cpu().numpy(), means_dense.cpu().numpy()) < 1e-9
    for t in range(0, T, 6):
        _, mean = ops.coverage_traj(pts, torch.from_numpy(poses[t].astype(np.float32)).to(dev),
                                    torch.from_numpy(quats[t]).to(dev), K, Wd, Hd)
        assert rel_err(means[t].item(), mean.item()) < 1e-6


def test_errors_are_loud(dev, mod):
    model, tools, ops = mod
    K, Wd, Hd = tools.load_intrinsics(dev)
    pts = torch.rand(100, 3)
    with pytest.raises(RuntimeError):  # CPU tensors: no fallback
        ops.coverage_pose(pts, torch.zeros(1, 3), torch.tensor([[1.0, 0, 0, 0]]), K.cpu(), Wd, Hd)
    from trajectory_optimization_b200 import _lib
    too_many = _lib.lib().cov_traj_max_poses() + 1
    with pytest.raises(RuntimeError, match="exceed"):
        ops.coverage_traj(pts.to(dev), torch.zeros(too_many, 3, device=dev),
                          torch.tensor([[1.0, 0, 0, 0]], device=dev).repeat(too_many, 1), K, Wd, Hd)
    with pytest.raises(AssertionError):
        model.ModelPose(pts, torch.zeros(3), torch.tensor([[1.0, 0, 0, 0]]), K, Wd, Hd, device=dev)
    # a misaligned view is handled (cloned), not rejected
    base = torch.rand(101, 3, device=dev)
    obs, total = ops.coverage_pose(base[1:], torch.zeros(1, 3, device=dev), torch.tensor([[1.0, 0, 0, 0]], device=dev),
                                   K, Wd, Hd)
    assert obs.shape == (100,)


def test_pruned_evaluation_is_bit_identical_to_dense(dev, mod):
    """The bound-based pruning of (point, pose) pairs must not change a single bit of any output."""
    from trajectory_optimization_b200 import _lib
    model, tools, ops = mod
    L = _lib.lib()
    gen = np.random.default_rng(5)
    pts = torch.from_numpy(_box(gen, 3_000_017)).to(dev)
    poses, yaw = _s_curve(33, 14.0)
    quats = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], 1).astype(np.float32)
    quats += gen.normal(0, 0.2, quats.shape).astype(np.float32)
    K, Wd, Hd = tools.load_intrinsics(dev)
    outs = []
    try:
        for mode in (1, 0):
            ops._MODE.dense = not mode
            P = torch.from_numpy(poses).to(dev).requires_grad_(True)
            Q = torch.from_numpy(quats).to(dev).requires_grad_(True)
            rewards, mean = ops.coverage_traj(pts, P, Q, K, Wd, Hd)
            gp, gq = torch.autograd.grad(mean, [P, Q])
            outs.append((rewards, mean, gp, gq))
    finally:
        ops._MODE.dense = not 1
    (r1, m1, gp1, gq1), (r0, m0, gp0, gq0) = outs
    assert torch.equal(r1, r0) and torch.equal(m1, m0)           # per-point outputs and their mean: every bit
    # the gradient accumulators see the same addends in a different fp32 summation order
    assert rel_err(gp1.cpu().numpy(), gp0.cpu().numpy()) < 2e-6 and rel_err(gq1.cpu().numpy(), gq0.cpu().numpy()) < 2e-6
    assert float((r1 != 0.5).float().mean()) > 1e-4  # the case does exercise gated pairs


def test_spatial_sort_is_a_stable_morton_permutation(dev, mod):
    model, tools, ops = mod
    gen = np.random.default_rng(3)
    for n in (1, 5, 1000, 250_007, 3_000_001):   # the last: several 4096-key tiles per block and a ragged end
        pts_np = _box(gen, n)
        if n >= 1000:
            pts_np[10] = pts_np[500]  # duplicate points keep their input order (stable sort)
        pts = torch.from_numpy(pts_np).to(dev)
        out, perm = ops.spatial_sort(pts)
        assert perm.dtype == torch.int32 and out.shape == pts.shape
        assert torch.equal(torch.sort(perm.long()).values, torch.arange(n, device=dev))
        assert torch.equal(out, pts[perm.long()])
        # restate the key on the host: 10 bits per axis on a cubic grid over the bounding box, z-y-x interleave
        lo = pts_np.min(0)
        ext = np.float32((pts_np.max(0) - lo).max())
        scale = np.float32(1023.999) / ext if ext > 0 else np.float32(0)
        cell = np.clip(((pts_np - lo) * scale), 0, 1023).astype(np.uint32)
        key = np.zeros(n, np.uint64)
        for b in range(10):
            for a in range(3):
                key |= ((cell[:, a].astype(np.uint64) >> b) & 1) << (3 * b + a)
        order = np.argsort(key, kind="stable")
        assert np.array_equal(order, perm.cpu().numpy())


@pytest.mark.parametrize("n,bits", [(1, (0, 32)), (31, (0, 32)), (4096, (0, 32)), (4097, (0, 32)), (100_003, (0, 32)),
                                    (100_003, (5, 19)), (2_500_001, (0, 30)), (5_000_000, (0, 8)), (5_000_000, (24, 32))])
def test_sort_pairs_is_a_stable_unsigned_radix_sort(n, bits, dev, mod):
    """The library's own pair sort (csrc/cov_radix.cuh) against numpy's stable argsort on the selected key bits."""
    model, tools, ops = mod
    gen = np.random.default_rng(n + bits[0])
    keys = gen.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        keys[::7] = keys[3]            # many equal keys: their values must keep the input order
        keys[5::11] &= 0xff            # small keys: high digits all zero
    vals = np.arange(n, dtype=np.int32)
    k, v = ops.sort_pairs(torch.from_numpy(keys.view(np.int32)).to(dev), torch.from_numpy(vals).to(dev), *bits)
    digit = (keys.astype(np.uint64) >> bits[0]) & ((1 << (bits[1] - bits[0])) - 1)
    order = np.argsort(digit, kind="stable")
    assert np.array_equal(v.cpu().numpy(), vals[order])
    assert np.array_equal(k.cpu().numpy().view(np.uint32), keys[order])


def test_model_traj_sorted_cloud_equals_caller_order(dev, mod):
    """ModelTraj on its Morton-ordered copy returns rewards in the caller's point order, bit-identical to the
    evaluation of the cloud as given, and the same objective and gradients."""
    model, tools, ops = mod
    gen = np.random.default_rng(8)
    pts = torch.from_numpy(_box(gen, 700_003))
    poses, yaw = _s_curve(24, 12.0)
    quats = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], 1).astype(np.float32)
    quats += gen.normal(0, 0.1, quats.shape).astype(np.float32)
    K, Wd, Hd = tools.load_intrinsics(dev)
    res = []
    for flag in (True, False):
        m = model.ModelTraj(pts, torch.from_numpy(poses), torch.from_numpy(quats), K, Wd, Hd, device=dev, spatial_sort=flag)
        loss = m(vis_wps_dist=0.0)
        loss.backward()
        res.append((m.rewards.detach().clone(), loss.detach().clone(), m.poses.grad.clone(), m.quats.grad.clone()))
    (r1, l1, gp1, gq1), (r0, l0, gp0, gq0) = res
    assert torch.equal(r1, r0)
    assert rel_err(l1.item(), l0.item()) < 1e-6
    assert rel_err(gp1.cpu().numpy(), gp0.cpu().numpy()) < 5e-6 and rel_err(gq1.cpu().numpy(), gq0.cpu().numpy()) < 5e-6
    # differentiating through the per-point vector (upstream gradient in the caller's order) also goes through the permutation
    m = model.ModelTraj(pts, torch.from_numpy(poses), torch.from_numpy(quats), K, Wd, Hd, device=dev)
    m(vis_wps_dist=0.0)
    wgt = torch.linspace(0.5, 1.5, pts.shape[0], device=dev)
    ga = torch.autograd.grad((m.rewards * wgt).sum(), [m.poses, m.quats])
    m0 = model.ModelTraj(pts, torch.from_numpy(poses), torch.from_numpy(quats), K, Wd, Hd, device=dev, spatial_sort=False)
    m0(vis_wps_dist=0.0)
    gb = torch.autograd.grad((m0.rewards * wgt).sum(), [m0.poses, m0.quats])
    assert rel_err(ga[0].cpu().numpy(), gb[0].cpu().numpy()) < 5e-6 and rel_err(ga[1].cpu().numpy(), gb[1].cpu().numpy()) < 5e-6


def test_tile_pruning_on_sorted_cloud_is_bit_identical_to_dense(dev, mod):
    """Where the tile-level pruning actually bites (a Morton-ordered cloud): normalisers, rewards and their mean
    must not change by a bit against the dense evaluation; compact-cloud poses (min > 0) are never pruned."""
    import ctypes
    from trajectory_optimization_b200 import _lib
    model, tools, ops = mod
    L = _lib.lib()
    gen = np.random.default_rng(21)
    K, Wd, Hd = tools.load_intrinsics(dev)
    cam = _lib.camera(Wd, Hd, 1.0, 5.0, 1e-6)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for n, compact in ((2_000_003, False), (300_001, True)):
        if compact:  # cloud in front of every camera and within a few metres: min_j m > 0 for all poses
            pts_np = _box(gen, n, (1.0, 1.0, 1.5), (4.0, 4.0, 4.5))
            poses = (gen.random((12, 3), dtype=np.float32) * 0.6 - 0.3).astype(np.float32)
            quats = (np.array([1.0, 0, 0, 0]) + gen.normal(0, 0.15, (12, 4))).astype(np.float32)
        else:
            pts_np = _box(gen, n)
            poses, yaw = _s_curve(40, 16.0)
            quats = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], 1).astype(np.float32)
        pts, perm = ops.spatial_sort(torch.from_numpy(pts_np).to(dev))
        boxes = ops.tile_boxes(pts)
        P, Q = torch.from_numpy(poses).to(dev), torch.from_numpy(quats).to(dev)
        W = P.shape[0]
        outs = []
        try:
            for mode in (1, 0):
                ops._MODE.dense = not mode
                mm = torch.empty(2 * W, device=dev)
                wsb = L.cov_traj_workspace_bytes(n, W)
                ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
                _lib.check(L.cov_traj_minmax(pts.data_ptr(), n, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(),
                                             ctypes.byref(cam), boxes.data_ptr(), mm.data_ptr(),
                                             ctypes.byref(_lib.traj_opts(dense=not mode)), ws.data_ptr(), wsb, stream), "minmax")
                Pg, Qg = P.clone().requires_grad_(True), Q.clone().requires_grad_(True)
                rewards, mean = ops.coverage_traj(pts, Pg, Qg, K, Wd, Hd, reward_index=perm, boxes=boxes)
                gp, gq = torch.autograd.grad(mean, [Pg, Qg])
                outs.append((mm, rewards, mean, gp, gq))
        finally:
            ops._MODE.dense = not 1
        (mm1, r1, m1, gp1, gq1), (mm0, r0, m0, gp0, gq0) = outs
        assert torch.equal(mm1, mm0) and torch.equal(r1, r0) and torch.equal(m1, m0)
        assert rel_err(gp1.cpu().numpy(), gp0.cpu().numpy()) < 2e-6 and rel_err(gq1.cpu().numpy(), gq0.cpu().numpy()) < 2e-6
        if compact:
            assert int((mm0[:W] > 0).sum()) >= 6  # most of these poses keep min_j m > 0 (never pruned), a few do not


def test_fused_rig_front_end_matches_torch_chain(dev, mod):
    """cov_rig_poses / cov_rig_poses_backward against the differentiable torch restatement of the same map
    (body (x, y, z, yaw) o extrinsics -> camera poses), through the trajectory objective to the 4 body parameters."""
    from trajectory_optimization_b200 import multicam
    model, tools, ops = mod
    gen = np.random.default_rng(17)
    pts = torch.from_numpy(_box(gen, 200_000)).to(dev)
    K, Wd, Hd = tools.load_intrinsics(dev)
    rig = multicam.ring_rig(5, lever=(0.3, -0.1, 0.2))
    rig7 = multicam.rig_tensor(rig, dev)
    poses, yaw = _s_curve(9, 8.0)
    body0 = torch.from_numpy(np.concatenate([poses, yaw[:, None].astype(np.float32)], 1)).to(dev)
    grads, losses = [], []
    for fused in (True, False):
        body = body0.clone().requires_grad_(True)
        if fused:
            t, q = multicam.camera_poses_fused(body, rig7)
        else:
            t, q = multicam.camera_poses_from_body(body, rig)
            t, q = t.reshape(-1, 3), q.reshape(-1, 4)
        if fused:
            t_f, q_f = t.detach().clone(), q.detach().clone()
        else:
            assert rel_err(t_f.cpu().numpy(), t.detach().cpu().numpy()) < 1e-6
            sign = torch.sign((q_f * q.detach()).sum(-1, keepdim=True))   # q and -q are the same rotation
            assert rel_err((q_f * sign).cpu().numpy(), q.detach().cpu().numpy()) < 1e-6
        _, mean = ops.coverage_traj(pts, t, q, K, Wd, Hd)
        loss = 1.0 / (mean + 1e-6) + 0.01 * (t ** 2).sum()   # exercises both gradient inputs of the backward kernel
        loss.backward()
        grads.append(body.grad.clone())
        losses.append(loss.item())
    assert rel_err(losses[0], losses[1]) < 1e-6
    assert rel_err(grads[0].cpu().numpy(), grads[1].cpu().numpy()) < 1e-5


def test_graphed_optimisation_step_matches_eager(dev, mod):
    """One CUDA graph per optimisation step (zero_grad + forward + backward + Adam) follows the eager trajectory."""
    from trajectory_optimization_b200.graphs import GraphedStep
    model, tools, ops = mod
    gen = np.random.default_rng(4)
    pts = torch.from_numpy(_box(gen, 150_000))
    K, Wd, Hd = tools.load_intrinsics(dev)
    poses, yaw = _s_curve(12, 9.0)
    quats = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], 1).astype(np.float32)
    finals = []
    for graphed in (False, True):
        m = model.ModelTraj(pts, torch.from_numpy(poses), torch.from_numpy(quats), K, Wd, Hd, device=dev)
        opt = torch.optim.Adam([{"params": [m.poses], "lr": 0.02}, {"params": [m.quats], "lr": 0.01}], capturable=True)
        if graphed:
            # one eager step on the default stream first: the autograd graph it leaves behind (model.loss[...]) must not
            # tie the capture to the default stream (GraphedStep detaches the model's state)
            opt.zero_grad()
            m().backward()
            opt.step()
            g = GraphedStep(m, opt, warmup=1)        # 1 more eager warm-up step, then 3 replays
            for _ in range(3):
                loss = g.step()
        else:
            for _ in range(5):
                opt.zero_grad()
                loss = m()
                loss.backward()
                opt.step()
        torch.cuda.synchronize()
        finals.append((loss.item(), m.poses.detach().clone(), m.quats.detach().clone()))
    assert rel_err(finals[1][0], finals[0][0]) < 1e-5
    assert rel_err(finals[1][1].cpu().numpy(), finals[0][1].cpu().numpy()) < 1e-5
    assert rel_err(finals[1][2].cpu().numpy(), finals[0][2].cpu().numpy()) < 1e-5
    mp = model.ModelPose(pts, torch.tensor([[6.0, 2.0, 0.0]]), torch.tensor([[0.92, 0.0, 0.0, 0.39]]), K, Wd, Hd, device=dev)
    optp = torch.optim.Adam([{"params": [mp.trans], "lr": 0.02}, {"params": [mp.quat], "lr": 0.02}], capturable=True)
    gp = GraphedStep(mp, optp)
    l0 = gp.step().item()
    for _ in range(20):
        l1 = gp.step().item()
    assert l1 < l0   # the optimiser is making progress on 1/(sum of observations)


@pytest.mark.parametrize("n,W,mode", [(1, 1, "box"), (31, 7, "box"), (129, 33, "box"), (4097, 100, "box"), (70_001, 500, "box"),
                                      (300_000, 64, "far"), (200_003, 40, "mixed"), (1_000_001, 1100, "box"),
                                      (70_003, -1, "box"), (66_000, -2, "box")])
def test_pruned_pipeline_equals_dense_on_edge_shapes(n, W, mode, dev, mod):
    """Cull -> work list -> tiles kernels against the dense kernels on ragged sizes, one pose, more poses than one
    mask word / one slot round, clouds that are almost entirely out of reach and clouds with a far-away half."""
    from trajectory_optimization_b200 import _lib
    model, tools, ops = mod
    L = _lib.lib()
    if W < 0:   # the largest pose tables: 1 and 2 points per thread, one block per SM, several mask words and slot rounds
        W = L.cov_traj_max_poses() if W == -1 else (L.cov_traj_max_poses() * 3) // 4
    gen = np.random.default_rng(n + W)
    pts_np = _box(gen, n)
    if mode == "far":
        pts_np[100:] += np.float32(500.0)           # all but 100 points are out of reach of every pose
    elif mode == "mixed":
        pts_np[: n // 2] += np.float32(300.0)
    poses = (gen.random((W, 3), dtype=np.float32) * np.array([30, 30, 2], np.float32) + np.array([-5, -5, -0.5], np.float32))
    quats = gen.normal(0, 1, (W, 4)).astype(np.float32)
    K, Wd, Hd = tools.load_intrinsics(dev)
    outs = []
    try:
        for prune in (1, 0):
            ops._MODE.dense = not prune
            for ordered in ((True, False) if prune else (False,)):
                pts = torch.from_numpy(pts_np).to(dev)
                perm = boxes = None
                if ordered:
                    pts, perm = ops.spatial_sort(pts)
                    boxes = ops.tile_boxes(pts)
                P = torch.from_numpy(poses).to(dev).requires_grad_(True)
                Q = torch.from_numpy(quats).to(dev).requires_grad_(True)
                rewards, mean = ops.coverage_traj(pts, P, Q, K, Wd, Hd, reward_index=perm, boxes=boxes)
                gp, gq = torch.autograd.grad(mean, [P, Q])
                outs.append((rewards, mean, gp, gq))
    finally:
        ops._MODE.dense = not 1
    dense = outs[-1]
    if bool(torch.isnan(dense[1])):   # one point: min == max, the reference's normalisation is 0/0 there too
        assert n == 1 and all(bool(torch.isnan(got[0]).all()) and bool(torch.isnan(got[1])) for got in outs)
        return
    for got in outs[:-1]:
        assert torch.equal(got[0], dense[0])
        assert rel_err(got[1].item(), dense[1].item()) < 1e-12
        scale_p, scale_q = float(dense[2].abs().max()), float(dense[3].abs().max())
        if scale_p > 0:
            assert rel_err(got[2].cpu().numpy(), dense[2].cpu().numpy()) < 5e-6
        if scale_q > 0:
            assert rel_err(got[3].cpu().numpy(), dense[3].cpu().numpy()) < 5e-6
    if mode == "far":
        assert float(dense[0][100:].min()) == 0.5 and float(dense[0][100:].max()) == 0.5


def test_pose_that_sees_nothing_gives_nan_like_the_reference(dev, mod):
    """max_j m = 0 for a pose makes the reference's normalisation 0/0: rewards, mean and loss are NaN (src/model.py:226-231).
    The kernels reproduce that instead of inventing a value; dense and pruned paths agree."""
    from trajectory_optimization_b200 import _lib
    model, tools, ops = mod
    gen = np.random.default_rng(2)
    pts_np = _box(gen, 5000) + np.float32(500.0)
    poses, yaw = _s_curve(3, 2.0)
    quats = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], 1).astype(np.float32)
    K, Wd, Hd = tools.load_intrinsics(dev)
    ref = orc.traj_objective(pts_np, poses, quats, K_np, IMG_W, IMG_H, dtype=np.float32)
    assert np.isnan(ref["vis"]) and np.isnan(ref["rewards"]).all()
    try:
        for prune in (1, 0):
            ops._MODE.dense = not prune
            rewards, mean = ops.coverage_traj(torch.from_numpy(pts_np).to(dev), torch.from_numpy(poses).to(dev),
                                              torch.from_numpy(quats).to(dev), K, Wd, Hd)
            assert bool(torch.isnan(mean)) and bool(torch.isnan(rewards).all())
    finally:
        ops._MODE.dense = not 1


def _import_dropin(name):
    """Import `model` / `tools` by MODULE NAME from the drop-in directory, the way the reference's nodes do
    (sys.path.append(<pkg>/src); from model import ModelTraj — src/trajectory_optimization.py:6-9)."""
    import importlib
    import os
    import sys
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "trajectory_optimization_b200", "dropin")
    sys.path.insert(0, d)
    try:
        sys.modules.pop(name, None)
        return importlib.import_module(name)
    finally:
        sys.path.remove(d)
        sys.modules.pop(name, None)


def test_reference_optimisation_loops_run_unchanged_on_the_dropin(dev, sample_inputs):
    """The loops of src/pose_optimization.py:129-137 and src/trajectory_optimization.py:106-116, verbatim, on the drop-in
    `model`/`tools` modules: 10 Adam steps follow the loss history and reach the parameters of the reference run on CPU
    (fixtures opt_pose_sample / opt_traj_sample, made by running the reference)."""
    model = _import_dropin("model")
    tools = _import_dropin("tools")
    K, Wd, Hd = tools.load_intrinsics(device=dev)
    pts = torch.from_numpy(sample_inputs["pts"]).float().to(dev)
    g = load_golden("opt_pose_sample")
    m = model.ModelPose(points=pts, trans0=torch.from_numpy(g["in_trans0"]).to(dev), q0=torch.from_numpy(g["in_quat0"]).to(dev),
                        intrins=K, img_width=Wd, img_height=Hd, device=dev).to(dev)
    optimizer = torch.optim.Adam([{"params": list([m.trans]), "lr": float(g["in_lr_pose"])},
                                  {"params": list([m.quat]), "lr": float(g["in_lr_quat"])}])
    hist = []
    for _ in range(int(g["in_steps"])):
        optimizer.zero_grad()
        loss = m(debug=False)
        loss.backward()
        optimizer.step()
        hist.append(loss.item())
    assert rel_err(np.array(hist), g["out_loss_history"]) < 1e-4
    assert rel_err(m.trans.detach().cpu().numpy(), g["out_trans"]) < 1e-3
    assert rel_err(m.quat.detach().cpu().numpy(), g["out_quat"]) < 1e-3

    g = load_golden("opt_traj_sample")
    poses = torch.from_numpy(sample_inputs["poses"]).float()
    quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]]).repeat(poses.shape[0], 1)
    mt = model.ModelTraj(points=pts, wps_poses=poses, wps_quats=quats, intrins=K, img_width=Wd, img_height=Hd,
                         device=dev).to(dev)
    optimizer = torch.optim.Adam([{"params": list([mt.poses]), "lr": float(g["in_lr_pose"])},
                                  {"params": list([mt.quats]), "lr": float(g["in_lr_quat"])}])
    hist, vis = [], []
    for _ in range(int(g["in_steps"])):
        optimizer.zero_grad()
        loss = mt()
        loss.backward()
        optimizer.step()
        hist.append(loss.item())
        vis.append(float(mt.loss["vis"]))
    assert rel_err(np.array(hist), g["out_loss_history"]) < 2e-4
    assert rel_err(np.array(vis), g["out_vis_history"]) < 2e-4
    assert rel_err(mt.poses.detach().cpu().numpy(), g["out_poses"]) < 1e-3
    assert rel_err(mt.quats.detach().cpu().numpy(), g["out_quats"]) < 1e-3
    assert rel_err(float(torch.mean(mt.rewards)), float(g["out_mean_reward"])) < 1e-4


def test_fused_regularisers_match_the_torch_terms(dev, mod):
    """cov_traj_regularizers against the same three terms written with torch ops (the reference's formulas, vectorised)
    and their autograd gradients, on a wiggly path, a path with a repeated waypoint, and the unmoved path (l2 = 0)."""
    model, tools, ops = mod
    gen = np.random.default_rng(12)
    K, Wd, Hd = tools.load_intrinsics(dev)
    pts = torch.from_numpy(_box(gen, 2000))
    for case in ("wiggly", "repeated", "unmoved", "long"):
        W = 300 if case == "long" else 17
        poses0, _ = _s_curve(W, 14.0)
        poses = poses0 + (0 if case == "unmoved" else gen.normal(0, 0.15, poses0.shape).astype(np.float32))
        if case == "repeated":
            poses[5] = poses[4]
        quats = np.tile(np.array([1.0, 0, 0, 0], np.float32), (W, 1))
        m = model.ModelTraj(pts, torch.from_numpy(poses0), torch.from_numpy(quats), K, Wd, Hd, device=dev)
        with torch.no_grad():
            m.poses.copy_(torch.from_numpy(poses))
        m.loss["vis"] = torch.zeros((), device=dev)
        reg = ops.traj_regularizers(m.poses, m.poses0, m.smoothness_weight, m.traj_length_weight, m.eps)
        wts = torch.tensor([0.7, 1.3, 2.1], device=dev)
        (g_fused,) = torch.autograd.grad((reg * wts).sum(), [m.poses])
        m._criterion_terms_torch()
        ref = torch.stack([m.loss["l2"], m.loss["smooth"], m.loss["length"]])
        (g_ref,) = torch.autograd.grad((ref * wts).sum(), [m.poses])
        assert rel_err(reg.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 2e-5, case
        if case == "repeated":   # |ab| = 0 at the repeated waypoint: both follow d|v|/dv = 0 there
            assert torch.isfinite(g_fused).all() and torch.isfinite(g_ref).all()
        assert rel_err(g_fused.cpu().numpy(), g_ref.cpu().numpy()) < (2e-3 if case == "unmoved" else 2e-4), case
    # through the model: the opt-in flag changes the loss and its gradient only at the fp32-vs-fp64 level
    outs = []
    for flag in (False, True):
        m = model.ModelTraj(pts, torch.from_numpy(poses0), torch.from_numpy(quats), K, Wd, Hd, device=dev, fused_regularizers=flag)
        with torch.no_grad():
            m.poses.add_(0.1 * torch.sin(torch.arange(m.poses.numel(), device=dev).reshape(m.poses.shape).float()))
        loss = m(vis_wps_dist=0.0)
        loss.backward()
        outs.append((loss.item(), m.poses.grad.clone()))
    assert rel_err(outs[1][0], outs[0][0]) < 1e-5
    assert rel_err(outs[1][1].cpu().numpy(), outs[0][1].cpu().numpy()) < 2e-4


def test_model_cloud_cache_follows_in_place_edits_and_set_points(dev, mod):
    """The Morton-ordered copy / permutation / boxes are keyed on identity AND version of `model.points`: an in-place
    edit, an assignment and `set_points` all re-order; `refresh_points_` rewrites the cached buffers in place."""
    model, tools, ops = mod
    gen = np.random.default_rng(21)
    a, b = _box(gen, 150_000), _box(gen, 150_000, (-5, -8, -1), (20, 25, 3))
    poses, _ = _s_curve(9)
    quats = (gen.standard_normal((9, 4)) * 0.3 + np.array([1.0, 0, 0, 0])).astype(np.float32)
    K, Wd, Hd = tools.load_intrinsics(dev)

    def fresh(pts):
        m = model.ModelTraj(torch.from_numpy(pts), torch.from_numpy(poses), torch.from_numpy(quats), K, Wd, Hd, device=dev)
        m(vis_wps_dist=0.0)
        return m.rewards.clone(), float(m.loss["vis"])

    ra, va = fresh(a)
    rb, vb = fresh(b)
    assert not torch.equal(ra, rb)
    m = model.ModelTraj(torch.from_numpy(a), torch.from_numpy(poses), torch.from_numpy(quats), K, Wd, Hd, device=dev)
    m(vis_wps_dist=0.0)
    assert torch.equal(m.rewards, ra)
    ptr = m._pts32.data_ptr()
    m.points.copy_(torch.from_numpy(b))              # in-place edit: same object, new version
    m(vis_wps_dist=0.0)
    assert torch.equal(m.rewards, rb) and float(m.loss["vis"]) == vb
    m.set_points(torch.from_numpy(a))
    m(vis_wps_dist=0.0)
    assert torch.equal(m.rewards, ra)
    m.points = torch.from_numpy(b).to(dev)           # plain assignment, as a caller of the reference would do
    m(vis_wps_dist=0.0)
    assert torch.equal(m.rewards, rb)
    ptr = m._pts32.data_ptr()
    m.refresh_points_(torch.from_numpy(a).to(dev))   # graph-safe refresh: buffers keep their addresses
    assert m._pts32.data_ptr() == ptr
    m(vis_wps_dist=0.0)
    assert torch.equal(m.rewards, ra) and float(m.loss["vis"]) == va
    mp = model.ModelPose(torch.from_numpy(a).double(), torch.zeros(1, 3), torch.tensor([[1.0, 0, 0, 0]]), K, Wd, Hd, device=dev)
    l0 = float(mp())
    mp.points.copy_(torch.from_numpy(b).double())    # ModelPose keeps an fp32 working copy of an fp64 cloud
    assert float(mp()) != l0


def test_ops_follow_the_tensors_device_not_the_current_device(mod):
    """`ModelPose(..., device='cuda:1')` while cuda:0 is current: kernels, streams and workspaces must be on
    the tensors' device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    model, tools, ops = mod
    g = load_golden("traj_box")
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    torch.cuda.set_device(d0)
    out = []
    for d in (d0, d1):
        K, Wd, Hd = tools.load_intrinsics(d)
        m = model.ModelTraj(torch.from_numpy(g["in_points"]), torch.from_numpy(g["in_poses"]), torch.from_numpy(g["in_quats"]),
                            K, Wd, Hd, device=d)
        loss = m(vis_wps_dist=0.0)
        loss.backward()
        assert m.rewards.device == d and m.poses.grad.device == d
        out.append((float(loss), m.rewards.cpu(), m.poses.grad.cpu()))
    assert torch.cuda.current_device() == 0
    assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])


def test_two_gpu_sharded_objective_with_in_kernel_exchange():
    """scripts/dist_check.py on two GPUs (NCCL group): the point-sharded models reproduce the unsharded loss, rewards and
    gradients, once through the in-kernel NVLink exchange (cov_peer_allreduce) and once through NCCL all-reduces."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for peer, port in (("1", "29631"), ("0", "29632")):
        env = dict(os.environ, COV_PEER_EXCHANGE=peer)
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(root, "scripts", "dist_check.py")],
                             capture_output=True, text=True, env=env, timeout=600, cwd=root)
        assert out.returncode == 0 and "dist_check OK" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])
        assert ("unavailable" not in out.stderr) or peer == "0"


def test_peer_exchange_kernel_single_rank(dev, mod):
    """cov_peer_allreduce with a world of one (the exchange buffer is a plain zeroed device buffer): the in-place result
    equals the input for all three kinds, over several calls (the slot parity alternates with the device-side epoch) and
    when replayed from a CUDA graph.  The multi-rank behaviour is covered by scripts/dist_check.py on >= 2 GPUs."""
    import ctypes
    from trajectory_optimization_b200 import _lib
    model, tools, ops = mod
    L = _lib.lib()
    W = 321
    n_mm, n_acc = 2 * W, W * _lib.ACC_STRIDE + 1
    b_mm = (L.cov_peer_region_bytes(_lib.PEER_MINMAX_F32, n_mm, 1) + 255) // 256 * 256
    b_acc = (L.cov_peer_region_bytes(_lib.PEER_SUM_F64, n_acc, 1) + 255) // 256 * 256
    buf = torch.zeros(b_mm + b_acc, dtype=torch.uint8, device=dev)
    peers = _lib.Peers()
    peers.ptr[0] = buf.data_ptr()
    peers.world, peers.rank = 1, 0
    g = torch.Generator(device=dev).manual_seed(1)
    for it in range(5):
        mm = torch.rand(n_mm, device=dev, generator=g) - 0.5
        acc = torch.randn(n_acc, device=dev, generator=g, dtype=torch.float64)
        mm0, acc0 = mm.clone(), acc.clone()
        ops._call("cov_peer_allreduce", mm, _lib.PEER_MINMAX_F32, mm.data_ptr(), n_mm, ctypes.byref(peers), 0)
        ops._call("cov_peer_allreduce", acc, _lib.PEER_SUM_F64, acc.data_ptr(), n_acc, ctypes.byref(peers), b_mm)
        assert torch.equal(mm, mm0) and torch.equal(acc, acc0)
    # graph replays draw fresh epochs from the device-side ticket
    acc = torch.randn(n_acc, device=dev, generator=g, dtype=torch.float64)
    acc0 = acc.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        ops._call("cov_peer_allreduce", acc, _lib.PEER_SUM_F64, acc.data_ptr(), n_acc, ctypes.byref(peers), b_mm)
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(acc, acc0)
    with pytest.raises(RuntimeError):
        ops._call("cov_peer_allreduce", acc, 7, acc.data_ptr(), n_acc, ctypes.byref(peers), b_mm)


def test_pruned_path_takes_up_to_2048_poses_per_call(dev, mod):
    """The pruned kernels keep no pose table in shared memory: one call takes 32 x 64 = 2048 poses (the dense kernels
    1233).  A sweep of 1500 one-pose trajectories runs as ONE pruned pass A and one sweep launch and must agree with the
    dense sweep (which chunks the poses)."""
    from trajectory_optimization_b200 import _lib
    model, tools, ops = mod
    L = _lib.lib()
    assert L.cov_traj_max_poses_pruned() == 2048 and L.cov_traj_max_poses() < 2048
    gen = np.random.default_rng(23)
    pts = torch.from_numpy(_box(gen, 120_000)).to(dev)
    T = 1500
    poses = (gen.random((T, 1, 3)) * np.array([30, 30, 2]) + np.array([-5, -5, -0.5])).astype(np.float32)
    quats = gen.normal(0, 1, (T, 1, 4)).astype(np.float32)
    K, Wd, Hd = tools.load_intrinsics(dev)
    P, Q = torch.from_numpy(poses), torch.from_numpy(quats)
    means = ops.sweep_rewards(pts, P, Q, K, Wd, Hd)
    with ops.evaluation(dense=True):
        means_dense = ops.sweep_rewards(pts, P, Q, K, Wd, Hd)
    assert means.shape == (T,) and rel_err(means.cpu().numpy(), means_dense.cpu().numpy()) < 1e-9
    # and the trajectory objective itself with 1500 poses in one pruned call
    Pg = P.reshape(T, 3).to(dev).requires_grad_(True)
    Qg = Q.reshape(T, 4).to(dev).requires_grad_(True)
    rewards, mean = ops.coverage_traj(pts, Pg, Qg, K, Wd, Hd)
    mean.backward()
    assert torch.isfinite(rewards).all() and float(rewards.min()) >= 0.5 and torch.isfinite(Pg.grad).all()
    with pytest.raises(RuntimeError):          # the dense path cannot take them in one call, and says so
        ops.coverage_traj(pts, Pg, Qg, K, Wd, Hd, dense=True)
