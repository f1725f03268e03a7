"""Pin the CPU oracle (oracle/coverage_oracle.py) to fixtures produced by RUNNING THE
REFERENCE (tests/golden/make_golden.py).  CPU only; a few seconds."""
import numpy as np
import pytest

from oracle import coverage_oracle as orc
from tests.conftest import load_golden, rel_err

K, IMG_W, IMG_H = orc.load_intrinsics()
# fp32 arithmetic in a different (but equivalent) op order: a few 1e-7 per op, <=2e-6 end to end.
TOL32 = 3e-6
# the closed-form gradients reproduce fp64 autograd to rounding
TOL64 = 1e-12

POSE_CASES = ["pose_sample", "pose_synth0", "pose_synth1", "pose_synth2", "pose_synth_clip"]
TRAJ_CASES = ["traj_sample", "traj_sample_all", "traj_box", "traj_compact", "traj_tiny"]


def _points(g, sample_inputs):
    return g["in_points"] if "in_points" in g else sample_inputs["pts"]


@pytest.mark.parametrize("name", POSE_CASES)
def test_pose_objective_matches_reference(name, sample_inputs):
    g = load_golden(name)
    pts = _points(g, sample_inputs)
    wgt = g["in_weight"] if "in_weight" in g else None
    args = (pts, g["in_trans"], g["in_quat"], K, IMG_W, IMG_H, float(g["in_min_d"]), float(g["in_max_d"]))
    r32 = orc.pose_objective(*args, weight=wgt, dtype=np.float32)
    assert rel_err(r32["loss"], g["out_loss"]) < TOL32
    assert rel_err(r32["obs"], g["out_obs"]) < TOL32
    assert rel_err(r32["g_trans"], g["out_g_trans"]) < TOL32
    assert rel_err(r32["g_quat"], g["out_g_quat"]) < TOL32
    r64 = orc.pose_objective(*args, weight=wgt, dtype=np.float64)
    assert rel_err(r64["loss"], g["out_loss64"]) < TOL64
    assert rel_err(r64["obs"], g["out_obs64"]) < TOL64
    assert rel_err(r64["g_trans"], g["out_g_trans64"]) < TOL64
    assert rel_err(r64["g_quat"], g["out_g_quat64"]) < TOL64


@pytest.mark.parametrize("name", TRAJ_CASES)
def test_traj_objective_matches_reference(name, sample_inputs):
    g = load_golden(name)
    pts = _points(g, sample_inputs)
    poses, quats = g["in_poses"], g["in_quats"]
    step = orc.wps_step_from_path(poses, float(g["in_vis_wps_dist"]))
    assert step == int(g["out_wps_step"])
    sel = np.arange(0, len(poses), step)
    for dt, sfx, tol in ((np.float32, "", TOL32), (np.float64, "64", TOL64)):
        r = orc.traj_objective(pts, poses[sel], quats[sel], K, IMG_W, IMG_H,
                               float(g["in_min_d"]), float(g["in_max_d"]), dtype=dt)
        gp = np.zeros(poses.shape)
        gq = np.zeros(quats.shape)
        gp[sel], gq[sel] = r["g_poses"], r["g_quats"]  # skipped waypoints get no visibility gradient
        assert rel_err(r["vis"], g["out_vis" + sfx]) < tol
        assert rel_err(r["rewards"], g["out_rewards" + sfx]) < tol
        assert rel_err(gp, g["out_gv_poses" + sfx]) < tol
        assert rel_err(gq, g["out_gv_quats" + sfx]) < tol
    t = orc.traj_criterion_terms(poses, poses, float(g["in_sw"]), float(g["in_lw"]), dtype=np.float32)
    assert rel_err(t["smooth"], g["out_smooth"]) < TOL32
    assert t["l2"] == g["out_l2"] == 0.0 and t["length"] == g["out_length"] == 0.0
    assert rel_err(g["out_vis"] + g["out_l2"] + g["out_smooth"] + g["out_length"], g["out_loss"]) < 1e-6


def test_traj_compact_exercises_tie_paths():
    """The compact cloud has min_j m > 0, so the arg-min and arg-max points carry gradient."""
    g = load_golden("traj_compact")
    part = orc.traj_partials(g["in_points"], g["in_poses"], g["in_quats"], K, IMG_W, IMG_H, dtype=np.float64)
    assert (part["mins"][:4] > 0).all() and part["mins"][4] == 0  # last pose: underflow ties at 0
    assert (part["amax_n"] == 1).all() and (part["amin_n"][:4] == 1).all() and part["amin_n"][4] > 1
    full = orc.traj_grads_from_partials(part, g["in_poses"], g["in_quats"], part["n"])
    part0 = dict(part, amax_n=np.zeros_like(part["amax_n"]), amin_n=np.zeros_like(part["amin_n"]))
    nomm = orc.traj_grads_from_partials(part0, g["in_poses"], g["in_quats"], part["n"])
    assert rel_err(nomm["g_poses"], full["g_poses"]) > 1e-2  # the min/max path is not negligible


def test_traj_sharded_partials_are_additive():
    """Accumulators of two point shards (with global min/max) add up to the unsharded ones."""
    g = load_golden("traj_box")
    pts, poses, quats = g["in_points"], g["in_poses"][::2], g["in_quats"][::2]
    mm = orc.traj_minmax(pts, poses, quats, K, IMG_W, IMG_H, dtype=np.float32)
    whole = orc.traj_partials(pts, poses, quats, K, IMG_W, IMG_H, dtype=np.float32, minmax=mm)
    cut = 7777
    a = orc.traj_partials(pts[:cut], poses, quats, K, IMG_W, IMG_H, dtype=np.float32, minmax=mm)
    b = orc.traj_partials(pts[cut:], poses, quats, K, IMG_W, IMG_H, dtype=np.float32, minmax=mm)
    summed = {k: a[k] + b[k] for k in a if k not in ("rewards", "mins", "maxs")}
    summed.update(mins=mm[0], maxs=mm[1])
    ga = orc.traj_grads_from_partials(whole, poses, quats, len(pts))
    gb = orc.traj_grads_from_partials(summed, poses, quats, len(pts))
    assert rel_err(gb["g_poses"], ga["g_poses"]) < 1e-9 and rel_err(gb["g_quats"], ga["g_quats"]) < 1e-9
    assert np.array_equal(np.concatenate([a["rewards"], b["rewards"]]), whole["rewards"])


@pytest.mark.parametrize("name", ["hpr_shell", "hpr_shell_small", "hpr_halfspace", "hpr_sample"])
def test_hpr_index_sets_bit_exact(name):
    g = load_golden(name)
    f, radius, n = orc.spherical_flip(g["in_points"], int(g["in_R_param"]))
    if "out_flipped" in g:
        assert np.array_equal(f.view(np.uint32), g["out_flipped"].view(np.uint32))  # bit-exact fp32 flip
    idx, mask = orc.hidden_pts_removal(g["in_points"], int(g["in_R_param"]))
    assert np.array_equal(idx, g["out_idx"])
    assert mask.dtype == np.float32 and mask.sum() == len(idx)


def test_frustum_cull_bit_exact():
    g = load_golden("cull")
    culled, dm, fm = orc.frustum_cull(g["in_points"].T, IMG_H, IMG_W, K, float(g["in_min_d"]), float(g["in_max_d"]))
    assert np.array_equal(culled, g["out_culled"])
    assert np.array_equal(dm, g["out_dist_mask"]) and np.array_equal(fm, g["out_fov_mask"])
    assert np.array_equal(orc.fov_mask(g["in_points"], IMG_H, IMG_W, K, binary=True), g["out_fov_binary_model"])


def test_helpers_match_reference():
    g = load_golden("helpers")
    cam = orc.to_camera_frame(g["in_points"], g["in_quat"], g["in_trans"], np.float32)
    assert np.abs(cam - g["out_cam"]).max() < 5e-6
    assert np.abs(orc.dist_mask(g["out_cam"]) - g["out_dist"]).max() < 1e-6
    assert np.abs(orc.fov_mask(g["out_cam"], IMG_H, IMG_W, K) - g["out_fov"]).max() < 1e-6


def test_sample_known_answers():
    """Digits recorded in SURVEY.md §8c from the reference on its shipped sample."""
    g = load_golden("pose_sample")
    assert abs(float(g["out_loss"]) - 5.38444088306278e-4) < 1e-12
    t = load_golden("traj_sample")
    assert int(t["out_wps_step"]) == 2
    assert abs(float(t["out_loss"]) - 6.954090595) < 1e-6
    assert abs(float(t["out_rewards"].mean()) - 0.5291025043) < 1e-7


@pytest.mark.parametrize("name", ["traj_box", "traj_compact"])
def test_torch_port_matches_reference(name):
    """oracle/torch_port.py (the CPU baseline bench.py times) reproduces the reference's numbers."""
    import torch
    from oracle import torch_port
    g = load_golden(name)
    step = int(g["out_wps_step"])
    K = torch.from_numpy(orc.K_DEFAULT.copy())
    poses, quats = torch.from_numpy(g["in_poses"][::step].copy()), torch.from_numpy(g["in_quats"][::step].copy())
    vis, gp, gq = torch_port.traj_step(torch.from_numpy(g["in_points"]), poses, quats, K, IMG_W, IMG_H)
    assert rel_err(vis.item(), g["out_vis"]) < 1e-6
    assert rel_err(gp.numpy(), g["out_gv_poses"][::step]) < 1e-5
    assert rel_err(gq.numpy(), g["out_gv_quats"][::step]) < 1e-5


def test_torch_port_pose_matches_reference(sample_inputs):
    import torch
    from oracle import torch_port
    g = load_golden("pose_sample")
    t = torch.from_numpy(g["in_trans"]).requires_grad_(True)
    q = torch.from_numpy(g["in_quat"]).requires_grad_(True)
    loss, obs = torch_port.pose_loss(torch.from_numpy(sample_inputs["pts"]), t, q, torch.from_numpy(orc.K_DEFAULT.copy()),
                                     IMG_W, IMG_H)
    loss.backward()
    assert rel_err(loss.item(), g["out_loss"]) < 1e-6 and rel_err(obs.detach().numpy(), g["out_obs"]) < 1e-6
    assert rel_err(t.grad.numpy(), g["out_g_trans"]) < 1e-5 and rel_err(q.grad.numpy(), g["out_g_quat"]) < 1e-5


PC2_CASES = ["pc2_xyz12", "pc2_xyzi_padded32", "pc2_unaligned_f64", "pc2_dense", "pc2_empty"]


@pytest.mark.parametrize("name", PC2_CASES)
def test_pointcloud2_decode_restatement_matches_reference(name):
    """oracle.pc2_to_xyz against the reference's pointcloud2_to_xyz_array (fixtures made by running it)."""
    g = load_golden(name)
    fields = list(zip([str(s) for s in g["in_field_names"]], [int(v) for v in g["in_field_offsets"]],
                      [int(v) for v in g["in_field_types"]]))
    n, step = int(g["in_n"]), int(g["in_point_step"])
    out = orc.pc2_to_xyz(g["in_data"].tobytes(), n, step, fields, remove_nans=True)
    out_all = orc.pc2_to_xyz(g["in_data"].tobytes(), n, step, fields, remove_nans=False)
    assert out.dtype == np.float64 and np.array_equal(out, g["out_xyz"].reshape(-1, 3))
    assert np.array_equal(out_all, g["out_xyz_all"].reshape(-1, 3), equal_nan=True)
    assert out.shape[0] <= n and (name != "pc2_dense" or out.shape[0] == n)
    payload, step_out, dense = orc.xyz_to_pc2(out.astype(np.float32))
    back = orc.pc2_to_xyz(payload, out.shape[0], step_out, [("x", 0, 7), ("y", 4, 7), ("z", 8, 7)], remove_nans=False)
    assert step_out == 12 and dense == 1 and np.array_equal(back.astype(np.float32), out.astype(np.float32))


def test_voxel_grid_restatement_properties():
    """pcl::VoxelGrid restatement (parity unpinned: no PCL here): one point per occupied voxel, every centroid inside its
    voxel, voxels in ascending index, pass-through limits inclusive, non-finite points dropped."""
    gen = np.random.default_rng(0)
    pts = (gen.random((20000, 3)) * np.array([6, 5, 8]) + np.array([-3, -2, -4])).astype(np.float32)
    pts[::97, 1] = np.nan
    pts[5] = [0.0, 0.0, 2.5]      # on the upper limit: kept
    pts[6] = [0.0, 0.0, 2.5001]   # just outside: dropped
    leaf = 0.25
    out = orc.voxel_grid(pts, leaf, 2, -2.5, 2.5)
    keep = np.isfinite(pts).all(1) & (pts[:, 2] >= -2.5) & (pts[:, 2] <= 2.5)
    q = pts[keep]
    inv = np.float32(1) / np.float32(leaf)
    cell = np.floor(q * inv).astype(np.int64)
    assert out.shape[0] == len(np.unique(cell, axis=0)) and out.dtype == np.float32
    ocell = np.floor(out * inv + 0.0).astype(np.int64)
    lo = ocell.astype(np.float32) * np.float32(leaf)
    assert (out >= lo - 1e-5).all() and (out <= lo + np.float32(leaf) + 1e-5).all()
    mn = cell.min(0)
    d = cell.max(0) - mn + 1
    lin = (ocell[:, 0] - mn[0]) + (ocell[:, 1] - mn[1]) * d[0] + (ocell[:, 2] - mn[2]) * d[0] * d[1]
    assert (np.diff(lin) > 0).all()
    assert out[:, 2].max() <= 2.5 and np.isfinite(out).all()
    assert orc.voxel_grid(pts, 100.0, None).shape == (8, 3)          # the cloud straddles the origin: 2 x 2 x 2 voxels
    one = orc.voxel_grid(np.nan_to_num(pts) + np.float32(10.0), 100.0, None)
    assert one.shape == (1, 3) and np.allclose(one[0], (np.nan_to_num(pts) + np.float32(10.0)).mean(0), rtol=1e-4)
    with pytest.raises(OverflowError):
        orc.voxel_grid(pts, 1e-4, None)
