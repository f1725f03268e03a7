"""Bit-exact parity of the binary frustum cull and the Katz HPR stage with the reference fixtures
and the CPU oracle (index sets must be identical on non-degenerate clouds)."""
import numpy as np
import pytest
import torch

from oracle import coverage_oracle as orc
from tests.conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
K_np, IMG_W, IMG_H = orc.load_intrinsics()


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def tools():
    from trajectory_optimization_b200 import tools, _lib
    _lib.lib()
    return tools


def test_cull_matches_reference_fixture(dev, tools):
    g = load_golden("cull")
    K, W, H = tools.load_intrinsics(dev)
    pts = torch.from_numpy(g["in_points"]).to(dev)
    culled, dm, fm = tools.get_cam_frustum_pts(pts.t(), H, W, K, float(g["in_min_d"]), float(g["in_max_d"]))
    assert np.array_equal(culled.detach().cpu().numpy(), g["out_culled"])
    assert np.array_equal(dm.detach().cpu().numpy(), g["out_dist_mask"]) and np.array_equal(fm.detach().cpu().numpy(), g["out_fov_mask"])
    from trajectory_optimization_b200 import model
    fb = model.get_fov_mask(pts, H, W, K, binary=True)
    assert np.array_equal(fb.detach().cpu().numpy(), g["out_fov_binary_model"])


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 2_000_003])
def test_cull_matches_oracle_ragged_sizes(n, dev, tools):
    gen = np.random.default_rng(n + 5)
    pts = (gen.random((n, 3), dtype=np.float32) * np.array([24, 24, 16], np.float32) - np.array([12, 12, 2], np.float32))
    K, W, H = tools.load_intrinsics(dev)
    culled, dm, fm = tools.get_cam_frustum_pts(torch.from_numpy(pts).to(dev).t(), H, W, K, 1.0, 10.0)
    ref_c, ref_dm, ref_fm = orc.frustum_cull(pts.T, IMG_H, IMG_W, K_np, 1.0, 10.0)
    assert np.array_equal(dm.detach().cpu().numpy(), ref_dm) and np.array_equal(fm.detach().cpu().numpy(), ref_fm)
    assert np.array_equal(culled.detach().cpu().numpy(), ref_c)


@pytest.mark.parametrize("name", ["hpr_shell", "hpr_shell_small", "hpr_halfspace"])
def test_flip_bit_exact_vs_reference(name, dev, tools):
    g = load_golden(name)
    f = tools.sphericalFlip(torch.from_numpy(g["in_points"]).to(dev), dev, int(g["in_R_param"]))
    assert np.array_equal(f.detach().cpu().numpy().view(np.uint32), g["out_flipped"].view(np.uint32))


def test_flip_bit_exact_vs_oracle_1m(dev, tools):
    gen = np.random.default_rng(1)
    d = gen.standard_normal((1_000_000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * gen.uniform(2, 8, (len(d), 1))).astype(np.float32)
    f = tools.sphericalFlip(torch.from_numpy(pts).to(dev), dev, 2)
    ref, _, _ = orc.spherical_flip(pts, 2)
    assert np.array_equal(f.detach().cpu().numpy().view(np.uint32), ref.view(np.uint32))


HPR_CASES = ["hpr_shell", "hpr_shell_small", "hpr_halfspace", "hpr_sample"]


@pytest.mark.parametrize("name", HPR_CASES)
def test_hidden_pts_removal_index_sets_bit_exact(name, dev, tools):
    """Visible-point index sets equal the reference's (Qhull) on the recorded fixtures, incl. the vertices[:-1] quirk."""
    g = load_golden(name)
    pts = torch.from_numpy(g["in_points"]).to(dev)
    vis, mask = tools.hidden_pts_removal(pts, dev, int(g["in_R_param"]))
    assert mask.dtype == torch.float32 and mask.shape == (len(pts),)
    idx = torch.nonzero(mask).reshape(-1).cpu().numpy()
    assert np.array_equal(idx, g["out_idx"])
    if "out_visible" in g:
        assert np.array_equal(vis.cpu().numpy(), g["out_visible"])
    hull = tools.convexHull(tools.sphericalFlip(pts, dev, 2), dev)
    assert hull.n_exact_fallback == 0  # every decision carried an fp64 certificate
    origin_is_vertex = name == "hpr_halfspace"
    assert (hull.vertices[-1] == len(pts)) == origin_is_vertex
    assert np.array_equal(hull.vertices[:-1], g["out_idx"])


def _hpr_shapes(gen, n):
    yield "gaussian blob off centre", gen.standard_normal((n, 3)) * np.array([1.0, 2.0, 0.5]) + np.array([6.0, 0.0, 1.0])
    yield "two clusters", np.concatenate([gen.standard_normal((n // 2, 3)) * 0.7 + np.array([4.0, 3.0, 0.0]),
                                          gen.standard_normal((n - n // 2, 3)) * 1.5 + np.array([-5.0, -2.0, 1.0])])
    yield "thick wall", gen.random((n, 3)) * np.array([0.3, 30, 10]) + np.array([5.0, -15, -5])
    yield "camera inside a box", (gen.random((n, 3)) - 0.5) * np.array([12, 9, 5])
    yield "ground plane with noise", np.stack([gen.uniform(-20, 20, n), gen.uniform(-20, 20, n),
                                               -1.5 + 0.05 * gen.standard_normal(n)], 1)


@pytest.mark.parametrize("n", [50, 1000, 20_000, 200_000])
def test_hidden_pts_removal_on_other_cloud_shapes(n, dev, tools):
    """Every stage of the hull pipeline (thread, warp, block per point, all-voxel sweep) sees work on these: clouds that do
    not surround the camera, clusters of different density, a near-planar cloud.  Index sets equal Qhull's."""
    gen = np.random.default_rng(n)
    for name, pts in _hpr_shapes(gen, n):
        pts = pts.astype(np.float32)
        ref_idx, _ = orc.hidden_pts_removal(pts, 2)
        vis, mask = tools.hidden_pts_removal(torch.from_numpy(pts).to(dev), dev, 2)
        assert np.array_equal(torch.nonzero(mask).reshape(-1).cpu().numpy(), ref_idx), name
        assert np.array_equal(vis.cpu().numpy(), pts[ref_idx]), name


@pytest.mark.parametrize("kind,n", [("shell", 1_000_000), ("halfspace", 300_000), ("tiny", 5), ("tiny", 64)])
def test_hidden_pts_removal_matches_oracle(kind, n, dev, tools):
    """BASELINE config 2 (1M-point shell cloud, camera inside) and a half-space cloud (origin is a hull vertex)."""
    gen = np.random.default_rng(1)
    if kind == "halfspace":
        pts = (gen.random((n, 3)) * np.array([20, 20, 4]) + np.array([-10, -10, 2])).astype(np.float32)
    else:
        d = gen.standard_normal((n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (d * gen.uniform(2, 8, (n, 1))).astype(np.float32)
    ref_idx, ref_mask = orc.hidden_pts_removal(pts, 2)
    vis, mask = tools.hidden_pts_removal(torch.from_numpy(pts).to(dev), dev, 2)
    assert np.array_equal(torch.nonzero(mask).reshape(-1).cpu().numpy(), ref_idx)
    assert np.array_equal(vis.cpu().numpy(), pts[ref_idx])


def test_model_pose_hpr_branch(dev, tools):
    """ModelPose.forward(hpr=True) multiplies the observations by the occlusion mask of the untransformed cloud
    (reference src/model.py:112-115)."""
    from trajectory_optimization_b200 import model
    g = load_golden("hpr_shell")
    pts = torch.from_numpy(g["in_points"]).to(dev)
    K, W, H = tools.load_intrinsics(dev)
    m = model.ModelPose(pts, torch.tensor([[0.1, 0.2, 0.0]]), torch.tensor([[1.0, 0.0, 0.0, 0.0]]), K, W, H, device=dev)
    loss_hpr = m(hpr=True)
    obs_hpr = m.observations.detach().clone()
    loss_hpr.backward()
    m2 = model.ModelPose(pts, torch.tensor([[0.1, 0.2, 0.0]]), torch.tensor([[1.0, 0.0, 0.0, 0.0]]), K, W, H, device=dev)
    m2()
    mask = torch.zeros(len(pts), device=dev)
    mask[torch.from_numpy(g["out_idx"]).to(dev)] = 1
    assert torch.equal(obs_hpr, (m2.observations.detach() * mask))
    ref = orc.pose_objective(g["in_points"], [0.1, 0.2, 0.0], [1.0, 0, 0, 0], K_np, IMG_W, IMG_H, weight=mask.cpu().numpy())
    assert abs(loss_hpr.item() - float(ref["loss"])) / float(ref["loss"]) < 1e-4
    assert np.abs(m.trans.grad.cpu().numpy().ravel() - ref["g_trans"]).max() / np.abs(ref["g_trans"]).max() < 1e-4


# ---------------- PointCloud2 codec (reference src/pointcloud_utils.py) ----------------
PC2_CASES = ["pc2_xyz12", "pc2_xyzi_padded32", "pc2_unaligned_f64", "pc2_dense", "pc2_empty"]


class _Field:
    def __init__(self, name, offset, datatype):
        self.name, self.offset, self.datatype, self.count = name, offset, datatype, 1


class _Msg:
    is_bigendian = False
    height = 1


@pytest.mark.parametrize("name", PC2_CASES)
def test_pointcloud2_decode_matches_reference(name, dev):
    """cov_pc2_to_xyz against the reference's pointcloud2_to_xyz_array output (fixture), bit for bit in fp32."""
    from trajectory_optimization_b200 import pointcloud_utils as pcu
    g = load_golden(name)
    msg = _Msg()
    msg.width, msg.point_step = int(g["in_n"]), int(g["in_point_step"])
    msg.fields = [_Field(str(nm), int(o), int(t)) for nm, o, t in zip(g["in_field_names"], g["in_field_offsets"], g["in_field_types"])]
    msg.data = g["in_data"].tobytes()
    out = pcu.pointcloud2_to_xyz_tensor(msg, device=dev)
    ref = g["out_xyz"].reshape(-1, 3)
    assert out.dtype == torch.float32 and out.shape == ref.shape
    assert np.array_equal(out.cpu().numpy(), ref.astype(np.float32))
    out_all = pcu.pointcloud2_to_xyz_tensor(msg, remove_nans=False, device=dev)
    assert np.array_equal(out_all.cpu().numpy(), g["out_xyz_all"].reshape(-1, 3).astype(np.float32), equal_nan=True)
    arr = pcu.pointcloud2_to_xyz_array(msg, device=dev)     # the reference's return type
    assert arr.dtype == np.float64 and arr.shape == ref.shape


@pytest.mark.parametrize("n", [1, 255, 257, 1_000_003])
def test_pointcloud2_codec_round_trip_and_oracle(n, dev):
    from trajectory_optimization_b200 import pointcloud_utils as pcu
    gen = np.random.default_rng(n)
    pts = gen.normal(0, 20, (n, 3)).astype(np.float32)
    inten = gen.random(n).astype(np.float32)
    bad = gen.random(n) < 0.03
    pts[bad, gen.integers(0, 3, int(bad.sum()))] = np.nan
    pts[gen.random(n) < 0.01, 1] = np.inf
    for extra in (None, inten):
        payload, dense = pcu.xyz_to_payload(torch.from_numpy(pts).to(dev), None if extra is None else torch.from_numpy(extra).to(dev))
        ref_bytes, step, ref_dense = orc.xyz_to_pc2(pts if extra is None else np.concatenate([pts, extra[:, None]], 1))
        assert payload.cpu().numpy().tobytes() == ref_bytes and dense == bool(ref_dense)
        back = pcu.payload_to_xyz(payload, n, step, 0, 4, 8)
        fields = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7)] + ([] if extra is None else [("i", 12, 7)])
        ref = orc.pc2_to_xyz(ref_bytes, n, step, fields)
        assert np.array_equal(back.cpu().numpy(), ref.astype(np.float32))
        assert back.shape[0] == int(np.isfinite(pts).all(1).sum())
    clean = torch.from_numpy(np.nan_to_num(pts, nan=0.0, posinf=1.0)).to(dev)
    assert pcu.xyz_to_payload(clean)[1] is True


def test_pointcloud2_errors_are_loud(dev):
    from trajectory_optimization_b200 import pointcloud_utils as pcu
    with pytest.raises(RuntimeError):
        pcu.payload_to_xyz(torch.zeros(48, dtype=torch.uint8), 4, 12, 0, 4, 8)          # CPU tensor
    with pytest.raises(RuntimeError, match="bad argument"):
        pcu.payload_to_xyz(torch.zeros(48, dtype=torch.uint8, device=dev), 4, 12, 0, 4, 10)   # z field runs past the record
    msg = _Msg()
    msg.width, msg.point_step, msg.data = 1, 12, bytes(12)
    msg.fields = [_Field("x", 0, 7), _Field("y", 4, 7)]
    with pytest.raises(ValueError):
        pcu.pointcloud2_to_xyz_tensor(msg, device=dev)


# ---------------- voxel-grid filter (launch/voxels_filtering.launch, pcl::VoxelGrid restated) ----------------
@pytest.mark.parametrize("n,leaf,axis", [(1, 0.1, 2), (1000, 0.1, 2), (20_000, 0.25, 2), (300_007, 0.1, None), (50_000, 50.0, 0)])
def test_voxel_grid_matches_oracle(n, leaf, axis, dev, tools):
    gen = np.random.default_rng(n)
    pts = (gen.random((n, 3)) * np.array([12, 9, 8]) + np.array([-6, -4, -4])).astype(np.float32)
    if n > 100:
        pts[::53, 0] = np.nan
        pts[7] = [np.inf, 0, 0]
        pts[11:13] = pts[10]                       # duplicates share a voxel
    name = None if axis is None else "xyz"[axis]
    out = tools.voxel_grid_filter(torch.from_numpy(pts).to(dev), leaf, name, -2.5, 2.5)
    ref = orc.voxel_grid(pts, leaf, axis, -2.5, 2.5)
    assert out.dtype == torch.float32 and tuple(out.shape) == ref.shape
    assert np.array_equal(out.cpu().numpy(), ref)           # same voxels, same order, same fp32 centroids
    out2 = tools.voxel_grid_filter(torch.from_numpy(pts).to(dev), leaf, name, -2.5, 2.5)
    assert torch.equal(out, out2)                            # deterministic


def test_voxel_grid_rejects_too_fine_a_grid(dev, tools):
    pts = torch.tensor([[0.0, 0.0, 0.0], [100.0, 100.0, 1.0]], device=dev)
    with pytest.raises(RuntimeError, match="too small"):
        tools.voxel_grid_filter(pts, 1e-3, None)
    assert tools.voxel_grid_filter(torch.zeros(0, 3, device=dev)).shape == (0, 3)


def test_standalone_helpers_match_reference_fixture(dev):
    """to_camera_frame / get_dist_mask / get_fov_mask (src/model.py:13-57) on CUDA tensors vs values recorded by running
    the reference (tests/golden/helpers.npz); fp32 tolerance 1e-5 relative."""
    from trajectory_optimization_b200 import model, tools as T
    g = load_golden("helpers")
    K, W, H = T.load_intrinsics(dev)
    pts = torch.from_numpy(g["in_points"]).to(dev)
    cam = model.to_camera_frame(pts, torch.from_numpy(g["in_quat"]).to(dev), torch.from_numpy(g["in_trans"]).to(dev))
    assert rel_err(cam.cpu().numpy(), g["out_cam"]) < 1e-5
    cam_ref = torch.from_numpy(g["out_cam"]).to(dev)
    assert rel_err(model.get_dist_mask(cam_ref, 1.0, 5.0).cpu().numpy(), g["out_dist"]) < 1e-5
    assert rel_err(model.get_fov_mask(cam_ref, H, W, K).cpu().numpy(), g["out_fov"]) < 1e-5
    # and they stay differentiable like the reference's
    q = torch.from_numpy(g["in_quat"]).to(dev).requires_grad_(True)
    t = torch.from_numpy(g["in_trans"]).to(dev).requires_grad_(True)
    c = model.to_camera_frame(pts, q, t)
    (model.get_dist_mask(c) * model.get_fov_mask(c, H, W, K)).sum().backward()
    assert torch.isfinite(q.grad).all() and torch.isfinite(t.grad).all() and float(t.grad.abs().max()) > 0


def test_pointcloud2_builders_take_numpy_like_the_reference(dev):
    """xyz/xyzi_array_to_pointcloud2 with the NUMPY arrays the unmodified nodes pass (src/pose_optimization.py:108-112,
    src/trajectory_optimization.py:147-157 -> src/tools.py:224-231): payload bytes and flags equal the reference's
    host encoder (np.asarray(points, np.float32).tobytes(), np.isfinite(points).all())."""
    import os
    import sys
    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims")
    sys.path.insert(0, shims)
    try:
        from trajectory_optimization_b200 import pointcloud_utils as pcu
        from trajectory_optimization_b200 import tools as T
        import rospy
        gen = np.random.default_rng(2)
        xyz64 = gen.normal(size=(1001, 3)) * 7
        xyzi = np.concatenate([xyz64.astype(np.float32), gen.random((1001, 1))], axis=1)   # float64, as np.concatenate gives
        for arr in (xyz64, xyz64.astype(np.float32), torch.from_numpy(xyz64), torch.from_numpy(xyz64).to(dev)):
            m = pcu.xyz_array_to_pointcloud2(arr, stamp=1.0, frame_id="map")
            assert m.data == np.asarray(xyz64, np.float32).tobytes() and m.is_dense == 1
            assert (m.width, m.height, m.point_step, m.row_step) == (1001, 1, 12, 1001)
        m = pcu.xyzi_array_to_pointcloud2(xyzi, frame_id="map")
        assert m.data == np.asarray(xyzi, np.float32).tobytes() and m.is_dense == 1 and m.point_step == 16
        xyzi[5, 3] = np.nan
        assert pcu.xyzi_array_to_pointcloud2(xyzi).is_dense == 0
        big = xyz64.copy()
        big[0, 0] = 1e300   # finite in fp64: the reference's flag looks at the input, the payload holds inf
        assert pcu.xyz_array_to_pointcloud2(big).is_dense == 1
        del rospy.PUBLISHED[:]
        T.publish_pointcloud(xyzi, "/pts/rewards", rospy.Time.now(), "map")   # the call the nodes make
        assert rospy.PUBLISHED[0][0] == "/pts/rewards" and len(rospy.PUBLISHED[0][1].data) == 16 * 1001
        # decode of what we encode, through the pinned staging path
        back = pcu.pointcloud2_to_xyz_array(pcu.xyz_array_to_pointcloud2(xyz64))
        assert back.dtype == np.float64 and np.array_equal(back, xyz64.astype(np.float32).astype(np.float64))
    finally:
        sys.path.remove(shims)


def test_multi_camera_visibility_matches_per_camera_oracle(dev, tools):
    """transform -> cull -> HPR for 5 cameras in one call (reference src/pc_processor.py:158-182 per camera): index sets
    equal the oracle's cull + Qhull HPR run on the SAME camera-frame points (bit-exact sets on a non-degenerate cloud)."""
    gen = np.random.default_rng(17)
    d = gen.normal(size=(60_000, 3))
    pts = (d / np.linalg.norm(d, axis=1, keepdims=True) * gen.uniform(1.5, 9.0, (60_000, 1))).astype(np.float32)
    yaws = np.deg2rad([0, 72, -72, 144, -144])
    # camera optical frames looking outwards around the vertical axis
    from trajectory_optimization_b200 import multicam
    R0 = multicam.R_BODY_OPTICAL.double()
    quats, trans = [], []
    for i, y in enumerate(yaws):
        Rz = torch.tensor([[np.cos(y), -np.sin(y), 0.0], [np.sin(y), np.cos(y), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
        quats.append(multicam.matrix_to_quat_wxyz((Rz @ R0)[None])[0])
        trans.append(torch.tensor([0.1 * i, -0.05 * i, 0.02 * i], dtype=torch.float64))
    K, W, H = tools.load_intrinsics(dev)
    res = tools.multi_camera_visibility(torch.from_numpy(pts), torch.stack(trans), torch.stack(quats), K, H, W, 1.0, 10.0,
                                        device=dev)
    assert [r["camera"] for r in res] == [0, 1, 2, 3, 4]
    from trajectory_optimization_b200.model import to_camera_frame
    for r in res:
        c = r["camera"]
        cam = to_camera_frame(torch.from_numpy(pts).to(dev), quats[c][None].float().to(dev), trans[c][None].float().to(dev))
        cam_np = cam.cpu().numpy()
        culled, dm, fm = orc.frustum_cull(cam_np.T, IMG_H, IMG_W, K_np, 1.0, 10.0)
        ref_idx = np.flatnonzero(dm & fm)
        assert np.array_equal(r["frustum_idx"].cpu().numpy(), ref_idx) and len(ref_idx) > 1000
        assert np.array_equal(r["frustum_points"].cpu().numpy(), culled)
        vis, _ = orc.hidden_pts_removal(culled, 2)
        assert np.array_equal(r["visible_idx"].cpu().numpy(), ref_idx[vis])
        assert np.array_equal(r["visible_points"].cpu().numpy(), culled[vis])
