"""The reference's node scripts run UNCHANGED on the drop-in.

The real `/root/reference/src/pose_optimization.py` and `trajectory_optimization.py` are loaded from where they lie
(never copied), with `oracle/shims` providing ROS and `trajectory_optimization_b200/dropin` ahead of the reference's
own `src/` on sys.path, so their `from model import ...`, `from tools import ...`, `from pointcloud_utils import ...`
resolve to this repository.  A test then builds `PoseOpt` / `TrajOpt` and drives `callback(pc_msg, pose_msg)` with fake
messages through the optimisation loop and every publisher.  The same callbacks are then run a second time with the
reference's OWN modules, and what the two runs publish (optimised pose / path) must agree.

There is no GPU in the build container, so the five coverage C-ABI calls and the two codec kernels are served by the
oracle-backed stand-ins of tests/_standins.py (device = cpu); everything around them — module names, signatures, numpy /
tensor conventions of the publishers, Adam on `.trans/.quat/.poses/.quats`, `.observations/.rewards/.loss` — is the
product's.  `/root/reference` does not exist on the GPU box: the module is skipped there."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("COV_REFERENCE_ROOT", "/root/reference")
SHIMS = os.path.join(ROOT, "oracle", "shims")
DROPIN = os.path.join(ROOT, "trajectory_optimization_b200", "dropin")
NODE_MODULES = ("model", "tools", "pointcloud_utils")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present")


class _Env:
    """sys.path / sys.modules arranged for one run of a node script, restored afterwards."""

    def __init__(self, dropin):
        self.dropin = dropin

    def __enter__(self):
        self.path, self.mods = list(sys.path), {k: sys.modules.get(k) for k in NODE_MODULES}
        for k in NODE_MODULES:
            sys.modules.pop(k, None)
        sys.path[:0] = [SHIMS] + ([DROPIN] if self.dropin else [os.path.join(REF, "src")])
        np.float = float  # src/pointcloud_utils.py:180 (removed from numpy >= 1.24)
        if not hasattr(np, "_fromstring_text"):  # binary np.fromstring (src/pointcloud_utils.py:71) left numpy 2.3
            np._fromstring_text = np.fromstring
            np.fromstring = lambda data, dtype=float, **kw: np.frombuffer(data, dtype=dtype).copy()
        import rospy
        rospy.PARAMS.clear()
        del rospy.PUBLISHED[:]
        return rospy

    def __exit__(self, *exc):
        sys.path[:] = self.path
        for k, v in self.mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def _load_node(name):
    spec = importlib.util.spec_from_file_location("refnode_" + name, os.path.join(REF, "src", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _standins():
    from tests._standins import NumpyCodec, OracleBackend
    from trajectory_optimization_b200 import ops, pointcloud_utils
    saved = ops._BACKEND, pointcloud_utils._CODEC
    ops._BACKEND, pointcloud_utils._CODEC = OracleBackend(), NumpyCodec()
    return saved


def _restore(saved):
    from trajectory_optimization_b200 import ops, pointcloud_utils
    ops._BACKEND, pointcloud_utils._CODEC = saved


def _cloud_msg(pts, frame="map"):
    from sensor_msgs.msg import PointCloud2, PointField
    m = PointCloud2()
    m.header.frame_id = frame
    m.height, m.width = 1, len(pts)
    m.fields = [PointField(n, 4 * i, PointField.FLOAT32, 1) for i, n in enumerate("xyz")]
    m.point_step, m.row_step = 12, 12 * len(pts)
    m.data = np.asarray(pts, np.float32).tobytes()
    return m


def _pose_msg(t, q_wxyz, frame="map"):
    from geometry_msgs.msg import PoseStamped
    m = PoseStamped()
    m.header.frame_id = frame
    m.pose.position.x, m.pose.position.y, m.pose.position.z = map(float, t)
    m.pose.orientation.w, m.pose.orientation.x, m.pose.orientation.y, m.pose.orientation.z = map(float, q_wxyz)
    return m


def _path_msg(poses, frame="map"):
    from nav_msgs.msg import Path
    p = Path()
    p.header.frame_id = frame
    p.poses = [_pose_msg(t, (1.0, 0.0, 0.0, 0.0), frame) for t in poses]
    return p


def _sample(n_pts=4000):
    from tests.conftest import load_golden
    s = load_golden("sample_inputs")
    pts = s["pts"][:: max(1, len(s["pts"]) // n_pts)].astype(np.float32).copy()
    pts[7] = np.nan  # one invalid return: pointcloud2_to_xyz_array must drop it
    return pts, s["poses"].astype(np.float32)


def _f(v):
    return np.array([float(x) for x in v], np.float64)


def _run_pose_node(dropin, pts):
    with _Env(dropin) as rospy:
        saved = _standins() if dropin else None
        try:
            node = _load_node("pose_optimization")
            if not dropin:  # numpy >= 2.3 has no ndarray.tostring (src/pointcloud_utils.py:311): the REFERENCE arm skips
                node.publish_pointcloud = lambda *a, **k: rospy.PUBLISHED.append(("skipped", None))  # the debug cloud
            rospy.PARAMS.update({"pose_opt/opt_steps": 40, "pose_opt/lr_pose": 0.02, "pose_opt/lr_quat": 0.02})
            opt = node.PoseOpt(pc_topic="/pts", input_pose_topic="/pose", device=torch.device("cpu"))
            opt.callback(_cloud_msg(pts), _pose_msg((6.0, 2.0, 0.0), (1.0, 0.0, 0.0, 0.0)))
            published = list(rospy.PUBLISHED)
            return opt, published, node
        finally:
            if saved:
                _restore(saved)


def test_pose_optimization_node_runs_unchanged_and_matches_the_reference_node():
    pts, _ = _sample()
    opt, pub, node = _run_pose_node(True, pts)
    import trajectory_optimization_b200.model as ours
    assert type(opt.model) is ours.ModelPose and node.ModelPose is ours.ModelPose
    assert opt.points.shape == (len(pts) - 1, 3)  # the NaN point was dropped by the drop-in codec
    topics = [t for t, _ in pub]
    assert topics.count("/odom") == 20 and topics.count("/tf") == 20 and topics.count("/camera/camera_info") == 20
    clouds = [m for t, m in pub if t == "/pts/rewards"]
    assert len(clouds) == 20
    # the debug cloud went through xyzi_array_to_pointcloud2 with the NUMPY (N,4) array the node builds
    last = clouds[-1]
    assert last.width == len(pts) - 1 and last.point_step == 16 and len(last.data) == 16 * last.width
    rec = np.frombuffer(last.data, np.float32).reshape(-1, 4)
    np.testing.assert_array_equal(rec[:, :3], pts[np.isfinite(pts).all(1)])
    assert last.is_dense == 1 and float(rec[:, 3].max()) > 0.0
    odom = [m for t, m in pub if t == "/odom"]
    ours_t = _f([odom[-1].pose.pose.position.x, odom[-1].pose.pose.position.y, odom[-1].pose.pose.position.z])
    ours_q = _f([getattr(odom[-1].pose.pose.orientation, k) for k in "xyzw"])
    assert np.abs(ours_t - np.array([6.0, 2.0, 0.0])).max() > 0.05  # Adam moved the pose

    ref_opt, ref_pub, _ = _run_pose_node(False, pts)
    assert type(ref_opt.model).__module__ == "model" and type(ref_opt.model) is not ours.ModelPose
    ref_odom = [m for t, m in ref_pub if t == "/odom"]
    assert len(ref_odom) == len(odom)
    for a, b in zip(odom, ref_odom):  # every published pose along the optimisation, not only the last
        ta = _f([a.pose.pose.position.x, a.pose.pose.position.y, a.pose.pose.position.z])
        tb = _f([b.pose.pose.position.x, b.pose.pose.position.y, b.pose.pose.position.z])
        qa, qb = (_f([getattr(m.pose.pose.orientation, k) for k in "xyzw"]) for m in (a, b))
        assert np.abs(ta - tb).max() < 2e-3 and np.abs(qa - qb).max() < 2e-3
    assert np.abs(ours_q).max() <= 1.0 + 1e-6


def _run_traj_node(dropin, pts, poses, rewards_cloud):
    with _Env(dropin) as rospy:
        saved = _standins() if dropin else None
        try:
            node = _load_node("trajectory_optimization")
            rospy.PARAMS.update({"traj_opt/opt_steps": 4, "traj_opt/lr_pose": 0.12, "traj_opt/lr_quat": 0.05})
            opt = node.TrajOpt(pc_topic="/cloud", input_path_topic="/path", publish_rewards_cloud=rewards_cloud,
                               device=torch.device("cpu"))
            opt.callback(_cloud_msg(pts), _path_msg(poses))
            return opt, list(rospy.PUBLISHED), node
        finally:
            if saved:
                _restore(saved)


def test_trajectory_optimization_node_runs_unchanged_and_matches_the_reference_node():
    pts, poses = _sample(3000)
    opt, pub, node = _run_traj_node(True, pts, poses, rewards_cloud=True)
    import trajectory_optimization_b200.model as ours
    assert node.ModelTraj is ours.ModelTraj
    paths = [m for t, m in pub if t == "/path/optimized"]
    clouds = [m for t, m in pub if t == "/cloud/rewards"]
    assert len(paths) == 1 and len(clouds) == 1 and len(paths[0].poses) == len(poses)
    rec = np.frombuffer(clouds[0].data, np.float32).reshape(-1, 4)
    assert clouds[0].point_step == 16 and rec.shape[0] == len(pts) - 1
    assert rec[:, 3].min() >= 0.5 and rec[:, 3].max() <= 1.0  # rewards, in the cloud's own order
    ours_xyz = np.array([_f([p.pose.position.x, p.pose.position.y, p.pose.position.z]) for p in paths[0].poses])
    ours_q = np.array([_f([getattr(p.pose.orientation, k) for k in "xyzw"]) for p in paths[0].poses])
    assert np.abs(ours_xyz - poses).max() > 0.05

    _, ref_pub, _ = _run_traj_node(False, pts, poses, rewards_cloud=False)
    ref_path = [m for t, m in ref_pub if t == "/path/optimized"][0]
    ref_xyz = np.array([_f([p.pose.position.x, p.pose.position.y, p.pose.position.z]) for p in ref_path.poses])
    ref_q = np.array([_f([getattr(p.pose.orientation, k) for k in "xyzw"]) for p in ref_path.poses])
    # 4 Adam steps on fp32 losses: the two arms see gradients that agree to ~1e-5; Adam's sign-like first steps turn
    # that into < 1e-3 on the waypoints
    assert np.abs(ours_xyz - ref_xyz).max() < 2e-3, np.abs(ours_xyz - ref_xyz).max()
    assert np.abs(np.abs(ours_q) - np.abs(ref_q)).max() < 2e-3


def test_publish_pointcloud_takes_what_the_reference_passes():
    """tools.publish_pointcloud with (N,3) and (N,4) NUMPY arrays, float64 (np.concatenate of fp32 points and rewards),
    and with a CPU tensor (src/tools.py:224-231 callers)."""
    with _Env(True) as rospy:
        saved = _standins()
        try:
            import tools
            g = np.random.default_rng(0)
            xyz = g.normal(size=(50, 3))
            xyzi = np.concatenate([xyz.astype(np.float32), g.random((50, 1))], axis=1)
            xyzi[3, 3] = np.inf
            tools.publish_pointcloud(xyz, "/a", rospy.Time.now(), "map")
            tools.publish_pointcloud(xyzi, "/b", rospy.Time.now(), "map")
            tools.publish_pointcloud(torch.from_numpy(xyz), "/c", rospy.Time.now(), "map")
            (ta, a), (tb, b), (tc, c) = rospy.PUBLISHED
            assert (ta, tb, tc) == ("/a", "/b", "/c")
            assert a.data == np.asarray(xyz, np.float32).tobytes() == c.data and a.is_dense == 1 and a.row_step == 50
            assert b.data == np.asarray(xyzi, np.float32).tobytes() and b.is_dense == 0 and b.point_step == 16
            assert [f.name for f in b.fields] == ["x", "y", "z", "i"] and [f.offset for f in b.fields] == [0, 4, 8, 12]
        finally:
            _restore(saved)
