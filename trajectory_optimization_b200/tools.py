"""Drop-in for the reference's `src/tools.py`: same public names and signatures.

Hot-path functions (Katz hidden-point removal, binary frustum cull, intrinsics) run on the
sm_100a kernels of libcovb200.so; the ROS publishers are thin message builders that import ROS
lazily, so the module loads on a box without ROS.  Citations: reference repository root.
"""
import numpy as np
import torch

from . import ops


# ------------------------------------------------------------------------------------------
# Katz hidden-point removal (src/tools.py:38-85)
# ------------------------------------------------------------------------------------------
def sphericalFlip(points, device, param):
    """src/tools.py:38-53 — fp32 spherical flip, bit-identical to the reference's torch CPU result."""
    flipped, _ = ops.spherical_flip(points.to(device), param)
    return flipped


class _Hull:
    """What callers use of scipy.spatial.ConvexHull here: `.vertices` (ascending int32, input order)."""

    def __init__(self, vertices, n_exact):
        self.vertices = vertices
        self.n_exact_fallback = n_exact


def convexHull(points, device):
    """src/tools.py:56-64 — hull of `points` plus the origin (appended as index N)."""
    mask, origin_is_vertex, n_exact = ops.hpr_hull_mask(points.to(device))
    idx = torch.nonzero(mask, as_tuple=False).reshape(-1).to(torch.int32).cpu().numpy()
    if origin_is_vertex:
        idx = np.concatenate([idx, np.array([points.shape[0]], dtype=np.int32)])
    return _Hull(idx, n_exact)


def hidden_pts_removal(pts: torch.Tensor, device, R_param: int = 2):
    """src/tools.py:67-85.  Returns (visible points (M,3), visibility mask float32 (N,)).
    Keeps the reference's `vertices[:-1]`: the origin is dropped when it is a hull vertex,
    otherwise the highest-index visible point is.  Decisions the hull stage could not certify (exactly coplanar or
    duplicate points: degenerate input, where Qhull's own answer depends on its joggle options) raise a warning."""
    pts = pts.to(device)
    flipped, _ = ops.spherical_flip(pts, R_param)
    mask, origin_is_vertex, n_uncertified = ops.hpr_hull_mask(flipped)
    if n_uncertified:
        import warnings
        warnings.warn(f"hidden_pts_removal: {n_uncertified} hull decision(s) could not be certified even at ~100 bits "
                      "(degenerate input: duplicate or exactly coplanar points); the visible set may differ from Qhull's "
                      "on those points", RuntimeWarning, stacklevel=2)
    idx = torch.nonzero(mask, as_tuple=False).reshape(-1)
    if not origin_is_vertex:
        idx = idx[:-1]
    visible_mask = torch.zeros(pts.size()[0], device=device)
    visible_mask[idx] = 1
    return pts[idx, :], visible_mask


@torch.no_grad()
def multi_camera_visibility(points, cam_trans, cam_quats, intrins, img_height, img_width, min_dist=1.0, max_dist=10.0,
                            R_param=2, device=None, group=None):
    """The per-camera visibility pipeline of the reference's point-cloud processor for C cameras in one call:
    ego -> camera frame (src/pc_processor.py:64-70), binary frustum cull (:72-83), Katz hidden-point removal
    (:178 -> src/tools.py:67-85), all on the device.  `cam_trans` (C,3), `cam_quats` (C,4) (w,x,y,z) are the camera poses
    in the cloud's frame (what `tf.lookupTransform(pc_frame, cam_frame)` returns, reordered); `intrins` (3,3) or (C,3,3).

    Cameras are independent, so with a `group` (one process per GPU, every rank holding the cloud) they are REPLICATED
    work: rank r takes cameras r, r + world, ....  Returns a list with one dict per LOCAL camera:
    camera (index), frustum_idx (int64 indices into `points`, ascending), frustum_points (M,3) camera frame,
    visible_idx (int64 indices into `points`), visible_points (V,3) camera frame."""
    from .model import to_camera_frame
    device = torch.device(device) if device is not None else (points.device if isinstance(points, torch.Tensor) and points.is_cuda
                                                               else torch.device("cuda", torch.cuda.current_device()))
    pts = torch.as_tensor(points, dtype=torch.float32).to(device)
    T = torch.as_tensor(cam_trans, dtype=torch.float32).to(device).reshape(-1, 3)
    Q = torch.as_tensor(cam_quats, dtype=torch.float32).to(device).reshape(-1, 4)
    K = torch.as_tensor(intrins, dtype=torch.float32).to(device)
    C = T.shape[0]
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    out = []
    for c in range(rank, C, world):
        cam = to_camera_frame(pts, Q[c:c + 1], T[c:c + 1]).contiguous()
        Kc = K if K.dim() == 2 else K[c]
        idx, _, _ = ops.frustum_cull(cam, Kc, img_width, img_height, min_dist, max_dist)
        fr = cam[idx]
        if fr.shape[0] >= 4:
            vis_pts, vis_mask = hidden_pts_removal(fr, device, R_param)
            vis_local = torch.nonzero(vis_mask, as_tuple=False).reshape(-1)
        else:  # fewer than a tetrahedron: Qhull raises in the reference; nothing can occlude anything
            vis_pts, vis_local = fr, torch.arange(fr.shape[0], device=device)
        out.append(dict(camera=c, frustum_idx=idx, frustum_points=fr, visible_idx=idx[vis_local], visible_points=vis_pts))
    return out


def hidden_pts_removal_o3d(pts):
    """src/tools.py:88-119 — Open3D variant; needs open3d, unused by the nodes."""
    import open3d as o3d
    flip = np.diag([1, -1, -1])
    p = (flip @ np.asarray(pts).T).T
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(p)
    diameter = np.linalg.norm(np.asarray(pcd.get_max_bound()) - np.asarray(pcd.get_min_bound()))
    if diameter > 0:
        _, pt_map = pcd.hidden_point_removal([0, 0, 0.0], diameter * 100)
        p = np.asarray(pcd.select_by_index(pt_map).points)
    return (flip.T @ p.T).T


# ------------------------------------------------------------------------------------------
# binary frustum cull (src/tools.py:176-187)
# ------------------------------------------------------------------------------------------
def get_cam_frustum_pts(points, img_height, img_width, intrins, min_dist=1.0, max_dist=10.0):
    """`points` is 3xN (camera frame) like the reference; returns (points (M,3), dist_mask, fov_mask)."""
    pts_nx3 = points[:3].t().contiguous()
    idx, dist_mask, fov_mask = ops.frustum_cull(pts_nx3, intrins, img_width, img_height, min_dist, max_dist)
    return pts_nx3[idx].to(points.dtype), dist_mask, fov_mask


def load_intrinsics(device=torch.device("cuda:0")):
    """src/tools.py:320-325."""
    width, height = 1232.0, 1616.0
    K = torch.tensor([[758.03967, 0.0, 621.46572],
                      [0.0, 761.62359, 756.86402],
                      [0.0, 0.0, 1.0]], dtype=torch.float32).to(device)
    return K, width, height


def denormalize(x, eps=1e-6):
    """src/tools.py:190-196 — 2..98 percentile stretch to [0, 1]."""
    lo, hi = np.percentile(x, 2), np.percentile(x, 98)
    return ((x - lo) / max(hi - lo, eps)).clip(0, 1)


def render_pc_image(verts, K, height, width, R=None, T=None, device=torch.device("cuda"), gamma=1.0e-1,
                    znear=1.0, zfar=10.0):
    """src/tools.py:122-173 — debug rendering through pytorch3d's Pulsar renderer.  Visualisation
    only and outside the coverage hot path; needs pytorch3d, which this build does not ship."""
    try:
        from pytorch3d.renderer import (PerspectiveCameras, PointsRasterizationSettings, PointsRasterizer,
                                        PulsarPointsRenderer)
        from pytorch3d.structures import Pointclouds
    except ImportError as e:  # pragma: no cover
        raise NotImplementedError("render_pc_image needs pytorch3d (debug visualisation, out of scope here)") from e
    rgb = verts - torch.min(verts)
    rgb = rgb / torch.max(rgb).to(device)
    cloud = Pointclouds(points=[verts], features=[rgb])
    R = torch.eye(3).unsqueeze(0).to(device) if R is None else R
    T = torch.zeros(1, 3).to(device) if T is None else T
    cameras = PerspectiveCameras(R=R, T=T, K=K, device=device)
    radius = 0.03 * torch.ones(verts.size()[0], dtype=torch.float32, device=device)
    settings = PointsRasterizationSettings(image_size=(width, height), radius=radius, points_per_pixel=1)
    renderer = PulsarPointsRenderer(rasterizer=PointsRasterizer(cameras=cameras, raster_settings=settings)).to(device)
    return renderer(cloud, gamma=(gamma,), znear=(znear,), zfar=(zfar,), radius_world=True,
                    bg_col=torch.ones((3,), dtype=torch.float32, device=device))[0]


# ------------------------------------------------------------------------------------------
# ROS message builders / publishers (src/tools.py:199-317).  ROS is imported on first use.
# ------------------------------------------------------------------------------------------
def _set_xyz(dst, v):
    dst.x, dst.y, dst.z = v[0], v[1], v[2]


def _set_quat_xyzw(dst, q):
    assert len(q) == 4
    dst.x, dst.y, dst.z, dst.w = q[0], q[1], q[2], q[3]


def publish_image(img, topic="/image/compressed"):
    import rospy
    from cv_bridge import CvBridge
    from sensor_msgs.msg import Image
    msg = CvBridge().cv2_to_imgmsg(np.uint8(255 * denormalize(img)), "bgr8")
    rospy.Publisher(topic, Image, queue_size=1).publish(msg)


def publish_odom(pose, quat, frame="/odom", topic="/odom_0"):
    import rospy
    from nav_msgs.msg import Odometry
    assert len(pose) == 3
    msg = Odometry()
    msg.header.stamp = rospy.Time.now()
    msg.header.frame_id = frame
    _set_xyz(msg.pose.pose.position, pose)
    _set_quat_xyzw(msg.pose.pose.orientation, quat)
    rospy.Publisher(topic, Odometry, queue_size=1).publish(msg)


def publish_pointcloud(points, topic_name, stamp, frame_id):
    import rospy
    from sensor_msgs.msg import PointCloud2
    from .pointcloud_utils import xyz_array_to_pointcloud2, xyzi_array_to_pointcloud2
    build = {3: xyz_array_to_pointcloud2, 4: xyzi_array_to_pointcloud2}[points.shape[1]]
    rospy.Publisher(topic_name, PointCloud2, queue_size=1).publish(build(points, stamp=stamp, frame_id=frame_id))


def publish_tf_pose(pose, quat, child_frame_id, frame_id="world"):
    import rospy
    import tf2_ros
    from geometry_msgs.msg import TransformStamped
    assert len(pose) == 3
    t = TransformStamped()
    t.header.stamp = rospy.Time.now()
    t.header.frame_id = frame_id
    t.child_frame_id = child_frame_id
    _set_xyz(t.transform.translation, pose)
    _set_quat_xyzw(t.transform.rotation, quat)
    tf2_ros.TransformBroadcaster().sendTransform(t)


def publish_camera_info(image_width=1232, image_height=1616,
                        K=(758.03967, 0.0, 621.46572, 0.0, 761.62359, 756.86402, 0.0, 0.0, 1.0),
                        D=(-0.20571, 0.04103, -0.00101, 0.00098, 0.0),
                        R=(1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0),
                        P=(638.81494, 0.0, 625.98561, 0.0, 0.0, 585.79797, 748.57858, 0.0, 0.0, 0.0, 1.0, 0.0),
                        topic_name="/camera_info", frame_id="camera_frame", distortion_model="plumb_bob"):
    import rospy
    from sensor_msgs.msg import CameraInfo
    msg = CameraInfo()
    msg.header.frame_id = frame_id
    msg.header.stamp = rospy.Time.now()
    msg.width, msg.height = image_width, image_height
    msg.K, msg.D, msg.R, msg.P = list(K), list(D), list(R), list(P)
    msg.distortion_model = distortion_model
    rospy.Publisher(topic_name, CameraInfo, queue_size=1).publish(msg)


def to_pose_stamped(pose, quat, stamp=None, frame_id="world"):
    import rospy
    from geometry_msgs.msg import PoseStamped
    assert len(pose) == 3
    msg = PoseStamped()
    msg.header.seq = 1
    msg.header.frame_id = frame_id
    _set_xyz(msg.pose.position, pose)
    _set_quat_xyzw(msg.pose.orientation, quat)
    msg.header.stamp = rospy.Time.now()  # the reference overwrites any stamp passed in (src/tools.py:284)
    return msg


def publish_pose(pose, quat, topic_name, stamp=None, frame_id="world"):
    import rospy
    from geometry_msgs.msg import PoseStamped
    msg = to_pose_stamped(pose, quat, stamp=stamp, frame_id=frame_id)
    rospy.Publisher(topic_name, PoseStamped, queue_size=1).publish(msg)


def publish_path(path_list, orient_list=None, topic_name="/path", frame_id="world"):
    import rospy
    from nav_msgs.msg import Path
    path = Path()
    orients = [[0, 0, 0, 1]] * len(path_list) if orient_list is None else orient_list
    for pose, orient in zip(path_list, orients):
        msg = to_pose_stamped(pose, orient, frame_id=frame_id)
        path.header = msg.header
        path.poses.append(msg)
    rospy.Publisher(topic_name, Path, queue_size=1).publish(path)


# ------------------------------------------------------------------------------------------
# voxel-grid filter (launch/voxels_filtering.launch:8-21: the pcl/VoxelGrid nodelet in front of the optimiser)
# ------------------------------------------------------------------------------------------
@torch.no_grad()
def voxel_grid_filter(points, leaf_size=0.1, filter_field_name="z", filter_limit_min=-2.5, filter_limit_max=2.5):
    """Downsample a cloud to one centroid per occupied voxel, after dropping non-finite points and those outside
    [filter_limit_min, filter_limit_max] along `filter_field_name` (None: no pass-through).  Defaults are the launch
    file's.  Returns an (M,3) fp32 CUDA tensor, voxels in ascending index order (x fastest), as pcl::VoxelGrid emits them."""
    from . import _lib
    L = _lib.lib()
    pts = ops._dev_f32(points, what="points")
    n = pts.shape[0]
    if n == 0:
        return pts.new_zeros((0, 3))
    axis = -1 if filter_field_name is None else {"x": 0, "y": 1, "z": 2}[filter_field_name]
    out = torch.empty_like(pts)
    cnt = torch.zeros(1, dtype=torch.int64, device=pts.device)
    info = torch.zeros(8, dtype=torch.int32, device=pts.device)
    ws_bytes = L.cov_voxel_grid_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
    ops._call("cov_voxel_grid", pts, pts.data_ptr(), n, float(leaf_size), axis, float(filter_limit_min),
              float(filter_limit_max), out.data_ptr(), cnt.data_ptr(), info.data_ptr(), ws.data_ptr(), ws_bytes)
    info_h = info.tolist()
    if info_h[6]:
        raise RuntimeError("voxel_grid_filter: leaf size is too small for the input dataset (integer voxel indices would "
                           "overflow) — pcl::VoxelGrid refuses the same input")
    return out[:int(cnt.item())]
