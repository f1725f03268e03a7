"""Put this directory on sys.path AHEAD of the reference's `src/` and `from tools import ...`
(src/pose_optimization.py:21-27, src/trajectory_optimization.py:19-21, src/pc_processor.py:27)
resolves to the B200 implementation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from trajectory_optimization_b200.tools import *  # noqa: F401,F403,E402
from trajectory_optimization_b200.tools import (convexHull, denormalize, get_cam_frustum_pts,  # noqa: F401,E402
                                                hidden_pts_removal, hidden_pts_removal_o3d, load_intrinsics,
                                                multi_camera_visibility,
                                                publish_camera_info, publish_image, publish_odom, publish_path,
                                                publish_pointcloud, publish_pose, publish_tf_pose, render_pc_image,
                                                sphericalFlip, to_pose_stamped)
