"""Put this directory on sys.path AHEAD of the reference's `src/` and
`from model import ModelPose, ModelTraj` (src/pose_optimization.py:10,
src/trajectory_optimization.py:9) resolves to the B200 implementation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from trajectory_optimization_b200.model import *  # noqa: F401,F403,E402
from trajectory_optimization_b200.model import (ModelPose, ModelTraj, get_dist_mask, get_fov_mask,  # noqa: F401,E402
                                                length_calc, mean_angle_calc, to_camera_frame)
