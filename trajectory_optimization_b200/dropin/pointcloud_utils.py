"""Drop-in for the reference's `src/pointcloud_utils.py` (module name `pointcloud_utils`): the codec entry points the
nodes import (src/pose_optimization.py:28, src/pc_processor.py:10), backed by libcovb200.so."""
from trajectory_optimization_b200.pointcloud_utils import (  # noqa: F401
    pointcloud2_to_xyz_array, pointcloud2_to_xyz_tensor, xyz_array_to_pointcloud2, xyzi_array_to_pointcloud2)
