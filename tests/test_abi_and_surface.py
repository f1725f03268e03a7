"""CPU checks of the drop-in boundary: the shared library loads and exports every symbol the header declares
(no compute calls), the ctypes table covers them, and the Python surface has the reference's names/signatures."""
import inspect
import os
import re
import sys

import pytest

from tests.conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "coverage_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cov_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from trajectory_optimization_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 19
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/coverage_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)
    # host-only queries work without a GPU
    assert L.cov_version() >= 100
    assert L.cov_traj_max_poses() >= 1024
    assert L.cov_traj_workspace_bytes(1000, 320) > 0 and L.cov_pose_workspace_bytes(1000) > 0
    assert L.cov_hpr_hull_workspace_bytes(1_000_000) > 16_000_000 and L.cov_cull_workspace_bytes(1000) > 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "trajectory_optimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"


def test_cpu_tensors_are_rejected_not_emulated():
    import torch
    from trajectory_optimization_b200 import model, tools
    K, W, H = tools.load_intrinsics(torch.device("cpu"))
    m = model.ModelPose(torch.rand(10, 3), torch.zeros(1, 3), torch.tensor([[1.0, 0, 0, 0]]), K, W, H,
                        device=torch.device("cpu"))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        m()


REF_SRC = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference checkout only exists in the build container")
def test_python_surface_matches_reference_signatures():
    import numpy as np
    np.float = float
    sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), REF_SRC]
    try:
        for mod in ("model", "tools"):
            sys.modules.pop(mod, None)
        import model as ref_model
        import tools as ref_tools
    finally:
        del sys.path[:2]
        for mod in ("model", "tools"):
            sys.modules.pop(mod, None)
    from trajectory_optimization_b200 import model, tools

    def params(fn):
        return [(p.name, p.default if p.default is not inspect._empty else None)
                for p in inspect.signature(fn).parameters.values()]

    for name in ("get_dist_mask", "get_fov_mask", "to_camera_frame", "length_calc", "mean_angle_calc"):
        assert [p[0] for p in params(getattr(model, name))] == [p[0] for p in params(getattr(ref_model, name))], name
    for cls in ("ModelPose", "ModelTraj"):
        ours, ref = params(getattr(model, cls).__init__), params(getattr(ref_model, cls).__init__)
        assert [p[0] for p in ours][:len(ref)] == [p[0] for p in ref], cls           # ours only appends optional kwargs
        assert [p[1] for p in ours][:len(ref) - 1] == [p[1] for p in ref][:-1], cls  # same defaults (device aside)
        assert [p[0] for p in params(getattr(model, cls).forward)] == [p[0] for p in params(getattr(ref_model, cls).forward)]
    for name in ("sphericalFlip", "convexHull", "hidden_pts_removal", "get_cam_frustum_pts", "load_intrinsics",
                 "render_pc_image", "publish_image", "publish_odom", "publish_pointcloud", "publish_tf_pose",
                 "publish_camera_info", "to_pose_stamped", "publish_pose", "publish_path", "denormalize",
                 "hidden_pts_removal_o3d"):
        assert [p[0] for p in params(getattr(tools, name))] == [p[0] for p in params(getattr(ref_tools, name))], name
