"""Bit-exact parity of the binary frustum cull and the Katz HPR stage with the reference fixtures
and the CPU oracle (index sets must be identical on non-degenerate clouds)."""
import numpy as np
import pytest
import torch

from oracle import coverage_oracle as orc
from tests.conftest import load_golden

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
