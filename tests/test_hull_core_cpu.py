"""The per-point extremeness test the CUDA hull kernel runs (csrc/hull_core.h, host+device code) compiled for
the host and checked against the Qhull vertex sets recorded from the reference (no GPU needed)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import coverage_oracle as orc
from tests.conftest import ROOT, load_golden


@pytest.fixture(scope="module")
def hull_host(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("hull") / "hull_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "hull_host.cpp")],
                   check=True)
    return exe


def _run(exe, flipped, tmp_path, *extra):
    src, dst = str(tmp_path / "pts.f32"), str(tmp_path / "mask.u8")
    np.ascontiguousarray(flipped, dtype=np.float32).tofile(src)
    out = subprocess.run([exe, src, str(len(flipped)), dst, *extra], check=True, capture_output=True, text=True).stdout
    stats = dict(kv.split("=") for kv in out.split())
    return np.flatnonzero(np.fromfile(dst, dtype=np.uint8)), {k: int(v) for k, v in stats.items()}


@pytest.mark.parametrize("name,origin_vertex", [("hpr_shell_small", 0), ("hpr_shell", 0), ("hpr_halfspace", 1),
                                                ("hpr_sample", 0)])
def test_extreme_points_equal_qhull_vertices(name, origin_vertex, hull_host, tmp_path):
    g = load_golden(name)
    flipped, _, _ = orc.spherical_flip(g["in_points"], int(g["in_R_param"]))
    idx, st = _run(hull_host, flipped, tmp_path)
    assert st["origin_vertex"] == origin_vertex and st["origin_cert"] == 1
    assert st["inside_uncert"] == 0 and st["overflow"] == 0 and st["extreme_uncert"] == 0  # every decision certified
    vis = idx if origin_vertex else idx[:-1]  # the reference's vertices[:-1] (src/tools.py:79)
    assert np.array_equal(vis, g["out_idx"])


@pytest.mark.parametrize("G", ["1", "3", "40"])
def test_grid_resolution_does_not_change_the_answer(G, hull_host, tmp_path):
    g = load_golden("hpr_shell_small")
    flipped, _, _ = orc.spherical_flip(g["in_points"], 2)
    idx, st = _run(hull_host, flipped, tmp_path, G)
    assert np.array_equal(idx[:-1], g["out_idx"]) and st["G"] == int(G)
