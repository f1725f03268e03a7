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


def test_determinant_sign_falls_back_to_double_double(tmp_path):
    """hull_det_sign: cases of fp32-valued rows whose determinant is far below the fp64 filter's bound (1e-15 of the
    permanent) but not zero must get their exact sign from the double-double stage; exactly singular ones stay 0."""
    from fractions import Fraction
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "det_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "det_host.cpp")], check=True)
    gen = np.random.default_rng(4)
    cases, exact = [], []
    for i in range(60):
        # integer rows of magnitude 2^25 (exact in fp64): b, c = (c0, b1 + 1, b2 + 1), a = b + c + (1, 0, 0), so that
        # det = b1 (b2 + 1) - b2 (b1 + 1) = b1 - b2 = k exactly, while the permanent is ~2^76: |det| / perm ~ 2^-70
        k = 0 if i % 3 == 0 else int(gen.integers(1, 1000)) * (1 if i % 2 else -1)
        b1 = int(gen.integers(2 ** 24, 2 ** 25))
        bb = [int(gen.integers(2 ** 24, 2 ** 25)), b1, b1 - k]
        cc = [int(gen.integers(2 ** 24, 2 ** 25)), bb[1] + 1, bb[2] + 1]
        aa = [bb[0] + cc[0] + 1, bb[1] + cc[1], bb[2] + cc[2]]
        rows = [[Fraction(x) for x in r] for r in (aa, bb, cc)]
        det = (rows[0][0] * (rows[1][1] * rows[2][2] - rows[1][2] * rows[2][1])
               + rows[0][1] * (rows[1][2] * rows[2][0] - rows[1][0] * rows[2][2])
               + rows[0][2] * (rows[1][0] * rows[2][1] - rows[1][1] * rows[2][0]))
        assert det == k
        # the fp64 filter alone cannot decide these
        A, B, Cc = (np.array(r, np.float64) for r in (aa, bb, cc))
        perm = (abs(A[0]) * (abs(B[1] * Cc[2]) + abs(B[2] * Cc[1])) + abs(A[1]) * (abs(B[2] * Cc[0]) + abs(B[0] * Cc[2]))
                + abs(A[2]) * (abs(B[0] * Cc[1]) + abs(B[1] * Cc[0])))
        assert abs(k) < 1e-15 * perm
        cases.append(np.array(aa + bb + cc, np.float64))
        exact.append(0 if k == 0 else (1 if k > 0 else -1))
    text = "\n".join(" ".join(repr(float(x)) for x in row) for row in cases) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    assert got == exact
    assert exact.count(0) >= 10 and sum(1 for e in exact if e != 0) >= 10   # both kinds were exercised
