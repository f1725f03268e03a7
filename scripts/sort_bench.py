#!/usr/bin/env python
"""Times the once-per-cloud ordering (`cov_spatial_sort` = bounding box, Morton keys, pair sort, gather) and the voxel-grid
filter on one GPU.  COV_B200_LIB selects another build of the library for A/B runs (e.g. round 2's CUB-based sort).
usage: sort_bench.py [tag [n ...]]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_optimization_b200 import _lib  # noqa: E402

if os.environ.get("COV_B200_LIB"):  # an older build may lack the newest entry points
    import ctypes
    h = ctypes.CDLL(os.environ["COV_B200_LIB"])
    for name in list(_lib.PROTOTYPES):
        if not hasattr(h, name):
            _lib.PROTOTYPES.pop(name)
from trajectory_optimization_b200 import ops, tools  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "own"
dev = torch.device("cuda:0")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


sizes = [int(a) for a in sys.argv[2:]] or [12_500_000, 100_000_000]
for n in sizes:
    gen = torch.Generator(device=dev).manual_seed(n)
    pts = torch.rand(n, 3, device=dev, generator=gen) * torch.tensor([40.0, 40.0, 5.0], device=dev)
    ms, (out, perm) = timed(lambda: ops.spatial_sort(pts))
    chk = int(perm[:: max(n // 1000, 1)].long().sum())
    ok = bool(torch.equal(out[:1000], pts[perm[:1000].long()]))
    print(json.dumps({"lib": tag, "what": "cov_spatial_sort", "n": n, "ms": ms, "GB_per_s_of_20B_per_point": n * 20 / ms / 1e6,
                      "perm_checksum": chk, "gather_ok": ok}), flush=True)
    if hasattr(ops, "sort_pairs") and "cov_sort_pairs" in _lib.PROTOTYPES:
        keys = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, dtype=torch.int32, generator=gen)
        vals = torch.arange(n, device=dev, dtype=torch.int32)
        ms2, _ = timed(lambda: ops.sort_pairs(keys, vals, 0, 32))
        print(json.dumps({"lib": tag, "what": "cov_sort_pairs 32 bits (incl. two clones)", "n": n, "ms": ms2}), flush=True)
        del keys, vals
    del pts, out, perm
    torch.cuda.empty_cache()

n = 10_000_000
gen = torch.Generator(device=dev).manual_seed(7)
pts = (torch.rand(n, 3, device=dev, generator=gen) - 0.5) * torch.tensor([40.0, 40.0, 5.0], device=dev)
ms, out = timed(lambda: tools.voxel_grid_filter(pts, 0.1, "z", -2.5, 2.5))
print(json.dumps({"lib": tag, "what": "voxel_grid_filter leaf 0.1", "n": n, "ms": ms, "voxels": int(out.shape[0]),
                  "checksum": float(out.double().sum())}), flush=True)
