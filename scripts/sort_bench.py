#!/usr/bin/env python
"""Times the once-per-cloud ordering (`cov_spatial_sort` = bounding box, Morton keys, pair sort, gather) and the voxel-grid
filter on one GPU.  COV_B200_LIB selects another build of the library for A/B runs (e.g. round 2's CUB-based sort).
usage: sort_bench.py [tag [n ...]]"""
import ctypes
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
L = _lib.lib()
for n in sizes:
    gen = torch.Generator(device=dev).manual_seed(n)
    pts = torch.rand(n, 3, device=dev, generator=gen) * torch.tensor([40.0, 40.0, 5.0], device=dev)
    # the C ABI with buffers allocated once: no allocator traffic inside the timed region
    out = torch.empty_like(pts)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    wsb = L.cov_spatial_sort_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run_sort():
        _lib.check(L.cov_spatial_sort(pts.data_ptr(), n, out.data_ptr(), perm.data_ptr(), ws.data_ptr(), wsb, stream), "sort")

    best = min(timed(run_sort, reps=5)[0] for _ in range(3))
    chk = int(perm[:: max(n // 1000, 1)].long().sum())
    ok = bool(torch.equal(out[:1000], pts[perm[:1000].long()]))
    print(json.dumps({"lib": tag, "what": "cov_spatial_sort (C ABI, preallocated; best of 3 x 5 calls)", "n": n, "ms": best,
                      "perm_checksum": chk, "gather_ok": ok}), flush=True)
    if "cov_sort_pairs" in _lib.PROTOTYPES:
        keys0 = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, dtype=torch.int32, generator=gen)
        keys, vals = keys0.clone(), torch.arange(n, device=dev, dtype=torch.int32)
        wsb2 = L.cov_sort_pairs_workspace_bytes(n)
        ws2 = ws if wsb2 <= wsb else torch.empty(wsb2, dtype=torch.uint8, device=dev)

        def run_pairs():  # re-sorting sorted keys moves the same bytes through the same kernels
            _lib.check(L.cov_sort_pairs(keys.data_ptr(), vals.data_ptr(), n, 0, 32, ws2.data_ptr(), wsb2, stream), "pairs")

        keys.copy_(keys0)
        first = timed(lambda: (keys.copy_(keys0), run_pairs()), reps=3)[0]
        copy_ms = timed(lambda: keys.copy_(keys0), reps=3)[0]
        print(json.dumps({"lib": tag, "what": "cov_sort_pairs, 32 random key bits", "n": n, "ms": first - copy_ms}), flush=True)
        del keys0, keys, vals, ws2
    del pts, out, perm, ws
    torch.cuda.empty_cache()

n = 10_000_000
gen = torch.Generator(device=dev).manual_seed(7)
pts = (torch.rand(n, 3, device=dev, generator=gen) - 0.5) * torch.tensor([40.0, 40.0, 5.0], device=dev)
ms, out = timed(lambda: tools.voxel_grid_filter(pts, 0.1, "z", -2.5, 2.5))
print(json.dumps({"lib": tag, "what": "voxel_grid_filter leaf 0.1", "n": n, "ms": ms, "voxels": int(out.shape[0]),
                  "checksum": float(out.double().sum())}), flush=True)
