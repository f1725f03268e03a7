#!/usr/bin/env python
"""FP32 issue-rate probes on the current GPU: scalar FFMA, packed FFMA2 (fma.rn.f32x2), MUFU.EX2."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trajectory_optimization_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
sink = torch.zeros(1, device=dev)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
iters = 8192
for name in ("cov_probe_fma", "cov_probe_fma2", "cov_probe_ex2"):
    fn = getattr(L, name)
    fn(iters, sink.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops = 0
    for _ in range(5):
        ops += fn(iters, sink.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {ops / (ms * 1e-3) / 1e12:.2f} T op/s ({'x2 flop' if 'fma' in name else 'ex2'})", flush=True)
