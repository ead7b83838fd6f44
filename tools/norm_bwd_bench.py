"""InstanceNorm backward (fused kernel) at the C3 training shapes, timed alone with CUDA events"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200 import ops  # noqa: E402
from pbt_b200._native import ACT_RELU, FP16, P8  # noqa: E402

dt = FP16
for (n, c, hw) in ((80, 128, 80), (80, 32, 80), (80, 128, 40), (80, 128, 20)):
    x = P8.empty(n, c, hw, hw, dt)
    x.t.normal_()
    ga = P8.empty(n, c, hw, hw, dt)
    ga.t.normal_()
    dx = P8.empty(n, c, hw, hw, dt)
    scale = torch.rand((n, c), device="cuda") + 0.5
    shift = torch.randn((n, c), device="cuda")
    sums = torch.zeros((n, 2, c), device="cuda")
    run = lambda: ops.norm_bwd(x, dt, scale=scale, shift=shift, act=ACT_RELU, ga=ga, sums=sums, kmul=scale, count=hw * hw, dx=dx)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = 3.0 * n * c * hw * hw * 2
    print(f"norm_bwd {n}x{c}x{hw}x{hw}: {ms * 1e3:7.1f} us  {nbytes / ms / 1e6:7.0f} GB/s algorithmic ({nbytes / ms / 1e6 / 6544.3:.2f} of HBM peak)", flush=True)
