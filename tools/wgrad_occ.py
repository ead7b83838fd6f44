"""wgrad kernel at the C3 training shapes: time vs the pixel-split factor (CTA budget multiplier)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200 import ops  # noqa: E402
from pbt_b200._native import FP16, P8  # noqa: E402

dt = FP16
N = 80
cases = [("res 128->128 3x3 @20x20", 128, 128, 3, 20), ("up2 256->128 3x3 @40x40", 256, 128, 3, 40),
         ("up1 192->128 3x3 @80x80", 192, 128, 3, 80), ("conv11 176->64 7x7 @80x80", 176, 64, 7, 80),
         ("smooth 64->64 3x3 @80x80", 64, 64, 3, 80), ("initial 16->32 7x7 @80x80", 16, 32, 7, 80),
         ("down1 s2d 128->64 2x2 @40x40", 128, 64, 2, 40), ("down2 s2d 256->128 2x2 @20x20", 256, 128, 2, 20)]
for name, cin, cout, k, hw in cases:
    x = P8.empty(N, cin, hw, hw, dt)
    x.t.normal_()
    g = P8.empty(N, cout, hw, hw, dt)
    g.t.normal_()
    dw = torch.zeros((k * k, cin, cout), device="cuda")
    for mult in (4, 2, 1, 8):
        run = lambda: ops.conv_wgrad(x, g, k, k, k // 2, k // 2, dt, dw, debug_flags=mult << 8)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.0 * N * hw * hw * k * k * cin * cout / ms / 1e9
        print(f"{name}: mult={mult}: {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s", flush=True)
