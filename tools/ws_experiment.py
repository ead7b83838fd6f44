"""weight-stationary MMA experiment (conv debug bit 5): correctness on the conv test cases, then timing at inference shapes"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as gc  # noqa: E402
from test_gpu_ops import CONV_CASES  # noqa: E402
from pbt_b200 import ops  # noqa: E402
from pbt_b200._native import ACT_RELU, FP16, P8  # noqa: E402

bad = 0
n = 0
for case in CONV_CASES:
    if case.get("T", 1) < 2 or case.get("pair") or case.get("cout") not in (64, 128, 256):
        continue
    ok, err, msg = gc.check_conv(**dict(case, debug_flags=32))
    n += 1
    if not ok:
        bad += 1
        print("FAIL", case, err, msg)
print(f"weight-stationary correctness: {n - bad}/{n} cases ok", flush=True)
dt = FP16
N = 4
for name, cin, cout, k, h, w, T, blk, cps in (("res 128->128 3x3 @270x480", 128, 128, 3, 270, 480, 2, 32, 0),
                                             ("up1-like 192->128 3x3 @1080p", 192, 128, 3, 1080, 1920, 2, 32, 0),
                                             ("smooth 64->64 3x3 @1080p cps4", 64, 64, 3, 1080, 1920, 2, 16, 4),
                                             ("conv11 176->64 7x7 @1080p T3 (non-pair)", 176, 64, 7, 1080, 1920, 3, 32, 0)):
    x = P8.empty(N, cin, h, w, dt)
    x.t.normal_()
    wt = torch.randn((cout, cin, k, k), device="cuda") * 0.05
    bias = torch.randn((cout,), device="cuda")
    wp = ops.pack_conv_weight(wt, cin, blk, dt)
    ref = None
    for flags in (0, 32):
        out = P8.empty(N, cout, h, w, dt)
        run = lambda: ops.conv_fwd(x, wp, cout, k, k, k // 2, k // 2, dt, blk_c=blk, tiles_per_cta=T, out=out, bias=bias, act=ACT_RELU,  # noqa: E731
                                   ctas_per_sm=cps, debug_flags=flags)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        diff = 0.0 if ref is None else (out.t.float() - ref).abs().max().item()
        if ref is None:
            ref = out.t.float().clone()
        print(f"{name}: ws={flags != 0}: {ms * 1e3:8.1f} us  {2.0 * N * h * w * k * k * cin * cout / ms / 1e9:7.1f} TFLOP/s  maxdiff {diff:.3g}", flush=True)
