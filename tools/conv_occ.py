"""default (2 CTAs/SM, 8 epilogue warps) vs small-footprint (4 CTAs/SM, 4 epilogue warps) conv configurations"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200 import ops  # noqa: E402
from pbt_b200._native import ACT_RELU, FP16, P8  # noqa: E402

dt = FP16
DBG = int(os.environ.get("PBT_DEBUG_FLAGS", "0"))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cases = [("smooth 64->64 3x3 @1080p", 64, 64, 3, 1080, 1920, 32, [(3, 0), (2, 4.16), (2, "pair4")]),
         ("initial 16->32 7x7 @1080p", 16, 32, 7, 1080, 1920, 16, [(3, 0), (3, 4), (2, 4)]),
         ("down1 s2d 128->64 2x2 @540p", 128, 64, 2, 540, 960, 32, [(2, 0), (2, 4), (3, 0)]),
         ("res 128->128 3x3 @270x480", 128, 128, 3, 270, 480, 32, [(2, 0), (2, "pair")]),
         ("up1-like 192->128 3x3 @1080p", 192, 128, 3, 1080, 1920, 32, [(2, 0), (2, "pair")]),
         ("conv11 176->64 7x7 @1080p", 176, 64, 7, 1080, 1920, 32, [(3, 0), (3, "pair"), (2, "pair4")])]
if len(sys.argv) > 2 and sys.argv[2] == "train":   # C3 training shapes: 80 patches of 80x80 (N = batch)
    N = 80
    cases = [("res 128->128 3x3 @20x20", 128, 128, 3, 20, 20, 32, [(1, 0), (1, 4), (2, "bt"), (3, "bt")]),
             ("down2 s2d 256->128 2x2 @20x20", 256, 128, 2, 20, 20, 32, [(1, 0), (1, 4), (2, "bt")]),
             ("up2 256->128 3x3 @40x40", 256, 128, 3, 40, 40, 32, [(2, 0), (1, 4), (2, "bt")]),
             ("down1 s2d 128->64 2x2 @40x40 ", 128, 64, 2, 40, 40, 32, [(2, 4), (2, "bt"), (3, "bt")]),
             ("up2 dgrad-like 128->128 3x3 @40x40", 128, 128, 3, 40, 40, 32, [(2, 0), (2, "bt")]),
             ("up1 192->128 3x3 @80x80", 192, 128, 3, 80, 80, 32, [(2, 0), (2, "pair"), (1, "pair")]),
             ("conv11 176->64 7x7 @80x80", 176, 64, 7, 80, 80, 32, [(2, 0), (2, "pair"), (1, "pair"), (2, "pair4")]),
             ("conv11 dgrad 64->160 7x7 @80x80", 64, 160, 7, 80, 80, 32, [(2, 0), (1, 0), (1, "pair")]),
             ("up1 dgrad 128->192 3x3 @80x80", 128, 192, 3, 80, 80, 32, [(2, 0), (1, 0), (1, "pair")]),
             ("up2 dgrad 128->256 3x3 @40x40", 128, 256, 3, 40, 40, 32, [(1, 0), (2, 0), (1, "pair")]),
             ("smooth 64->64 3x3 @80x80", 64, 64, 3, 80, 80, 32, [(2, 0), (2, 4), (2, 4.16), (1, 4)]),
             ("initial 16->32 7x7 @80x80", 16, 32, 7, 80, 80, 16, [(2, 0), (2, 4), (3, 4)]),
             ("down1 s2d 128->64 2x2 @40x40", 128, 64, 2, 40, 40, 32, [(2, 0), (2, 4), (1, 4)]),
             ("down1 dgrad 64->128 2x2 @40x40", 64, 128, 2, 40, 40, 32, [(2, 0), (1, 4), (1, 0)]),
             ("down2 dgrad 128->256 2x2 @20x20", 128, 256, 2, 20, 20, 32, [(1, 0), (2, 0), (1, "pair")])]
for name, cin, cout, k, h, w, blk, cfgs in cases:
    x = P8.empty(N, cin, h, w, dt)
    x.t.normal_()
    wt = torch.randn((cout, cin, k, k), device="cuda") * 0.05
    bias = torch.randn((cout,), device="cuda")
    ref = None
    for T, cps in cfgs:
        try:
            b = blk
            bt = cps == "bt"
            if bt:
                cps = 0
            pair = cps in ("pair", "pair4")
            if pair:
                cps = 4.16 if cps == "pair4" else 0
            if cps != int(cps):
                b, cps = 16, int(cps)
            if cps == 4 and k == 7 and cin > 16:
                b = 16
            wp = ops.pack_conv_weight(wt, cin, b, dt, pair=pair)
            out = P8.empty(N, cout, h, w, dt)
            run = lambda: ops.conv_fwd(x, wp, cout, k, k, k // 2, k // 2, dt, blk_c=b, tiles_per_cta=T, out=out, bias=bias,
                                       act=ACT_RELU, ctas_per_sm=cps, cta_pair=pair, debug_flags=DBG, batch_tiles=bt)
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
            tf = 2.0 * N * h * w * k * k * cin * cout / ms / 1e9
            print(f"{name}: T={T} cps={'pair' if pair else ('bt' if bt else (cps or 2))} blk={b}: {ms * 1e3:8.1f} us ({ms * 1e3 / N:8.1f} us/frame)  {tf:7.1f} TFLOP/s  maxdiff vs first {diff:.3g}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{name}: T={T} cps={cps}: EXC {e}", flush=True)
