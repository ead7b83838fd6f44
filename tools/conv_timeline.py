"""per-CTA phase timeline of the conv kernel (debug_buf instrumentation): where does a CTA's lifetime go?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200 import ops  # noqa: E402
from pbt_b200._native import FP16, P8  # noqa: E402

dt = FP16
cases = [("up1 192->128 3x3 T2 UPSAMPLE-ON-LOAD", 192, 128, 3, 1080, 1920, 2, 32, True, 0, False, True),
         ("conv11 176->64 7x7 T3 PAIR", 176, 64, 7, 1080, 1920, 3, 32, False, 0, True),
         ("conv11 176->64 7x7 T2 blk16 PAIR cps4", 176, 64, 7, 1080, 1920, 2, 16, False, 4, True),
         ("conv11 176->64 7x7 T2 blk16 cps4", 176, 64, 7, 1080, 1920, 2, 16, False, 4, False),
         ("smooth 64->64 3x3 T2 blk16 cps4", 64, 64, 3, 1080, 1920, 2, 16, False, 4),
         ("initial 16->32 7x7 T3 cps4", 16, 32, 7, 1080, 1920, 3, 16, True, 4),
         ("res 128->128 3x3 @270x480 T2", 128, 128, 3, 270, 480, 2, 32, False),
         ("smooth 64->64 3x3 @1080x1920 T3", 64, 64, 3, 1080, 1920, 3, 32, False),
         ("initial 16->32 7x7 T3 +stats", 16, 32, 7, 1080, 1920, 3, 16, True),
         ("conv11 176->64 7x7 T3", 176, 64, 7, 1080, 1920, 3, 32, False),
         ("up1 192->128 3x3 T2", 192, 128, 3, 1080, 1920, 2, 32, False)]
FLAGS = int(sys.argv[1]) if len(sys.argv) > 1 else 0
NIMG = 1
if len(sys.argv) > 2 and sys.argv[2] == "train":     # C3 training shapes: 80 patches
    NIMG = 80
    cases = [("res 128->128 3x3 @20x20 x80 T2 BATCH TILES +stats", 128, 128, 3, 20, 20, 2, 32, True, 0, False, False, True),
             ("res 128->128 3x3 @20x20 x80 T1 cps4 +stats", 128, 128, 3, 20, 20, 1, 32, True, 4),
             ("res 128->128 3x3 @20x20 x80 T1 cps2 +stats", 128, 128, 3, 20, 20, 1, 32, True, 0),
             ("up2 256->128 3x3 @40x40 x80 T2", 256, 128, 3, 40, 40, 2, 32, True, 0),
             ("smooth 64->64 3x3 @80x80 x80 T2 blk16 cps4", 64, 64, 3, 80, 80, 2, 16, False, 4)]
for name, cin, cout, k, h, w, T, blk, stats, *rest in cases:
    cps = rest[0] if rest else 0
    pair = rest[1] if len(rest) > 1 else False
    up = rest[2] if len(rest) > 2 else False
    bt = rest[3] if len(rest) > 3 else False
    x = P8.empty(NIMG, cin, h // 2 if up else h, w // 2 if up else w, dt)
    x.t.normal_()
    wp = ops.pack_conv_weight(torch.randn((cout, cin, k, k), device="cuda") * 0.05, cin, blk, dt, pair=pair)
    out = P8.empty(NIMG, cout, h, w, dt)
    tiles = ops.conv_num_tiles(h, w, 1 if bt else T)
    part = torch.empty((NIMG, tiles, 2, cout), device="cuda") if stats else None
    n_cta = tiles * ((NIMG + T - 1) // T if bt else NIMG)
    dbg = torch.zeros((n_cta, 8), dtype=torch.int64, device="cuda")
    for _ in range(2):
        ops.conv_fwd(x, wp, cout, k, k, k // 2, k // 2, dt, blk_c=blk, tiles_per_cta=T, out=out, stats_partial=part, debug_buf=dbg, debug_flags=FLAGS | 128, upsample2x=up,
                     ctas_per_sm=cps, cta_pair=pair, batch_tiles=bt)
    torch.cuda.synchronize()
    d = dbg.cpu().double()
    t0 = d[:, 0:1]
    rel = d[:, 1:7] - t0
    med = rel.median(0).values
    names = ["setup done", "first A landed", "MMA issue done", "acc ready (epi start)", "epilogue done", "exit"]
    print(f"== {name}: {n_cta} CTAs; median cycles since CTA start:")
    for nm, v in zip(names, med.tolist()):
        print(f"     {nm:24s} {v:10.0f}")
    # per-SM occupancy of time: CTAs per SM and total span
    sm = d[:, 7].long()
    span = []
    for s in sm.unique().tolist()[:8]:
        m = sm == s
        span.append((int(m.sum()), float(d[m, 6].max() - d[m, 0].min())))
    print("     per-SM (n_ctas, span cycles) samples:", span)
    # average number of co-resident CTAs per SM = sum of CTA lifetimes / busy span (block-launch gaps show up here)
    conc = []
    for s_ in sm.unique().tolist():
        m = sm == s_
        conc.append(float((d[m, 6] - d[m, 0]).sum() / (d[m, 6].max() - d[m, 0].min())))
    print(f"     mean co-resident CTAs per SM: {sum(conc) / len(conc):.2f}; median CTA lifetime {float((d[:, 6] - d[:, 0]).median()):.0f} cycles")
