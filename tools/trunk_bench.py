"""Fused residual trunk (csrc/res_trunk.cu) against the layer-by-layer launches it replaces, training shape of config C3
(80 images, 128 channels, 20x20 maps, 7 blocks):   python tools/trunk_bench.py [N H W BLOCKS]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200 import ops  # noqa: E402
from pbt_b200._native import ACT_NONE, ACT_RELU, FP16, P8  # noqa: E402

n, h, w, nb = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (80, 20, 20, 7)))
C, dt = 128, FP16
torch.manual_seed(0)
r0 = torch.randn(n, C, h, w, device="cuda")
ws = [(torch.randn(C, C, 3, 3, device="cuda") * 0.03) for _ in range(2 * nb)]
packed = [ops.pack_conv_weight(wt, C, 32, dt) for wt in ws]
E = lambda: P8.empty(n, C, h, w, dt)  # noqa: E731
a = [P8.from_nchw(torch.relu(r0), dt)] + [E() for _ in range(nb)]
raw_a, hmid, raw_b = [E() for _ in range(nb)], [E() for _ in range(nb)], [E() for _ in range(nb)]
stats = [[(torch.empty(n, C, device="cuda"), torch.empty(n, C, device="cuda")) for _ in range(nb)] for _ in range(2)]
res0 = r0.reshape(n, C // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous()
res = [res0.clone(), res0.clone()]
last16 = E()


def fused():
    ops.res_trunk_fwd(a[:nb], raw_a, hmid, raw_b, packed[0::2], packed[1::2], stats[0], stats[1], res[0], last16, dt)


T = 1
tiles = ops.conv_num_tiles(h, w, T)
part = torch.empty(n, tiles, 2, C, device="cuda")


def layered():
    """the launches of pbt_b200/generator.py's training trunk: conv (batch tiles) -> finalize -> apply, twice per block"""
    rc, rn = res[0], res[1]
    for b in range(nb):
        for half in range(2):
            xin = a[b] if half == 0 else hmid[b]
            raw = raw_a[b] if half == 0 else raw_b[b]
            sc, sh = stats[half][b]
            ops.conv_fwd(xin, packed[2 * b + half], C, 3, 3, 1, 1, dt, blk_c=32, tiles_per_cta=2, out=raw, stats_partial=part,
                         ctas_per_sm=0, batch_tiles=True)
            ops.norm_finalize(part, n, tiles, C, h * w, sc, sh, eps=1e-5)
            if half == 0:
                ops.norm_apply(raw, dt, scale=sc, shift=sh, act=ACT_RELU, out=hmid[b])
            else:
                ops.norm_apply(raw, dt, scale=sc, shift=sh, act=ACT_NONE, residual32=rc, out32=rn, out=last16 if b == nb - 1 else None,
                               out_relu=a[b + 1])
                rc, rn = rn, rc


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


flops = 2.0 * n * h * w * C * C * 9 * 2 * nb
tf, tl = timed(fused), timed(layered)
print(f"residual trunk {n} x 128 x {h}x{w}, {nb} blocks: fused {tf:.1f} us ({flops / tf / 1e6:.0f} TFLOP/s), "
      f"layer by layer {tl:.1f} us ({6 * nb} launches, {flops / tl / 1e6:.0f} TFLOP/s)")

import ctypes  # noqa: E402
from pbt_b200._native import lib  # noqa: E402
if hasattr(lib(), "pbt_debug_trunk_cycles"):
    buf = (ctypes.c_longlong * 8)()
    lib().pbt_debug_trunk_cycles(buf)
    fused()
    torch.cuda.synchronize()
    lib().pbt_debug_trunk_cycles(buf)
    names = ["MMA warp waits for weights", "MMA warp waits for the map", "epilogue waits for accumulators (x4 warps)", "pass 1", "pass 2", "pass 3"]
    if any(buf):      # only a -DPBT_TRUNK_DBG=1 build of csrc/res_trunk.cu counts
        print("block 0 cycles per conv: " + ", ".join(f"{nm} {buf[i] / (2 * nb):.0f}" for i, nm in enumerate(names)))
