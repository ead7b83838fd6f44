"""GPU bring-up battery: runs every native op against its float64 restatement, one group per
subprocess (a trapped kernel poisons its CUDA context), and prints a compact report.
usage: python tools/bringup.py [group ...]      (no args = all groups)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def g_conv_basic():
    import gpu_checks as gc
    from pbt_b200._native import BF16, FP16
    cases = [
        dict(name="1x1 c16->16", cin=16, cout=16),
        dict(name="1x1 c32->16 blk32", cin=32, cout=16, blk_c=32),
        dict(name="1x1 c64->16 blk32 (2 blocks)", cin=64, cout=16, blk_c=32),
        dict(name="1x1 c48->16 blk32 (partial block)", cin=48, cout=16, blk_c=32),
        dict(name="1x1 c16->64", cin=16, cout=64),
        dict(name="1x1 c16->128", cin=16, cout=128),
        dict(name="1x1 c16->160", cin=16, cout=160),
        dict(name="1x1 fp16", cin=16, cout=16, dt=FP16),
        dict(name="3x3 c16->16", cin=16, cout=16, kh=3, kw=3, pad_t=1, pad_l=1),
        dict(name="7x7 c16->32", cin=16, cout=32, kh=7, kw=7, pad_t=3, pad_l=3),
        dict(name="2x2 pad(1,1) c32->16", cin=32, cout=16, kh=2, kw=2, pad_t=1, pad_l=1, blk_c=32),
        dict(name="2x2 pad(0,0) c32->16", cin=32, cout=16, kh=2, kw=2, pad_t=0, pad_l=0, blk_c=32),
    ]
    for c in cases:
        name = c.pop("name")
        ok, err, msg = gc.check_conv(**c)
        print(f"[conv_basic] {name}: {'PASS' if ok else 'FAIL'} err={err:.4g} {msg}", flush=True)
        if not ok and name == "1x1 c32->16 blk32":
            for fl in (1, 2, 3):
                ok2, err2, msg2 = gc.check_conv(debug_flags=fl, **c)
                print(f"    debug_flags={fl}: {'PASS' if ok2 else 'FAIL'} err={err2:.4g}", flush=True)


def g_conv_tiles():
    import gpu_checks as gc
    from pbt_b200._native import ACT_LEAKY, ACT_RELU, BF16, FP16
    cases = [
        dict(name="T2 3x3 32x32", cin=32, cout=32, h=32, w=32, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32),
        dict(name="T3 3x3 32x48", cin=32, cout=32, h=32, w=48, kh=3, kw=3, pad_t=1, pad_l=1, T=3, blk_c=32),
        dict(name="T3 7x7 48x72 n2", n=2, cin=32, cout=64, h=48, w=72, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=32),
        dict(name="partial tiles 20x20 n3", n=3, cin=128, cout=128, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32),
        dict(name="8x8 n5 c128 blk64", n=5, cin=128, cout=128, h=8, w=8, kh=3, kw=3, pad_t=1, pad_l=1, T=1, blk_c=64),
        dict(name="views in/out", n=2, cin=32, cout=32, h=16, w=16, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32,
             in_off=16, in_extra=8, out_off=8, out_extra=16),
        dict(name="c176->64 7x7 (conv11 shape)", cin=176, cout=64, h=32, w=32, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=32),
        dict(name="c256->128 3x3 (up2 shape)", cin=256, cout=128, h=16, w=16, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32),
        dict(name="bias+relu", cin=32, cout=32, h=16, w=16, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True, act=ACT_RELU),
        dict(name="bias+leaky+affine", cin=32, cout=32, h=16, w=16, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True,
             act=ACT_LEAKY, affine=True, integer=False),
        dict(name="mask+addend+out32", cin=32, cout=32, h=16, w=24, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, mask=True,
             addend=True, out32=True),
        dict(name="out32 only", cin=32, cout=32, h=16, w=24, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, out32=True,
             store16=False),
        dict(name="stats", n=2, cin=32, cout=64, h=24, w=40, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, stats=True),
        dict(name="stats partial tiles", n=2, cin=32, cout=128, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32,
             stats=True, integer=False),
        dict(name="head", n=2, cin=64, cout=64, h=16, w=24, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True,
             act=ACT_RELU, head=True, integer=False),
        dict(name="real-valued bf16 c128", cin=128, cout=128, h=32, w=32, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32,
             integer=False),
        dict(name="real-valued fp16 c128", cin=128, cout=128, h=32, w=32, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32,
             integer=False, dt=FP16),
    ]
    for c in cases:
        name = c.pop("name")
        ok, err, msg = gc.check_conv(**c)
        print(f"[conv_tiles] {name}: {'PASS' if ok else 'FAIL'} err={err:.4g} {msg}", flush=True)


def g_conv_up():
    import gpu_checks as gc
    from pbt_b200._native import BF16, FP16
    for name, kw in [("64->32 20x28 T2", dict()), ("192->128 16x16 T2", dict(cin=192, cout=128, h=16, w=16)),
                     ("256->128 10x12 T2 n3", dict(n=3, cin=256, cout=128, h=10, w=12)),
                     ("64->64 24x24 T3 bf16", dict(cin=64, cout=64, h=24, w=24, T=3, dt=BF16)),
                     ("32->16 5x7 T1 (partial tiles)", dict(cin=32, cout=16, h=5, w=7, T=1))]:
        try:
            ok, err, msg = gc.check_conv_upsample(**kw)
            print(f"[conv_up] {name}: {'PASS' if ok else 'FAIL'} {msg}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[conv_up] {name}: EXC {type(e).__name__}: {e}", flush=True)


def g_conv_nrm():
    import gpu_checks as gc
    for name, kw in [("64+32->32", dict()), ("128+48->64", dict(cpre=128, cin=48, cout=64, h=40, w=50)),
                     ("256+0->256 blk64", dict(cpre=256, cin=0, cout=256, h=20, w=20, blk_c=64)),
                     ("64+0->16 T3 bf16 leaky", dict(cpre=64, cin=0, cout=16, h=9, w=70, T=3, dt=0, act="leaky")),
                     ("16+16->16 T1 blk16 none", dict(n=3, cpre=16, cin=16, cout=16, h=5, w=7, T=1, blk_c=16, act="none"))]:
        try:
            ok, err, msg = gc.check_conv_norm_on_load(**kw)
            print(f"[conv_nrm] {name}: {'PASS' if ok else 'FAIL'} {msg}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[conv_nrm] {name}: EXC {type(e).__name__}: {e}", flush=True)


def g_wgrad():
    import gpu_checks as gc
    from pbt_b200._native import FP16
    cases = [
        dict(name="1x1 c16->16", cin=16, cout=16),
        dict(name="1x1 c128->16", cin=128, cout=16),
        dict(name="1x1 c16->128", cin=16, cout=128),
        dict(name="3x3 c16->16", cin=16, cout=16, kh=3, kw=3, pad_t=1, pad_l=1),
        dict(name="3x3 c128->128 32x32 n2", n=2, cin=128, cout=128, h=32, w=32, kh=3, kw=3, pad_t=1, pad_l=1),
        dict(name="7x7 c176->64 32x32", cin=176, cout=64, h=32, w=32, kh=7, kw=7, pad_t=3, pad_l=3),
        dict(name="2x2 c256->128 pad(1,1) 20x20 n3", n=3, cin=256, cout=128, h=20, w=20, kh=2, kw=2, pad_t=1, pad_l=1),
        dict(name="3x3 c192->128 fp16 real", cin=192, cout=128, h=24, w=24, kh=3, kw=3, pad_t=1, pad_l=1, dt=FP16,
             integer=False, inv_scale=0.25),
    ]
    for c in cases:
        name = c.pop("name")
        ok, err, msg = gc.check_wgrad(**c)
        print(f"[wgrad] {name}: {'PASS' if ok else 'FAIL'} err={err:.4g} {msg}", flush=True)
        if not ok and name == "1x1 c16->16":
            for fl in (1, 2, 3):
                ok2, err2, msg2 = gc.check_wgrad(debug_flags=fl, **c)
                print(f"    debug_flags={fl}: {'PASS' if ok2 else 'FAIL'} err={err2:.4g}", flush=True)


def g_elementwise():
    import gpu_checks as gc
    from pbt_b200._native import BF16, FP16
    for name, fn, kw in [
        ("layout bf16", gc.check_layout_roundtrip, dict(dt=BF16)),
        ("layout fp16", gc.check_layout_roundtrip, dict(dt=FP16)),
        ("u8", gc.check_u8, {}),
        ("upsample bf16", gc.check_upsample, dict(dt=BF16)),
        ("upsample fp16", gc.check_upsample, dict(dt=FP16)),
        ("norm_bwd inst fp16", gc.check_norm_bwd, dict(dt=FP16)),
        ("norm_bwd inst s2d fp16", gc.check_norm_bwd, dict(dt=FP16, s2d=True)),
        ("norm_bwd batch fp16", gc.check_norm_bwd, dict(dt=FP16, batch_mode=True)),
        ("norm_bwd inst bf16", gc.check_norm_bwd, dict(dt=BF16)),
        ("head_bwd fp16", gc.check_head_bwd, dict(dt=FP16)),
        ("head_bwd bf16", gc.check_head_bwd, dict(dt=BF16)),
        ("grad_scale", gc.check_grad_scale, {}),
        ("gather P32", gc.check_gather, dict(patch=32)),
        ("gather P80", gc.check_gather, dict(patch=80, n_patches=9)),
        ("gather P7 (odd)", gc.check_gather, dict(patch=7, n_patches=9)),
    ]:
        try:
            ok, err, msg = fn(**kw)
            print(f"[elementwise] {name}: {'PASS' if ok else 'FAIL'} err={err:.4g} {msg}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[elementwise] {name}: EXC {type(e).__name__}: {e}", flush=True)


def g_norm():
    import gpu_checks as gc
    from pbt_b200._native import BF16, FP16
    for name, kw in [("inst bf16", dict(dt=BF16)), ("inst fp16", dict(dt=FP16)), ("batch fp16", dict(dt=FP16, batch_mode=True))]:
        try:
            ok, err, msg = gc.check_norm(**kw)
            print(f"[norm] {name}: {'PASS' if ok else 'FAIL'} err={err:.4g} {msg}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[norm] {name}: EXC {type(e).__name__}: {e}", flush=True)


def g_conv_perf():
    """first performance look: the big layer shapes of one 1080p frame"""
    import torch
    from pbt_b200 import ops
    from pbt_b200._native import BF16, P8
    dt = BF16
    H, W = 1080, 1920
    shapes = [
        ("res 128->128 3x3 @270x480", 128, 128, 3, H // 4, W // 4, 2, 32),
        ("up2 256->128 3x3 @540x960", 256, 128, 3, H // 2, W // 2, 2, 32),
        ("up1 192->128 3x3 @1080x1920", 192, 128, 3, H, W, 2, 32),
        ("up1 T3", 192, 128, 3, H, W, 3, 32),
        ("up1 T2 blk64", 192, 128, 3, H, W, 2, 64),
        ("conv11 176->64 7x7 @1080x1920 T3", 176, 64, 7, H, W, 3, 32),
        ("conv11 T2", 176, 64, 7, H, W, 2, 32),
        ("conv11 T3 blk16", 176, 64, 7, H, W, 3, 16),
        ("smooth 64->64 3x3 @1080x1920", 64, 64, 3, H, W, 3, 32),
        ("initial 16->32 7x7", 16, 32, 7, H, W, 3, 16),
    ]
    for name, cin, cout, k, h, w, T, blk in shapes:
        x = P8.empty(1, cin, h, w, dt)
        x.t.normal_()
        wt = torch.randn((cout, cin, k, k), device="cuda") * 0.05
        wp = ops.pack_conv_weight(wt, cin, blk, dt)
        out = P8.empty(1, cout, h, w, dt)
        for _ in range(2):
            ops.conv_fwd(x, wp, cout, k, k, k // 2, k // 2, dt, blk_c=blk, tiles_per_cta=T, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            ops.conv_fwd(x, wp, cout, k, k, k // 2, k // 2, dt, blk_c=blk, tiles_per_cta=T, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * h * w * cin * cout * k * k
        print(f"[conv_perf] {name}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (padded-K flops)", flush=True)


GROUPS = {"conv_nrm": g_conv_nrm, "conv_up": g_conv_up, "conv_basic": g_conv_basic, "conv_tiles": g_conv_tiles, "wgrad": g_wgrad, "elementwise": g_elementwise,
          "norm": g_norm, "conv_perf": g_conv_perf}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        GROUPS[sys.argv[2]]()
        sys.exit(0)
    groups = sys.argv[1:] or list(GROUPS)
    for gname in groups:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", gname], timeout=300,
                               capture_output=True, text=True)
            out = r.stdout + ("\n[stderr tail]\n" + r.stderr[-3000:] if r.returncode != 0 else "")
            print(out, flush=True)
            print(f"== group {gname}: exit {r.returncode} in {time.time() - t0:.1f}s", flush=True)
        except subprocess.TimeoutExpired as e:
            print((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), flush=True)
            print(f"== group {gname}: TIMEOUT", flush=True)
