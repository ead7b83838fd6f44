"""Cost of the perceptual term at the C3 training shape (80 patches of 80x80, taps [0, 3, 5]): the native path
(pbt_b200/perceptual.py) against the reference expression on the tensor library (cuDNN fp32, and fp16 autocast), value + gradient
w.r.t. the generated patches, CUDA events.

    python tools/perceptual_step_bench.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_gpu_perceptual import _module, _patches  # noqa: E402


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from pbt_b200._native import LAUNCHES
    n, p = 80, 80
    mod = _module([0, 3, 5], False).cuda()
    y, t = _patches(n, p, seed=1)
    y, t = y.cuda(), t.cuda()

    def native():
        yg = y.detach().requires_grad_(True)
        mod.feature_mse(yg, t).backward()
        return yg.grad

    def library(autocast):
        def run():
            yg = y.detach().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.float16, enabled=autocast):
                loss = ((mod(yg)[1] - mod(t)[1]) ** 2).mean()
            loss.backward()
            return yg.grad
        return run

    g_nat, g_ref = native(), library(False)()
    rel = float((g_nat - g_ref).norm() / g_ref.norm())
    l0 = LAUNCHES[0]
    native()
    launches = LAUNCHES[0] - l0
    ms_nat = timed(native)
    ms_fp32 = timed(library(False))
    ms_amp = timed(library(True))
    flops = 2 * n * 2 * p * p * 9 * (3 * 64 + 64 * 64 + 64 * 128 / 4) + 2 * n * p * p * 9 * (3 * 64 + 64 * 64 + 64 * 128 / 4)
    print(f"perceptual term, {n} x 3 x {p}x{p}, taps [0, 3, 5]: native {ms_nat:.3f} ms ({launches} launches, "
          f"{flops / ms_nat / 1e9:.0f} TFLOP/s of conv work), tensor library fp32 {ms_fp32:.3f} ms, fp16 autocast {ms_amp:.3f} ms; "
          f"gradient rel L2 native vs fp32 {rel:.2e}")


if __name__ == "__main__":
    main()
