"""quick device timing of the native generator: full-frame inference and a patch-training step"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200.generator import GeneratorJ  # noqa: E402


def timeit(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps


def F(cin):
    return 2017664 + 9408 * cin


for dtype in ("fp16", "bf16"):
    for (n, cin, h, w) in ((1, 3, 1080, 1920), (1, 6, 960, 540), (1, 5, 2160, 3840)):
        torch.manual_seed(0)
        g = GeneratorJ(input_channels=cin, use_bias=True).cuda().eval()
        g.operand_dtype = dtype
        x = torch.rand(n, cin, h, w, device="cuda") * 2 - 1
        with torch.no_grad():
            ms, wall = timeit(lambda: g(x))
        tf = F(cin) * h * w * n / ms / 1e9
        print(f"[infer {dtype}] {n}x{cin}x{h}x{w}: {ms:.2f} ms/frame (host {wall:.2f} ms) -> {1e3 / ms * n:.1f} frames/s, {tf:.0f} TFLOP/s algorithmic",
              flush=True)
        del g, x
        torch.cuda.empty_cache()
    for (n, cin, p) in ((40, 3, 32), (80, 9, 80)):
        torch.manual_seed(0)
        g = GeneratorJ(input_channels=cin, use_bias=True).cuda().train()
        g.operand_dtype = dtype
        opt = torch.optim.Adam(g.parameters(), lr=4e-4, weight_decay=1e-5)
        x = torch.rand(n, cin, p, p, device="cuda") * 2 - 1
        t = torch.rand(n, 3, p, p, device="cuda") * 2 - 1

        def step():
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.l1_loss(g(x), t) * 4.0
            loss.backward()
            torch.nn.utils.clip_grad_norm_(g.parameters(), 0.5)
            opt.step()

        ms, wall = timeit(step)
        tf = 3 * F(cin) * p * p * n / ms / 1e9
        print(f"[train {dtype}] batch {n} x {cin}ch x {p}^2: {ms:.2f} ms/step (host {wall:.2f} ms) -> {n * 1e3 / ms:.0f} patches/s, {tf:.0f} TFLOP/s algorithmic",
              flush=True)
        del g, opt
        torch.cuda.empty_cache()
