"""frames/s of FrameStylizer.stylize_device against frames_per_pass: python tools/frames_per_pass_sweep.py [H W CIN]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200.generator import GeneratorJ  # noqa: E402
from pbt_b200.inference import FrameStylizer  # noqa: E402

h, w, cin = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (960, 540, 6)))
torch.manual_seed(0)
g = GeneratorJ(input_channels=cin, use_bias=True).cuda().eval()
sty = FrameStylizer(g)
x = torch.randint(0, 256, (16, h, w, cin), dtype=torch.uint8, device="cuda")
out = torch.empty((16, h, w, 3), dtype=torch.uint8, device="cuda")
for fpp in (1, 2, 4, 8):
    sty.frames_per_pass = fpp
    for _ in range(2):
        sty.stylize_device(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        sty.stylize_device(x, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (5 * 16)
    print(f"[{h}x{w}x{cin}] frames_per_pass={fpp}: {ms:.3f} ms/frame -> {1e3 / ms:.1f} frames/s", flush=True)
