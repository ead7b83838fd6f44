"""a few full-frame inference passes (for ncu launch lists): python tools/infer_once.py [H W CIN FRAMES]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200.generator import GeneratorJ  # noqa: E402
from pbt_b200.inference import FrameStylizer  # noqa: E402

h, w, cin, frames = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (1080, 1920, 3, 3)))
torch.manual_seed(0)
g = GeneratorJ(input_channels=cin, use_bias=True).cuda().eval()
sty = FrameStylizer(g)
x = torch.randint(0, 256, (frames, h, w, cin), dtype=torch.uint8, device="cuda")
out = sty.stylize_device(x)
torch.cuda.synchronize()
print("checksum", int(out.sum()))
