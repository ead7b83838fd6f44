"""two 4-frame 1080p generator passes (for `ncu --set full` of the conv11 launch of the second pass):
   ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel --launch-skip 41 --launch-count 1 ...
(22 conv launches per pass: initial, 2 strided, 14 residual, 2 decoder, conv11 = the 20th, 2 smoothers)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbt_b200.generator import GeneratorJ  # noqa: E402
from pbt_b200.inference import FrameStylizer  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "gen_c3_trained.npz"))
g = GeneratorJ(input_channels=3, use_bias=True)
g.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
sty = FrameStylizer(g.cuda().eval())
gen = torch.Generator(device="cuda").manual_seed(1)
frames = torch.randint(0, 256, (4, 1080, 1920, 3), generator=gen, device="cuda", dtype=torch.uint8)
for _ in range(2):
    out = sty.stylize_device(frames)
torch.cuda.synchronize()
print("checksum", int(out.sum()))
