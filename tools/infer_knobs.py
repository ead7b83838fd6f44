"""inference-time kernel configuration experiments: one 4-frame 1080p pass (the bench's unit) under each switch setting.
   python tools/infer_knobs.py [H W CIN]"""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbt_b200.generator import GeneratorJ  # noqa: E402
from pbt_b200.inference import FrameStylizer  # noqa: E402

H, W, CIN = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1080, 1920, 3)
z = np.load(os.path.join(ROOT, "tests", "golden", "gen_c3_trained.npz" if CIN == 3 else f"gen_cin{CIN}_trained.npz"))
g = GeneratorJ(input_channels=CIN, use_bias=True)
g.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
g = g.cuda().eval()
sty = FrameStylizer(g)
n = sty.pass_size(H, W)
gen = torch.Generator(device="cuda").manual_seed(1)
frames = torch.randint(0, 256, (4 * n, H, W, CIN), generator=gen, device="cuda", dtype=torch.uint8)
out = torch.empty((4 * n, H, W, 3), dtype=torch.uint8, device="cuda")
base = None
for pr, pu, ps in itertools.product((False, True), repeat=3):
    eng = g._engine
    eng.pair_res, eng.pair_up, eng.pair_smooth = pr, pu, ps
    try:
        for _ in range(2):
            sty.stylize_device(frames, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 6
        for _ in range(reps):
            sty.stylize_device(frames, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps / 4
        if base is None:
            base = out.clone()
        d = (out.int() - base.int()).abs()
        print(f"pair res={int(pr)} up={int(pu)} smooth={int(ps)}: {ms:.3f} ms per {n}-frame pass = {n * 1e3 / ms:.1f} frames/s; "
              f"max u8 diff vs default {int(d.max())}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"pair res={int(pr)} up={int(pu)} smooth={int(ps)}: FAILED {e}", flush=True)
