"""interleaved A/B of engine switches on the bench's unit (one 4-frame 1080p pass): every configuration is timed ROUNDS times
in rotation (12 passes each time), so that slow thermal / power drift hits all of them alike.
   python tools/infer_ab.py "ws_up=0,ws_res=0" "ws_up=1,ws_res=0" ..."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbt_b200.generator import GeneratorJ  # noqa: E402
from pbt_b200.inference import FrameStylizer  # noqa: E402

H, W, CIN, ROUNDS = 1080, 1920, 3, 5
z = np.load(os.path.join(ROOT, "tests", "golden", "gen_c3_trained.npz"))
g = GeneratorJ(input_channels=CIN, use_bias=True)
g.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
sty = FrameStylizer(g.cuda().eval())
gen = torch.Generator(device="cuda").manual_seed(1)
frames = torch.randint(0, 256, (12, H, W, CIN), generator=gen, device="cuda", dtype=torch.uint8)
out = torch.empty((12, H, W, 3), dtype=torch.uint8, device="cuda")
configs = [dict((kv.split("=")[0], kv.split("=")[1] == "1") for kv in a.split(",")) for a in sys.argv[1:]]
times = [[] for _ in configs]
base = None
for _ in range(3):
    sty.stylize_device(frames, out)
for r in range(ROUNDS):
    for ci, cfg in enumerate(configs):
        for k, v in cfg.items():
            assert hasattr(sty.eng, k), k
            setattr(sty.eng, k, v)
        sty.eng._ws.clear()
        sty.eng._wslots.clear()
        sty.stylize_device(frames[:4], out[:4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            sty.stylize_device(frames, out)
        e1.record()
        torch.cuda.synchronize()
        times[ci].append(e0.elapsed_time(e1) / 12)
        if base is None:
            base = out.clone()
        assert int((out.int() - base.int()).abs().max()) <= 1
for cfg, t in zip(configs, times):
    t = sorted(t)
    print(f"{cfg}: median {t[len(t) // 2]:.3f} ms per 4-frame pass ({4e3 / t[len(t) // 2]:.1f} frames/s), min {t[0]:.3f}, max {t[-1]:.3f}", flush=True)
