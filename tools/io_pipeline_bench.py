"""directory inference from PNG files: sequential loop vs the decode | GPU | encode pipeline of the root generator.py"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import generator as driver  # noqa: E402
from pbt_b200.config import compose  # noqa: E402
from pbt_b200.generator import GeneratorJ  # noqa: E402

N, H, W = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (48, 1080, 1920)))
tmp = tempfile.mkdtemp(prefix="pbt_io_")
try:
    rng = np.random.RandomState(0)
    for sub in ("input", "mask"):
        os.makedirs(os.path.join(tmp, sub))
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(N):
        base = rng.randint(0, 256, (H // 16, W // 16, 3)).astype(np.uint8)
        img = np.asarray(Image.fromarray(base).resize((W, H), Image.BILINEAR))
        Image.fromarray(img).save(os.path.join(tmp, "input", f"{i:03d}.png"))
        m = ((((yy - H / 2) / (H * 0.4)) ** 2 + ((xx - W / 2) / (W * 0.4)) ** 2) <= 1).astype(np.uint8) * 255
        Image.fromarray(m, mode="L").save(os.path.join(tmp, "mask", f"{i:03d}.png"))
    torch.manual_seed(0)
    g = GeneratorJ(input_channels=3, use_bias=True)
    ckpt = os.path.join(tmp, "last.ckpt")
    torch.save({"state_dict": {"generator." + k: v for k, v in g.state_dict().items()}}, ckpt)
    for workers in (0, 8, 16):
        out = os.path.join(tmp, f"out_{workers}")
        cfg = compose(os.path.join(ROOT, "config"), "inference",
                      [f"paths.checkpoint={ckpt}", f"paths.input_dir={tmp}/input", f"paths.mask_dir={tmp}/mask",
                       f"paths.output_dir={out}", "paths.additional_channels={}", f"inference.io_workers={workers}"])
        inf = driver.StyleTransferInference(cfg)
        inf.process_image(os.path.join(tmp, "input", "000.png"), os.path.join(tmp, "mask", "000.png"), os.path.join(out, "warm.png"))
        os.remove(os.path.join(out, "warm.png"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        inf.process_directory()
        dt = time.perf_counter() - t0
        n_out = len(os.listdir(out))
        print(f"io_workers={workers:2d}: {n_out} frames {H}x{W} in {dt:.2f} s -> {n_out / dt:.1f} frames/s (PNG in, PNG out)", flush=True)
    a = np.asarray(Image.open(os.path.join(tmp, "out_0", "005.png")))
    b = np.asarray(Image.open(os.path.join(tmp, "out_8", "005.png")))
    print("pipelined output identical to sequential:", bool(np.array_equal(a, b)))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
