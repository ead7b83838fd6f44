"""whole training pipeline at the C3 shape: sampler draws (host) + patch gather + graphed G-only step, through the
reference-shaped StyleTransferModel / loader (not just the generator step on a fixed batch)"""
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
import lightning_model as lm  # noqa: E402
from pbt_b200.config import compose  # noqa: E402

K, H, W = 6, 1080, 1920
tmp = tempfile.mkdtemp(prefix="pbt_train_")
try:
    rng = np.random.RandomState(0)
    yy, xx = np.mgrid[0:H, 0:W]
    for sub in ("input", "output", "mask", "gauss", "flow"):
        os.makedirs(os.path.join(tmp, sub))
    for i in range(K):
        for sub in ("input", "output", "gauss", "flow"):
            base = rng.randint(0, 256, (H // 24, W // 24, 3)).astype(np.uint8)
            Image.fromarray(base).resize((W, H), Image.BILINEAR).save(os.path.join(tmp, sub, f"{i:03d}.png"), compress_level=1)
        cy, cx = rng.randint(300, 780), rng.randint(500, 1400)
        m = ((((yy - cy) / 260.0) ** 2 + ((xx - cx) / 420.0) ** 2) <= 1).astype(np.uint8) * 255
        Image.fromarray(m, mode="L").save(os.path.join(tmp, "mask", f"{i:03d}.png"), compress_level=1)
    cfg = compose(os.path.join(ROOT, "config"), "config",
                  [f"data.dir_pre={tmp}/input", f"data.dir_post={tmp}/output", f"data.dir_mask={tmp}/mask", "data.patch_size=80",
                   f"data.additional_channels.point_vector.path={tmp}/gauss", f"+data.additional_channels.flow.path={tmp}/flow",
                   "+data.additional_channels.flow.depth=3",
                   "training.batch_size=80", f"training.output_dir={tmp}/out"])
    torch.manual_seed(0)
    np.random.seed(0)
    model = lm.StyleTransferModel(cfg.model.generator, cfg.model.discriminator, cfg.training, cfg.optimizer, cfg.data,
                                  cfg.model.perception_loss).cuda()
    model.setup("fit")
    model._optimizers = model.configure_optimizers()
    model.train()
    print("input channels", model.generator.input_channels, "dataset length", len(model.train_dataset), flush=True)
    loader = iter(model.train_dataloader())
    t_s = t_g = 0.0
    for step in range(70):
        if step == 20:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            t_s = 0.0
        a = time.perf_counter()
        batch = next(loader)
        t_s += time.perf_counter() - a
        out = model.graphed_training_step(batch, step)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"pipeline: {dt / 50 * 1e3:.2f} ms/step -> {80 * 50 / dt:.0f} patches/s (host time in the sampler: {t_s / 50 * 1e3:.2f} ms/step); "
          f"loss {float(out['loss']):.4f}", flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
