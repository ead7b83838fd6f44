"""End-to-end GPU tests of the reference-shaped entry points: train.py-style fit (device sampler -> native
generator step, eager and CUDA-graph) -> Lightning-style checkpoint -> generator.py-style directory inference."""
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINI = os.path.join(ROOT, "tests", "golden", "mini_dataset")


def _cfg(tmp, **over):
    from pbt_b200.config import compose
    ov = [f"data.dir_pre={MINI}/input", f"data.dir_post={MINI}/output", f"data.dir_mask={MINI}/mask", "data.patch_size=32",
          f"data.additional_channels.point_vector.path={MINI}/guide", "training.batch_size=8", f"training.output_dir={tmp}",
          "+training.max_steps=6", "training.max_epochs=1", "training.log_every_n_steps=2"]
    ov += [f"{k}={v}" for k, v in over.items()]
    return compose(os.path.join(ROOT, "config"), "config", ov)


@pytest.mark.parametrize("graph", [True, False])
def test_fit_checkpoint_and_directory_inference(tmp_path, graph):
    sys.path.insert(0, ROOT)
    import generator as infer_driver
    import lightning_model as lm
    from pbt_b200.config import compose
    from pbt_b200.trainer import Trainer
    cfg = _cfg(str(tmp_path), **{"+training.cuda_graph": str(graph).lower()})
    torch.manual_seed(0)
    np.random.seed(0)
    model = lm.StyleTransferModel(cfg.model.generator, cfg.model.discriminator, cfg.training, cfg.optimizer, cfg.data,
                                  cfg.model.perception_loss)
    assert model.generator.input_channels == 6
    w0 = model.generator.conv11[0].weight.detach().clone()
    tr = Trainer(max_epochs=1, max_steps=6, output_dir=str(tmp_path), log_every_n_steps=2)
    tr.fit(model)
    assert tr.global_step == 6
    assert "g_total_loss" in tr.logged and np.isfinite(tr.logged["g_total_loss"])
    assert not torch.equal(w0.cuda(), model.generator.conv11[0].weight.detach())      # the optimiser moved the weights
    ckpt = os.path.join(str(tmp_path), "checkpoints", "last.ckpt")
    sd = torch.load(ckpt, map_location="cpu")["state_dict"]
    assert "generator.initial_conv.0.weight" in sd and sd["generator.initial_conv.0.weight"].shape[1] == 6
    # directory inference with the reference-shaped driver
    out_dir = os.path.join(str(tmp_path), "stylised")
    icfg = compose(os.path.join(ROOT, "config"), "inference",
                   [f"paths.checkpoint={ckpt}", f"paths.input_dir={MINI}/input", f"paths.mask_dir={MINI}/mask",
                    f"paths.output_dir={out_dir}", f"paths.additional_channels.point_vector.path={MINI}/guide"])
    infer_driver.StyleTransferInference(icfg).process_directory()
    outs = sorted(os.listdir(out_dir))
    assert outs == sorted(os.listdir(os.path.join(MINI, "input")))
    img = np.asarray(Image.open(os.path.join(out_dir, outs[0])))
    src = np.asarray(Image.open(os.path.join(MINI, "input", outs[0])))
    assert img.shape == src.shape and img.dtype == np.uint8
    # outside the (eroded) mask the frame is the input, inside it is the generator output
    m = np.asarray(Image.open(os.path.join(MINI, "mask", outs[0])).convert("L")) > 128
    assert np.array_equal(img[~m], src[~m])
    assert not np.array_equal(img[m], src[m])
    # the reference's own windowed mode (inference.tiled=true): same driver, same files, windows + Gaussian blend
    out_t = os.path.join(str(tmp_path), "stylised_tiled")
    tcfg = compose(os.path.join(ROOT, "config"), "inference",
                   [f"paths.checkpoint={ckpt}", f"paths.input_dir={MINI}/input", f"paths.mask_dir={MINI}/mask",
                    f"paths.output_dir={out_t}", f"paths.additional_channels.point_vector.path={MINI}/guide",
                    "inference.tiled=true", "data.patch_size=32"])
    infer_driver.StyleTransferInference(tcfg).process_directory()
    assert sorted(os.listdir(out_t)) == outs
    img_t = np.asarray(Image.open(os.path.join(out_t, outs[0])))
    assert img_t.shape == src.shape and np.array_equal(img_t[~m], src[~m]) and not np.array_equal(img_t[m], src[m])


def test_frame_stylizer_host_path_matches_device_path():
    from pbt_b200.generator import GeneratorJ
    from pbt_b200.inference import FrameStylizer
    torch.manual_seed(0)
    g = GeneratorJ(input_channels=3, use_bias=True).cuda()
    sty = FrameStylizer(g)
    frames = torch.randint(0, 256, (3, 64, 96, 3), dtype=torch.uint8)
    dev = sty.stylize_device(frames.cuda()).cpu()
    host_in, host_out = frames.pin_memory(), torch.empty((3, 64, 96, 3), dtype=torch.uint8).pin_memory()
    sty.stylize_host(host_in, host_out)
    torch.cuda.synchronize()
    assert torch.equal(dev, host_out)
    # against the oracle in uint8 space
    from oracle import generator_oracle as go
    sd = {k: v.detach().cpu() for k, v in g.state_dict().items()}
    x = ((frames.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5
    ref = torch.cat([go.frame_to_uint8(go.generator_forward(sd, x[i:i + 1])) for i in range(3)])
    assert (dev.int() - ref.int()).abs().max().item() <= 2
