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


def test_checkpoint_carries_optimizer_state_and_fit_resumes(tmp_path):
    """checkpoints hold `optimizer_states` (Lightning's key) next to the state_dict; `Trainer.fit(model, ckpt_path=...)` restores
    weights, Adam moments / step counters and the global step, and training continues from there"""
    sys.path.insert(0, ROOT)
    import lightning_model as lm
    from pbt_b200.trainer import Trainer
    cfg = _cfg(str(tmp_path))
    torch.manual_seed(0)
    np.random.seed(0)
    build = lambda: lm.StyleTransferModel(cfg.model.generator, cfg.model.discriminator, cfg.training, cfg.optimizer, cfg.data,  # noqa: E731
                                          cfg.model.perception_loss)
    model = build()
    tr = Trainer(max_epochs=1, max_steps=4, output_dir=str(tmp_path), log_every_n_steps=2)
    tr.fit(model)
    ckpt_path = os.path.join(str(tmp_path), "checkpoints", "last.ckpt")
    ckpt = torch.load(ckpt_path, map_location="cpu")
    assert ckpt["global_step"] == 4 and len(ckpt["optimizer_states"]) == len(model.optimizers())
    st0 = ckpt["optimizer_states"][0]["state"]
    assert len(st0) == len(list(model.generator.parameters())) and float(st0[0]["step"]) == 4
    assert float(st0[0]["exp_avg"].abs().sum()) > 0
    w_ck = ckpt["state_dict"]["generator.conv11.0.weight"]
    # a fresh model resumes: same weights and moments as the checkpoint, then moves on
    model2 = build()
    tr2 = Trainer(max_epochs=2, max_steps=6, output_dir=str(tmp_path / "resumed"), log_every_n_steps=2)
    tr2.fit(model2, ckpt_path=ckpt_path)
    assert tr2.global_step == 6
    opt_g = model2.optimizers()[0]
    p0 = next(iter(model2.generator.parameters()))
    assert float(opt_g.state[p0]["step"]) == 6                      # 4 restored + 2 new steps
    assert not torch.equal(model2.generator.conv11[0].weight.detach().cpu(), w_ck)


def test_process_directory_shards_the_frame_list_by_rank(tmp_path, monkeypatch):
    """under torchrun every process stylises its contiguous range of the frame list (reference loop generator.py:674-705 is
    the one-process case); no process group is needed"""
    sys.path.insert(0, ROOT)
    import generator as infer_driver
    from pbt_b200.config import compose
    from pbt_b200.generator import GeneratorJ
    torch.manual_seed(0)
    g = GeneratorJ(input_channels=6, use_bias=True)
    ckpt = os.path.join(str(tmp_path), "g.ckpt")
    torch.save({"state_dict": {"generator." + k: v for k, v in g.state_dict().items()}}, ckpt)
    names = sorted(os.listdir(os.path.join(MINI, "input")))
    seen = []
    for rank in range(2):
        out_dir = os.path.join(str(tmp_path), f"out{rank}")
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("WORLD_SIZE", "2")
        monkeypatch.setenv("LOCAL_RANK", "0")
        icfg = compose(os.path.join(ROOT, "config"), "inference",
                       [f"paths.checkpoint={ckpt}", f"paths.input_dir={MINI}/input", f"paths.mask_dir={MINI}/mask",
                        f"paths.output_dir={out_dir}", f"paths.additional_channels.point_vector.path={MINI}/guide"])
        infer_driver.StyleTransferInference(icfg).process_directory()
        seen.append(sorted(os.listdir(out_dir)))
    assert seen[0] + seen[1] == names and len(seen[0]) == 3 and len(seen[1]) == 2      # 5 frames: ranges [0,3) and [3,5)
