"""CPU-side tests: C-ABI surface, host logic (sampler bookkeeping, weight packing, config), drop-in contracts."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from pbt_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "pbt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pbt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pbt.h but not exported"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    L = _native.lib()
    assert L.pbt_abi_version() == 2
    assert L.pbt_error_string(0) == b"ok" and b"argument" in L.pbt_error_string(-1)
    assert L.pbt_conv_num_tiles(1080, 1920, 3) == 68 * 80
    # no GPU here: a compute call must fail loudly, never fall back
    if not torch.cuda.is_available():
        d = _native.ConvDesc()
        assert L.pbt_conv_fwd(ctypes.byref(d), None) != 0


def test_ostree_equals_list_pop():
    from pbt_b200.sampler import _OsTree
    rnd = random.Random(5)
    for n in (1, 2, 7, 64, 1000):
        tree, ref = _OsTree(n), list(range(n))
        for _ in range(2):                      # second pass exercises reset()
            while ref:
                k = rnd.randrange(len(ref))
                assert len(tree) == len(ref)
                assert tree.take(k) == ref.pop(k)
            with pytest.raises(IndexError):
                tree.take(0)
            tree.reset()
            ref = list(range(n))


def test_pack_conv_weight_layout():
    from pbt_b200 import ops
    w = torch.randn(32, 20, 3, 3)
    cin_pad, blk = 48, 32
    flat = ops.pack_conv_weight(w, cin_pad, blk, 1).float()
    off = 0
    wq = w.half().float()
    for c0 in range(0, cin_pad, blk):
        kc = min(blk, cin_pad - c0)
        blkt = flat[off:off + 9 * kc * 32].reshape(9, kc // 8, 32, 8)
        off += 9 * kc * 32
        for tap in range(9):
            for k8 in range(kc // 8):
                for k in range(8):
                    ci = c0 + k8 * 8 + k
                    exp = wq[:, ci, tap // 3, tap % 3] if ci < 20 else torch.zeros(32)
                    assert torch.equal(blkt[tap, k8, :, k], exp)
    assert off == flat.numel()


def test_s2d_weight_is_the_stride2_conv():
    from pbt_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(2, 8, 12, 16)
    w = torch.randn(5, 8, 3, 3)
    ref = F.conv2d(x, w, stride=2, padding=1)
    s2d = torch.cat([x[:, :, py::2, px::2] for py in range(2) for px in range(2)], 1)
    got = F.conv2d(F.pad(s2d, (1, 0, 1, 0)), ops.s2d_weight(w))
    assert torch.allclose(ref, got, atol=1e-5)
    assert torch.equal(ops.s2d_weight_grad(ops.s2d_weight(w), 8), w)


def test_dgrad_weight_is_the_input_gradient():
    from pbt_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(1, 6, 9, 11, requires_grad=True)
    w = torch.randn(4, 6, 3, 3)
    gy = torch.randn(1, 4, 9, 11)
    F.conv2d(x, w, padding=1).backward(gy)
    got = F.conv2d(gy, ops.dgrad_weight(w), padding=1)
    assert torch.allclose(x.grad, got, atol=1e-5)


def test_generator_state_dict_matches_reference_layout():
    from pbt_b200.generator import GeneratorJ
    z = np.load(os.path.join(ROOT, "tests", "golden", "gen_c3_trained.npz"))
    torch.manual_seed(0)
    g = GeneratorJ(input_channels=3, use_bias=True)
    sd = g.state_dict()
    assert list(sd.keys()) == list(z.files)
    for k in z.files:
        assert tuple(sd[k].shape) == z[k].shape, k
    g.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
    assert sum(p.numel() for p in g.parameters()) == 3265027
    assert sum(p.numel() for p in GeneratorJ(input_channels=9, use_bias=True).parameters()) == 3293251
    no_bias = GeneratorJ(input_channels=3, use_bias=False).state_dict()
    assert "initial_conv.0.bias" not in no_bias and "output.0.bias" in no_bias


def test_generator_has_no_cpu_path_and_validates_input():
    from pbt_b200.generator import GeneratorJ
    g = GeneratorJ(input_channels=3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g(torch.zeros(1, 3, 32, 32))
    gg = GeneratorJ(input_channels=3, norm_layer="batch_norm")
    gg._check_supported()                            # both norm options of the reference constructor are built
    assert "initial_conv.1.running_var" in gg.state_dict() and "resnet_blocks.0.block.5.weight" in gg.state_dict()
    plain = GeneratorJ(input_channels=3, norm_layer="none")           # any other string: no norm layers (reference :83-87)
    plain._check_supported()
    assert "resnet_blocks.0.block.3.weight" in plain.state_dict() and "initial_conv.1.weight" not in plain.state_dict()
    bare = GeneratorJ(input_channels=3, append_smoothers=False)
    bare._check_supported()
    assert not any(k.startswith("smoothers.") for k in bare.state_dict())
    with pytest.raises(NotImplementedError):
        GeneratorJ(input_channels=3, filters=[24, 64, 128, 128, 128, 64])._check_supported()


def test_config_compose_and_overrides():
    from pbt_b200.config import compose
    cfg = compose(os.path.join(ROOT, "config"), "config", ["training.batch_size=40", "data.patch_size=32", "+training.max_steps=7"])
    assert cfg.training.batch_size == 40 and cfg.data.patch_size == 32 and cfg.training.max_steps == 7
    assert cfg.model.generator.args.input_channels == "auto"
    assert cfg.model.generator.args.filters == [32, 64, 128, 128, 128, 64]
    assert cfg.optimizer.generator.lr == 0.0004 and cfg.training.gradient_clip_val == 0.5
    assert cfg.hydra.run.dir.startswith("outputs/20")
    inf = compose(os.path.join(ROOT, "config"), "inference")
    assert inf.data.dir_pre == inf.paths.input_dir and inf.data.patch_size == 80


def test_style_transfer_model_auto_channels():
    import lightning_model as lm
    from pbt_b200.config import compose
    cfg = compose(os.path.join(ROOT, "config"), "config")
    m = lm.StyleTransferModel(cfg.model.generator, None, cfg.training, cfg.optimizer, cfg.data, None)
    assert m.generator.input_channels == 6          # RGB + point_vector (depth 3)
    (opt,) = m.configure_optimizers()
    assert opt.defaults["lr"] == 0.0004 and opt.defaults["weight_decay"] == 1e-5 and opt.defaults["betas"] == (0.9, 0.999)


def test_shard_range_partitions_exactly():
    from pbt_b200.parallel import shard_range
    for n in (0, 1, 7, 500, 2000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 4, 4)


def test_frames_per_pass_follows_the_frame_size():
    from pbt_b200.inference import FrameStylizer
    sty = FrameStylizer.__new__(FrameStylizer)           # host logic only: no engine, no device
    assert sty.pass_size(540, 960) == 8 and sty.pass_size(1080, 1920) == 4 and sty.pass_size(2160, 3840) == 2
    assert sty.pass_size(64, 96) == 8 and sty.pass_size(4320, 7680) == 1
    sty.frames_per_pass = 3                              # explicit setting wins
    assert sty.pass_size(1080, 1920) == 3


def test_bench_reference_arm_prints_the_contract_line():
    # the driver runs `bench.py --impl reference` beside the native arm and computes the ratio itself
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["metric"] == "960x540 stylized frames/s" and line["unit"] == "frames/s" and line["value"] > 0
    # whole frames through the unmodified reference module when baseline/_ref exists (build() copies it), else the port
    ref_copy = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "src", "models", "generator.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_copy else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["frame"] == [960, 540, 6] and line["config"]["frames_per_step"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_grad_bucket_groups_follow_the_backward_sweep():
    """flat gradient bucket: tail | decoder | one group per residual block (last first) | encoder; 16-byte aligned slices"""
    import torch
    from pbt_b200.generator import GeneratorJ
    from pbt_b200.parallel import GradBucket
    g = GeneratorJ(input_channels=9, use_bias=True)
    named = list(g.named_parameters())
    b = GradBucket(named)
    assert len(b.group_bounds) == 2 + 7 + 1
    assert b.group_index("output.0.weight") == 0 and b.group_index("conv11.0.bias") == 0
    assert b.group_index("upsample2.1.weight") == 1
    assert b.group_index("resnet_blocks.6.block.1.weight") == 2 and b.group_index("resnet_blocks.0.block.4.bias") == 8
    assert b.group_index("initial_conv.0.weight") == 9 and b.group_index("downsample2.0.weight") == 9
    assert all(lo % 4 == 0 for lo, _ in b.slices.values())
    spans = sorted(b.slices.values())
    assert all(a[1] <= c[0] for a, c in zip(spans, spans[1:]))                   # no overlap
    assert all(b.views[n].shape == p.shape and b.views[n].data_ptr() == b.flat[b.slices[n][0]:].data_ptr() for n, p in named)
    assert not b.aliased_by_param_grads()
    named[3][1].grad = b.views[named[3][0]]
    assert b.aliased_by_param_grads()


def test_index_loader_gives_every_rank_the_same_number_of_steps():
    """DistributedSampler semantics (pad by wrapping, strided shard): unequal step counts would deadlock the all-reduce"""
    from lightning_model import _IndexLoader

    class _DS:
        def __len__(self):
            return 161

        def sample_batch(self, idx):
            return list(idx)

    seen = []
    for rank in range(2):
        import torch
        torch.manual_seed(4)                      # stands in for the seed broadcast (no process group in this test)
        ld = _IndexLoader(_DS(), 80, rank, 2)
        batches = list(ld)
        assert len(batches) == len(ld) == 2 and sum(len(b) for b in batches) == 81
        seen += [i for b in batches for i in b]
    assert set(seen) == set(range(161)) and len(seen) == 162       # one wrapped-around index, as DistributedSampler pads


def test_non_fused_adam_path_never_writes_non_finite_weights():
    """optimizer.*.fused=false uses torch.optim.Adam, which has no skip-on-overflow: `_clip_and_step` replaces non-finite
    gradients (fp16 overflow in a backward sweep) by zeros so that the weights stay finite"""
    import torch
    from lightning_model import StyleTransferModel
    tcfg = {"batch_size": 2, "reconstruction_weight": 4.0, "adversarial_weight": 0.5, "use_image_loss": True,
            "reconstruction_criterion": "L1Loss", "adversarial_criterion": "MSELoss", "use_gradient_clipping": True,
            "gradient_clip_val": 0.5, "cuda_graph": False, "use_adversarial_loss": False}
    adam = {"lr": 4e-4, "betas": [0.9, 0.999], "weight_decay": 1e-5, "fused": False}
    m = StyleTransferModel({"args": {"input_channels": 3, "use_bias": True}}, None, tcfg, {"generator": dict(adam)},
                           {"additional_channels": {}})
    (opt,) = m.configure_optimizers()
    assert isinstance(opt, torch.optim.Adam)
    w0 = [p.detach().clone() for p in m.generator.parameters()]
    for i, p in enumerate(m.generator.parameters()):
        p.grad = torch.full_like(p, float("inf") if i == 3 else 1e-3)
    m._clip_and_step(opt, m.generator)
    assert all(torch.isfinite(p).all() for p in m.generator.parameters())
    for i, p in enumerate(m.generator.parameters()):
        p.grad = torch.full_like(p, 1e-3)
    m._clip_and_step(opt, m.generator)
    assert all(torch.isfinite(p).all() for p in m.generator.parameters())
    assert any(not torch.equal(a, p.detach()) for a, p in zip(w0, m.generator.parameters()))


def test_perceptual_plan_follows_the_reference_tap_semantics():
    """host logic of the native perceptual loss (pbt_b200/perceptual.py): which tensors are tapped / rectified, and which
    configurations are handed back to the reference expression"""
    import torch
    from pbt_b200 import perceptual
    from src.models.perception import PerceptualVGG19
    C, R, M = torch.nn.Conv2d, torch.nn.ReLU, torch.nn.MaxPool2d
    stack = lambda: torch.nn.Sequential(C(3, 64, 3, padding=1), R(True), C(64, 64, 3, padding=1), R(True), M(2, 2),  # noqa: E731
                                        C(64, 128, 3, padding=1), R(True))
    mod = PerceptualVGG19.from_features(stack(), [5, 0, 3], use_normalization=False)
    assert mod.feature_layers == [0, 3, 5] and not any(p.requires_grad for p in mod.parameters())
    nodes = perceptual.plan(mod)
    # conv 0 is tapped as a view and rectified by the in-place ReLU behind it; conv 5's ReLU (index 6) is never executed
    assert [(n.kind, n.index, n.relu, n.taps) for n in nodes] == [("conv", 0, True, 1), ("conv", 2, True, 1), ("pool", 4, False, 0),
                                                                   ("conv", 5, False, 1)]
    both = perceptual.plan(PerceptualVGG19.from_features(stack(), [0, 1], use_normalization=False))
    assert [(n.relu, n.taps) for n in both] == [(True, 2)]          # index 0 and index 1 are the same (rectified) tensor, twice
    x = torch.zeros(2, 3, 32, 32)
    assert mod.native_unsupported(x) == "not on a CUDA device"
    y = torch.rand(2, 3, 16, 16) * 2 - 1
    assert float(mod.feature_mse(y, y)) == 0.0                        # the reference expression on the CPU
    wide = torch.nn.Sequential(C(3, 512, 3, padding=1))
    assert "wider than 256" in perceptual.plan(PerceptualVGG19.from_features(wide, [0], use_normalization=False))
    bn = torch.nn.Sequential(C(3, 64, 3, padding=1), torch.nn.BatchNorm2d(64), R(True))
    assert "no native kernel" in perceptual.plan(PerceptualVGG19.from_features(bn, [2], use_normalization=False))
    k5 = torch.nn.Sequential(C(3, 64, 5, padding=2))
    assert "3x3" in perceptual.plan(PerceptualVGG19.from_features(k5, [0], use_normalization=False))
    tr = PerceptualVGG19.from_features(stack(), [0], use_normalization=False, requires_grad=True)
    assert "trainable" in perceptual.plan(tr)
