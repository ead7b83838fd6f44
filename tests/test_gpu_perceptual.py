"""Perceptual (VGG feature) term of the generator loss on the native kernels (reference
src/models/perception.py:93-143, lightning_model.py:270-275).  The checker is the reference expression itself evaluated by the
tensor library in fp32: `PerceptualVGG19.forward` of the drop-in module is line-for-line the reference's (flattened taps that are
views of the running activation), `((fake - target) ** 2).mean()` and autograd give the value and the gradient.
The ImageNet weights are not available offline, so the stack is a seeded VGG19 prefix (torchvision's layer order and init)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _vgg_prefix(depth: int, seed: int = 5):
    from pbt_b200.perceptual import vgg19_prefix
    return vgg19_prefix(depth, seed)


def _module(layers, norm, depth=None, seed=5):
    from src.models.perception import PerceptualVGG19
    return PerceptualVGG19.from_features(_vgg_prefix(depth or max(layers) + 1, seed), layers, use_normalization=norm)


def _patches(n, p, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.nn.functional.interpolate(torch.rand(2 * n, 3, p // 4, p // 4, generator=g), size=(p, p), mode="bilinear")
    x = (low + 0.15 * torch.randn(2 * n, 3, p, p, generator=g)).clamp(0, 1) * 2 - 1
    return x[:n].contiguous(), x[n:].contiguous()


def _reference(mod, y, t):
    """the reference expression in fp32 on the tensor library (lightning_model.py:272-274)"""
    y = y.clone().requires_grad_(True)
    loss = ((mod(y)[1] - mod(t.detach())[1]) ** 2).mean()
    (gy,) = torch.autograd.grad(loss, y)
    return loss.detach(), gy


def _cmp(got, want):
    got, want = got.double().flatten().cpu(), want.double().flatten().cpu()
    rel = float((got - want).norm() / want.norm())
    cos = float(torch.dot(got, want) / (got.norm() * want.norm()))
    return rel, cos


# gradient tolerance (relative L2, cosine) per case.  The shipped taps sit <= 3 convs deep.  The last case is a single tap behind
# 8 convs, 7 ReLUs and 2 poolings on two patches: there 16-bit operands flip ReLU masks / pooling winners, and an fp32 tensor-library
# pipeline whose operands are merely ROUNDED to fp16 between layers already differs from fp32 by 3.3e-2 / 0.99946 (measured)
@pytest.mark.parametrize("layers,norm,n,p,tol", [([0, 3, 5], False, 8, 32, (1e-2, 0.9999)), ([0, 3, 5], True, 6, 80, (1e-2, 0.9999)),
                                                 ([5, 0, 3], False, 1, 16, (1e-2, 0.9999)), ([1, 4, 7, 9], True, 4, 32, (1e-2, 0.9999)),
                                                 ([2], False, 3, 24, (1e-2, 0.9999)), ([12], True, 2, 32, (6e-2, 0.998))])
def test_feature_mse_matches_the_reference_expression(layers, norm, n, p, tol):
    mod = _module(layers, norm).cuda()
    y, t = _patches(n, p, seed=n * 100 + p)
    y, t = y.cuda(), t.cuda()
    assert mod.native_unsupported(y) is None
    want_loss, want_g = _reference(mod, y, t)
    yg = y.clone().requires_grad_(True)
    loss = mod.feature_mse(yg, t)
    (loss * 6.0).backward()
    assert loss.shape == () and loss.dtype == torch.float32
    assert float(loss.detach()) == pytest.approx(float(want_loss), rel=3e-3)
    rel, cos = _cmp(yg.grad / 6.0, want_g)
    print(f"layers {layers} norm {norm}: loss {float(loss.detach()):.6f} vs {float(want_loss):.6f}; gradient rel L2 {rel:.2e}, cosine {cos:.6f}")
    assert rel < tol[0] and cos > tol[1]
    # value only (validation): same number, no gradient work
    with torch.no_grad():
        again = mod.feature_mse(y, t)
    assert float(again) == float(loss)
    # perceptual_loss of the reference API routes through the same kernels
    assert float(mod.perceptual_loss(y, t)) == float(loss)


def test_feature_mse_is_reproducible_and_zero_for_equal_inputs():
    mod = _module([0, 3, 5], False).cuda()
    y, t = _patches(8, 32, seed=3)
    y, t = y.cuda(), t.cuda()
    outs = []
    for _ in range(3):
        yg = y.clone().requires_grad_(True)
        loss = mod.feature_mse(yg, t)
        loss.backward()
        outs.append((loss.clone(), yg.grad.clone()))
    for l, g in outs[1:]:
        assert torch.equal(l, outs[0][0]) and torch.equal(g, outs[0][1])       # fixed summation order, no float atomics
    yg = y.clone().requires_grad_(True)
    loss = mod.feature_mse(yg, y.clone())
    loss.backward()
    assert float(loss) == 0.0 and float(yg.grad.abs().max()) == 0.0


def test_unsupported_configurations_take_the_reference_expression():
    y, t = _patches(2, 32, seed=9)
    y, t = y.cuda(), t.cuda()
    mod = _module([0, 3, 5], False).cuda()
    assert "multiple of 2" in mod.native_unsupported(y[:, :, :31, :31])
    odd = mod.feature_mse(y[:, :, :31, :31].contiguous(), t[:, :, :31, :31].contiguous())
    want = ((mod(y[:, :, :31, :31])[1] - mod(t[:, :, :31, :31])[1]) ** 2).mean()
    assert float(odd) == pytest.approx(float(want), rel=1e-6)
    trainable = _module([0, 3, 5], False)
    for q in trainable.parameters():
        q.requires_grad = True
    trainable = trainable.cuda()
    assert "trainable" in trainable.native_unsupported(y)
    assert mod.native_unsupported(y.cpu()) == "not on a CUDA device"
    from pbt_b200 import perceptual
    with pytest.raises(RuntimeError, match="native perceptual loss"):
        perceptual.PerceptualEngine(mod).loss_and_grad(y[:, :, :31, :31].contiguous(), t[:, :, :31, :31].contiguous(), True)
    with pytest.raises(ValueError, match="differ in shape"):
        perceptual.PerceptualEngine(mod).loss_and_grad(y, t[:1], True)


def test_maxpool_and_its_transpose_match_the_tensor_library():
    from pbt_b200 import ops
    from pbt_b200._native import FP16, P8
    torch.manual_seed(4)
    for (n, c, h, w) in ((3, 16, 8, 12), (2, 64, 20, 20), (1, 8, 7, 9)):
        x = torch.randn(n, c, h, w, device="cuda").half().float()
        x[:, :, ::3, ::2] = 0.0
        x = torch.relu(x)                                  # many ties at zero, as behind a ReLU
        x[0, 0, 0:2, 0:2] = 1.5                            # a tie between non-zero values: the first one takes the gradient
        xp = P8.from_nchw(x, FP16)
        yp = P8.empty(n, c, h // 2, w // 2, FP16)
        ops.maxpool2(xp, yp, FP16)
        xr = x.clone().requires_grad_(True)
        yr = torch.nn.functional.max_pool2d(xr, 2, 2)
        assert torch.equal(yp.to_nchw(), yr.detach())
        gy = torch.randn_like(yr).half().float()
        (gx_ref,) = torch.autograd.grad(yr, xr, gy)
        half = max(1, n - 1)                                # the gradient may cover only the leading images
        gxp = P8(torch.full((half, c // 8, h, w, 8), 7.0, dtype=torch.float16, device="cuda"))
        ops.maxpool2_bwd(xp, P8.from_nchw(gy, FP16), gxp, FP16)
        assert torch.equal(gxp.to_nchw(), gx_ref[:half])


def test_feature_mse_kernel_flags():
    from pbt_b200 import ops
    from pbt_b200._native import FP16, P8
    torch.manual_seed(6)
    n, c, h, w = 3, 24, 10, 14
    f = torch.randn(2 * n, c, h, w, device="cuda").half().float()
    g0 = torch.randn(n, c, h, w, device="cuda").half().float()
    fp = P8.from_nchw(f, FP16)
    partial, counter = torch.empty(4096, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    d = f[:n] - f[n:]
    for accumulate in (False, True):
        for relu in (False, True):
            for tap in (False, True):
                if not tap and not relu and not accumulate:
                    continue
                gp = P8.from_nchw(g0, FP16)
                loss = torch.full((), 2.0, device="cuda")
                ops.feature_mse(fp, n, FP16, g=gp, grad_mul=0.5, accumulate=accumulate, relu=relu, tap=tap, partial=partial,
                                counter=counter, loss=loss, loss_mul=0.25)
                want = (g0 if accumulate else torch.zeros_like(g0)) + (0.5 * d if tap else 0.0)
                if relu:
                    want = want * (f[:n] > 0)
                assert torch.equal(gp.to_nchw()[:, :c], want.half().float()), (accumulate, relu, tap)
                want_loss = 2.0 + (0.25 * float((d.double() ** 2).sum()) if tap else 0.0)
                assert float(loss) == pytest.approx(want_loss, rel=1e-5)
                assert int(counter) == 0
    with pytest.raises(RuntimeError, match="feature_mse"):
        ops.feature_mse(fp, n + 1, FP16, tap=True, partial=partial, counter=counter, loss=torch.zeros((), device="cuda"))
    with pytest.raises(RuntimeError, match="feature_mse"):
        ops.feature_mse(fp, n, FP16, tap=False)
    with pytest.raises(RuntimeError, match="maxpool2"):
        ops.maxpool2(fp, P8.empty(2 * n, c, h // 2, w // 2 + 1, FP16), FP16)


def test_generator_step_with_the_native_perceptual_term():
    """StyleTransferModel._generator_step with a VGG-prefix perceptual module: the whole step (native generator, native taps)
    against the fp32 oracle of the generator + the reference expression of the taps"""
    import os
    from lightning_model import StyleTransferModel
    from oracle import generator_oracle as go
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g_sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(gold, "gen_c3_trained.npz")).items()}
    vec = np.load(os.path.join(gold, "gen_c3_vectors.npz"))
    x, post = torch.from_numpy(vec["x"][:8]).contiguous(), torch.from_numpy(vec["target"][:8]).contiguous()
    train_cfg = {"batch_size": 8, "reconstruction_weight": 4.0, "adversarial_weight": 0.5, "use_image_loss": True,
                 "reconstruction_criterion": "L1Loss", "adversarial_criterion": "MSELoss", "use_gradient_clipping": True,
                 "gradient_clip_val": 0.5, "cuda_graph": False}
    adam = {"lr": 0.0004, "betas": [0.9, 0.999], "weight_decay": 0.00001}
    m = StyleTransferModel({"args": {"input_channels": 3, "use_bias": True}}, None, train_cfg, {"generator": dict(adam)},
                           {"additional_channels": {}})
    m.generator.load_state_dict(g_sd, strict=True)
    m.perception_loss_model, m.perception_loss_weight = _module([0, 3, 5], False), 6.0
    m = m.cuda().train()
    from pbt_b200._native import LAUNCHES
    before = LAUNCHES[0]
    out = m._generator_step(x.cuda(), {"post": post.cuda()})
    out["loss"].backward()
    assert getattr(m.perception_loss_model, "_native_engine", None) is not None and LAUNCHES[0] > before
    names = [k for k, v in g_sd.items() if v.is_floating_point() and "running_" not in k]
    leaves = {k: g_sd[k].clone().requires_grad_(True) for k in names}
    y = go.generator_forward({**g_sd, **leaves}, x, training=True)
    taps = _module([0, 3, 5], False)
    rec = (y - post).abs().mean() * 4.0
    per = ((taps(y)[1] - taps(post)[1]) ** 2).mean() * 6.0
    ref = dict(zip(names, torch.autograd.grad(rec + per, [leaves[k] for k in names], allow_unused=True)))
    assert float(out["g_perception_loss"]) == pytest.approx(float(per), rel=2e-2)
    assert float(out["g_total_loss"]) == pytest.approx(float(rec + per), rel=1e-2)
    for k in ("output.0.weight", "smoothers.3.weight", "conv11.0.weight", "upsample1.1.weight", "resnet_blocks.3.block.1.weight",
              "initial_conv.0.weight"):
        got, want = dict(m.generator.named_parameters())[k].grad.cpu(), ref[k]
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0))
        peak = float(want.abs().max())
        ps = 10 * np.log10(peak * peak / float(((got.double() - want.double()) ** 2).mean()))
        print(f"{k}: cosine {cos:.5f} psnr {ps:.1f} dB")
        assert cos > 0.99 and ps >= 40.0, k


def test_three_term_step_graphed_equals_eager():
    """the reference's default generator loss (L1 x 4 + VGG taps x 6 + adversarial x 0.5, config/model/default.yaml) as one
    CUDA-graph replay: same losses and same weights as the eager step, two steps in a row"""
    import os
    from lightning_model import StyleTransferModel
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g_sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(gold, "gen_c3_trained.npz")).items()}
    vec = np.load(os.path.join(gold, "gen_c3_vectors.npz"))
    x, post = torch.from_numpy(vec["x"][:8]).contiguous().cuda(), torch.from_numpy(vec["target"][:8]).contiguous().cuda()
    tcfg = {"batch_size": 8, "reconstruction_weight": 4.0, "adversarial_weight": 0.5, "use_image_loss": True,
            "reconstruction_criterion": "L1Loss", "adversarial_criterion": "MSELoss", "use_gradient_clipping": True,
            "gradient_clip_val": 0.5}
    adam = {"lr": 0.0004, "betas": [0.9, 0.999], "weight_decay": 0.00001}
    d_args = dict(input_channels=3, num_filters=12, n_layers=2, use_noise=False, noise_sigma=0.2, norm_layer="instance_norm", use_bias=True)
    runs = {}
    for graphed in (False, True):
        torch.manual_seed(2024)
        m = StyleTransferModel({"args": {"input_channels": 3, "use_bias": True}}, {"type": "DiscriminatorN_IN", "args": dict(d_args)},
                               dict(tcfg, cuda_graph=graphed), {"generator": dict(adam), "discriminator": dict(adam)},
                               {"additional_channels": {}})
        m.generator.load_state_dict(g_sd, strict=True)
        m.perception_loss_model, m.perception_loss_weight = _module([0, 3, 5], False), 6.0
        m = m.cuda().train()
        m._optimizers = m.configure_optimizers()
        hist = []
        for step in range(2):
            out = m.graphed_training_step({"combined_input": x, "post": post}, step) if graphed else m.full_step(x, post)
            hist.append({k: float(v) for k, v in out.items() if k != "loss"})
        runs[graphed] = (hist, {k: v.detach().clone() for k, v in m.generator.state_dict().items()})
        assert getattr(m.perception_loss_model, "_native_engine", None) is not None
    for a, b in zip(*[runs[g][0] for g in (False, True)]):
        assert set(a) == set(b) and "g_perception_loss" in a and a["g_perception_loss"] > 0
        for k in a:
            assert b[k] == pytest.approx(a[k], rel=2e-3, abs=1e-6), k
    for k, v in runs[False][1].items():
        if v.is_floating_point():
            assert (v - runs[True][1][k]).abs().max().item() <= 1.7e-3, k     # two Adam steps of <= 4e-4 each; sign noise on ~0 gradients
