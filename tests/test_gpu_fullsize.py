"""Full-size checks at the BASELINE.json configurations (C4 full-HD inference, C3-sized patch training): the oracle on
one full frame / one full batch, plus size-independent properties (batch invariance, run-to-run determinism, uint8
pipeline consistency, gradient linearity in the loss weight)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
H, W = 1080, 1920


def psnr(a, b, peak):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 200.0 if mse == 0 else 10 * math.log10(peak * peak / mse)


@pytest.fixture(scope="module")
def trained_sd():
    z = np.load(os.path.join(GOLD, "gen_c3_trained.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="module")
def gen(trained_sd):
    from pbt_b200.generator import GeneratorJ
    g = GeneratorJ(input_channels=3, use_bias=True)
    g.load_state_dict(trained_sd, strict=True)
    return g.cuda().eval()


def synthetic_frames(n, seed=1234):
    """the bench's C4 input recipe: low-frequency noise + fine detail, uint8 HWC"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    low = torch.rand((n, 3, 68, 120), generator=g, device="cuda") * 255.0
    up = torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)
    det = torch.rand((n, 3, H, W), generator=g, device="cuda") * 16.0 - 8.0
    return (up + det).clamp_(0, 255).round_().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def test_full_hd_frame_matches_oracle(gen, trained_sd):
    """one whole 1920x1080 frame through every full-size code path (persistent CTAs, CTA pair, normalise- and
    upsample-on-load) against the fp32 oracle on the host cores"""
    from oracle import generator_oracle as go
    u8 = synthetic_frames(1)
    x = ((u8.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5
    with torch.no_grad():
        y = gen(x).cpu()
    sd = {k: v.float() for k, v in trained_sd.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = go.generator_forward(sd, x.cpu())
    err = (y - ref).abs().max().item()
    p = psnr(y, ref, 2.0)
    print(f"1080p frame vs oracle: max_abs={err:.5f} psnr={p:.1f} dB")
    assert err <= 2e-2 and p >= 40.0, (err, p)


def test_full_hd_batch_invariance_determinism_and_u8_pipeline(gen):
    from pbt_b200.inference import FrameStylizer
    u8 = synthetic_frames(3, seed=77)
    sty = FrameStylizer(gen)
    sty.frames_per_pass = 2
    a = sty.stylize_device(u8).clone()             # passes of 2 + 1 frames
    b = sty.stylize_device(u8).clone()
    assert torch.equal(a, b), "two identical runs differ"
    sty.frames_per_pass = 1
    c = sty.stylize_device(u8)
    assert torch.equal(a, c), "a frame's result depends on which frames share its pass (InstanceNorm must be per frame)"
    assert int(a.min()) >= 0 and int(a.max()) <= 255 and a.float().std() > 1.0
    # the uint8 pipeline equals: normalise -> module forward -> clamp, (x+1)*127.5, round (reference generator.py:643-647)
    x = ((u8[:1].permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5
    with torch.no_grad():
        y = gen(x)
    q = ((y.clamp(-1, 1) + 1) * 127.5).clamp(0, 255).round().to(torch.uint8).permute(0, 2, 3, 1)
    diff = (q.int() - a[:1].int()).abs()
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 1e-3


def test_c3_sized_training_step_matches_oracle_and_is_linear_in_the_loss_weight(trained_sd):
    """batch 80 x 80x80 (the C3 shape, Cin 3 so that the reference-trained weights apply): one-step gradients against the
    fp32 oracle, and gradients of 2*loss against 2*gradients of loss"""
    from oracle import generator_oracle as go
    from pbt_b200.generator import GeneratorJ
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    frame, yframe = torch.from_numpy(vec["frame"])[0], torch.from_numpy(vec["y_frame"])[0]
    rng = np.random.RandomState(5)
    xs, ts = [], []
    for _ in range(80):
        y0, x0 = rng.randint(0, frame.shape[1] - 80), rng.randint(0, frame.shape[2] - 80)
        xs.append(frame[:, y0:y0 + 80, x0:x0 + 80])
        ts.append(yframe[:, y0:y0 + 80, x0:x0 + 80].flip(2))
    x, t = torch.stack(xs).contiguous(), torch.stack(ts).contiguous()

    def native_grads(weight):
        g = GeneratorJ(input_channels=3, use_bias=True)
        g.load_state_dict(trained_sd, strict=True)
        g = g.cuda().train()
        loss = torch.nn.functional.l1_loss(g(x.cuda()), t.cuda()) * weight
        loss.backward()
        torch.cuda.synchronize()
        return float(loss), {k: p.grad.detach().cpu() for k, p in g.named_parameters()}

    l1, g1 = native_grads(4.0)
    l2, g2 = native_grads(8.0)
    assert abs(l2 - 2 * l1) <= 1e-5 * abs(l2)
    for k in g1:
        peak = float(g1[k].abs().max())
        if peak == 0.0:
            continue
        # power-of-two loss weights only shift the fp16 gradient scale; atomics reorder fp32 sums
        assert float((g2[k] - 2 * g1[k]).abs().max()) <= 2e-3 * 2 * peak, k
    sd = {k: v.float() for k, v in trained_sd.items()}
    torch.set_num_threads(os.cpu_count() or 1)
    _, rloss, rg = go.loss_and_grads(sd, x, t)
    assert abs(l1 - float(rloss)) <= 2e-3 * abs(float(rloss))
    worst = 200.0
    for k, ref in rg.items():
        peak = float(ref.abs().max())
        if peak < 1e-12 or float(g1[k].abs().max()) == 0.0:
            continue
        worst = min(worst, psnr(g1[k], ref, peak))
    print(f"C3-sized step vs oracle: loss {l1:.5f} / {float(rloss):.5f}, worst gradient PSNR {worst:.1f} dB")
    assert worst >= 40.0, worst


def test_fp16_gradient_overflow_backs_off_and_skips_the_step(trained_sd):
    """an absurd gradient-scale target overflows fp16 in the sweep: the fused optimiser must skip those steps, the engine
    must shrink its scale on the device, and training must recover without a single non-finite weight"""
    from pbt_b200.generator import GeneratorJ
    from pbt_b200.optim import FusedClipAdam
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    x, t = torch.from_numpy(vec["x"]).cuda(), torch.from_numpy(vec["target"]).cuda()
    g = GeneratorJ(input_channels=3, use_bias=True)
    g.load_state_dict(trained_sd, strict=True)
    g = g.cuda().train()
    opt = FusedClipAdam(g.parameters(), lr=4e-4, weight_decay=1e-5, max_grad_norm=0.5)
    g(x[:1])
    g._engine.grad_scale_target = 1.0e9
    w0 = g.conv11[0].weight.detach().clone()
    losses = []
    for _ in range(12):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.l1_loss(g(x), t) * 4.0
        loss.backward()
        opt.step()
        losses.append(float(loss))
    adj = g._engine.grad_scale_adjust().cpu()
    assert opt.skipped_steps >= 1 and float(adj[2]) >= 1 and float(adj[0]) < 1.0, (opt.skipped_steps, adj)
    assert float(opt.state[g.conv11[0].weight]["step"]) == 12 - opt.skipped_steps
    assert all(torch.isfinite(p).all() for p in g.parameters())
    assert not torch.equal(w0, g.conv11[0].weight.detach())       # it recovered and trained
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0]


def test_4k_frame_matches_oracle_and_is_batch_invariant():
    """C5 of BASELINE.json: 3840x2160, Cin 5.  One frame against the fp32 oracle on the host cores (about a minute), plus
    determinism / batch invariance - 4x the units of full HD, every conv layer runs persistent CTAs here."""
    from oracle import generator_oracle as go
    from pbt_b200.generator import GeneratorJ
    z = np.load(os.path.join(GOLD, "gen_cin5_trained.npz"))      # reference-trained Cin-5 weights (oracle/make_golden_guides.py)
    sd = {k: torch.from_numpy(z[k]) for k in z.files}
    g = GeneratorJ(input_channels=5, use_bias=True)
    g.load_state_dict(sd, strict=True)
    g = g.cuda().eval()
    gen = torch.Generator(device="cuda").manual_seed(3)
    low = torch.rand((2, 5, 135, 240), generator=gen, device="cuda") * 2 - 1
    x = torch.nn.functional.interpolate(low, size=(2160, 3840), mode="bilinear", align_corners=False)
    x = (x + (torch.rand(x.shape, generator=gen, device="cuda") - 0.5) * 0.1).clamp_(-1, 1)
    with torch.no_grad():
        y2 = g(x)                      # two frames in one pass
        y0 = g(x[:1])
        y0b = g(x[:1])
    assert torch.equal(y0, y0b)
    assert torch.equal(y2[:1], y0), "a frame's result depends on its batch neighbours"
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = go.generator_forward(sd, x[:1].cpu())
    err, p = (y0.cpu() - ref).abs().max().item(), psnr(y0.cpu(), ref, 2.0)
    print(f"4K frame vs oracle: max_abs={err:.5f} psnr={p:.1f} dB")
    assert err <= 2e-2 and p >= 40.0, (err, p)
