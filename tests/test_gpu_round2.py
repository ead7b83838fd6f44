"""GPU tests of the round-2 additions: sampler augmentation branch and in-memory constructor, native mask erosion and
composite, the sharded frame loop (`FrameStylizer.stylize_video`), flat gradient bucket semantics, and - on a box with two
or more GPUs - NCCL data-parallel parity under torchrun (tests/dist_gpu_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _mini(s):
    return os.path.join(GOLD, "mini_dataset", s)


def test_sampler_augmentation_branch_bit_exact_vs_reference_golden():
    """augmentation_factor=2 (reference src/data/dataset.py:276-292): both draws of every item and the extra patches"""
    from pbt_b200.sampler import StyleTransferDataset
    z = np.load(os.path.join(GOLD, "sampler_aug_golden.npz"))
    ds = StyleTransferDataset(_mini("input"), _mini("output"), _mini("mask"), 32, augmentation_factor=2,
                              additional_channels={"guide": {"path": _mini("guide"), "depth": 3}})
    assert len(ds) == int(z["length"])
    log = z["log"]
    np.random.seed(321)
    for bi in range(len(log) // 8):
        rows = log[bi * 8:(bi + 1) * 8]
        batch = ds.sample_batch(rows[:, 0].tolist())
        assert torch.equal(batch["positions"], torch.from_numpy(rows[:, 1:4])), bi
        assert ds.last_patch_positions == [rows[-1, 2:4].tolist(), rows[-1, 4:6].tolist()], bi
        if bi < 2:
            for key in ("pre", "post", "channel_guide", "already", "channel_guide_aug"):
                assert np.array_equal(batch[key].cpu().numpy(), z[f"b{bi}_{key}"]), (bi, key)


def test_dataset_from_arrays_equals_directory_dataset():
    from PIL import Image
    from pbt_b200.sampler import StyleTransferDataset
    names = sorted(os.listdir(_mini("input")))
    rgb = lambda d: [np.asarray(Image.open(os.path.join(_mini(d), n)).convert("RGB")) for n in names]  # noqa: E731
    masks = [np.asarray(Image.open(os.path.join(_mini("mask"), n)).convert("L")) for n in names]
    a = StyleTransferDataset(_mini("input"), _mini("output"), _mini("mask"), 32,
                             additional_channels={"guide": {"path": _mini("guide"), "depth": 3}})
    b = StyleTransferDataset.from_arrays(rgb("input"), rgb("output"), masks, 32, additional={"guide": rgb("guide")})
    assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a.valid_indices, b.valid_indices))
    idx = list(range(0, 40, 3))
    np.random.seed(5)
    ba = a.sample_batch(idx)
    np.random.seed(5)
    bb = b.sample_batch(idx)
    for k in ("combined_input", "post", "positions"):
        assert torch.equal(ba[k], bb[k]), k


def test_mask_erode_and_composite_kernels_match_the_reference_expressions():
    from pbt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    n, h, w = 2, 61, 83
    m = (torch.rand((n, h, w), generator=g, device="cuda") > 0.03).to(torch.uint8) * 255
    m[:, 20:45, 30:70] = 255
    out = torch.empty((n, h, w), device="cuda")
    ops.mask_erode7(m.contiguous(), out)
    # reference generator.py:327-351 on the 0/1 mask
    t = (m.float() / 255.0)[:, None]
    conv = torch.nn.functional.conv2d(t, torch.ones((1, 1, 7, 7), device="cuda"), padding=3)
    ref = torch.where(conv < 49, torch.zeros_like(conv), conv) / 49
    assert torch.equal(out, ref[:, 0]) and 0 < float(out.mean()) < 1
    # composite rgb*(1-m) + y*m (:562-563) then clamp / (x+1)*127.5 / round (:643-647)
    frame = torch.randint(0, 256, (n, h, w, 5), generator=g, device="cuda", dtype=torch.uint8)
    y = torch.rand((n, 3, h, w), generator=g, device="cuda") * 2.4 - 1.2
    mask = torch.rand((n, h, w), generator=g, device="cuda").round()
    mask[0, :10] = 0.5
    res = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    ops.composite_to_u8(y, res, frame, mask)
    rgb = ((frame[..., :3].permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5
    comp = rgb * (1 - mask[:, None]) + y * mask[:, None]
    q = ((comp.clamp(-1, 1) + 1) * 127.5).clamp(0, 255).permute(0, 2, 3, 1).round().to(torch.uint8)
    assert torch.equal(res, q)
    ops.composite_to_u8(y, res)
    q = ((y.clamp(-1, 1) + 1) * 127.5).clamp(0, 255).permute(0, 2, 3, 1).round().to(torch.uint8)
    assert torch.equal(res, q)


def _small_gen(cin=3):
    from pbt_b200.generator import GeneratorJ
    z = np.load(os.path.join(GOLD, "gen_c3_trained.npz" if cin == 3 else f"gen_cin{cin}_trained.npz"))
    g = GeneratorJ(input_channels=cin, use_bias=True)
    g.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
    return g.cuda()


def test_stylize_video_shards_cover_the_video_exactly_once():
    """every (rank, world) share written by stylize_video is the single-process result of those frames, the shares tile
    the video, device- and host-resident videos agree (reference frame loop generator.py:674-705, sharded)"""
    from pbt_b200.inference import FrameStylizer
    g = torch.Generator(device="cuda").manual_seed(3)
    video = torch.randint(0, 256, (11, 64, 96, 3), generator=g, device="cuda", dtype=torch.uint8)
    sty = FrameStylizer(_small_gen())
    sty.frames_per_pass = 2
    whole = sty.stylize_device(video).clone()
    for world in (1, 2, 3, 4):
        out = torch.zeros_like(whole)
        covered = torch.zeros(11, dtype=torch.int32)
        for rank in range(world):
            lo, hi = sty.stylize_video(video, out, rank, world)
            covered[lo:hi] += 1
        assert covered.eq(1).all() and torch.equal(out, whole), world
    hin = video.cpu().pin_memory()
    hout = torch.zeros((11, 64, 96, 3), dtype=torch.uint8).pin_memory()
    for rank in range(3):
        sty.stylize_video(hin, hout, rank, 3)
    torch.cuda.synchronize()
    assert torch.equal(hout, whole.cpu())
    mask = (torch.rand((11, 64, 96), generator=g, device="cuda") > 0.5).float()
    comp = sty.stylize_device(video, masks=mask)
    assert torch.equal(comp[mask.bool()], whole[mask.bool()]) and torch.equal(comp[~mask.bool()], video[~mask.bool()])


def test_gradient_bucket_aliases_param_grads_and_survives_accumulation():
    """p.grad are views of one flat fp32 bucket (no per-tensor copies; stable addresses for the fused optimiser and the
    all-reduce); a second backward without zero_grad must ACCUMULATE like autograd does for the reference module"""
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    x, t = torch.from_numpy(vec["x"][:8]).cuda(), torch.from_numpy(vec["target"][:8]).cuda()
    g = _small_gen().train()

    def sweep():
        (torch.nn.functional.l1_loss(g(x), t) * 4.0).backward()

    sweep()
    flat = g._engine.grad_bucket().flat
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
    assert all(lo <= p.grad.data_ptr() < hi for p in g.parameters())
    ptrs = [p.grad.data_ptr() for p in g.parameters()]
    first = [p.grad.clone() for p in g.parameters()]
    g.zero_grad(set_to_none=True)
    sweep()
    assert ptrs == [p.grad.data_ptr() for p in g.parameters()]
    sweep()                                            # no zero_grad: gradients add up
    for p, f in zip(g.parameters(), first):
        peak = float(f.abs().max())
        assert float((p.grad - 2 * f).abs().max()) <= 4e-3 * peak + 1e-12     # wgrad atomics reorder fp32 sums run to run


def test_backward_of_a_stale_forward_raises():
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    x = torch.from_numpy(vec["x"][:4]).cuda()
    g = _small_gen().train()
    y1 = g(x)
    y2 = g(x * 0.5)
    with pytest.raises(RuntimeError, match="overwritten"):
        y1.sum().backward()
    y2.sum().backward()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_nccl_data_parallel_parity_under_torchrun():
    """two ranks: broadcast of the initial weights, NCCL-reduced gradient == mean of rank gradients, graph-replayed steps
    keep the replicas bit-identical, the sharded inference covers a video exactly (tests/dist_gpu_check.py)"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DIST_GPU_CHECK_OK" in r.stdout


@pytest.mark.parametrize("cfg", [dict(num_filters=12, n_layers=2, n=6, hw=80), dict(num_filters=12, n_layers=2, n=160, hw=80),
                                 dict(num_filters=8, n_layers=3, n=5, hw=64, use_bias=False), dict(num_filters=16, n_layers=1, n=3, hw=40)],
                         ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_native_critic_matches_the_tensor_library_expression(cfg):
    """DiscriminatorN_IN on the native kernels (pbt_b200/critic.py) against the same module through torch fp32 ops (the
    reference's expression, src/models/discriminator.py:105-149): logits, parameter gradients of the critic loss, and the
    gradient w.r.t. the input patches of the generator's adversarial loss (lightning_model.py:277-283,294-319)"""
    from src.models.discriminator import DiscriminatorN_IN
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(cfg)
    n, hw = cfg.pop("n"), cfg.pop("hw")
    torch.manual_seed(3)
    d = DiscriminatorN_IN(input_channels=3, **cfg).cuda().train()
    with torch.no_grad():                       # away from the N(0, 0.02) init: trained-like magnitudes
        for p in d.parameters():
            p.mul_(3.0).add_(torch.randn_like(p) * 0.01)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((n, 3, hw, hw), generator=g, device="cuda") * 2 - 1
    x[: n // 2] = torch.nn.functional.avg_pool2d(x[: n // 2], 5, 1, 2)         # smooth "real" half, noisy "fake" half
    lab = torch.cat([torch.ones(n // 2), torch.zeros(n - n // 2)]).cuda().view(n, 1, 1, 1)

    def run(native, with_input):
        d.native = native
        d.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(with_input)
        if with_input:
            for p in d.parameters():
                p.requires_grad_(False)
        try:
            logits, _ = d(xi)
            loss = torch.nn.functional.mse_loss(logits, lab.expand_as(logits))
            loss.backward()
        finally:
            for p in d.parameters():
                p.requires_grad_(True)
        return logits.detach(), float(loss), [None if p.grad is None else p.grad.detach().clone() for p in d.parameters()], xi.grad

    ref_l, ref_loss, ref_g, _ = run(False, False)
    nat_l, nat_loss, nat_g, _ = run(True, False)
    assert d._engine is not None, d._native_reason
    assert nat_l.shape == ref_l.shape
    err = float((nat_l - ref_l).abs().max())
    assert err <= 2e-2 * max(1.0, float(ref_l.abs().max())) and abs(nat_loss - ref_loss) <= 1e-2 * abs(ref_loss), (err, nat_loss, ref_loss)
    for (name, _), a, b in zip(d.named_parameters(), nat_g, ref_g):
        if name.endswith(".bias") and name.startswith(("intermediate.", "pre_output.")):
            # a bias in front of an InstanceNorm has a mathematically zero gradient: exact zeros here, round-off noise there
            wpeak = float(ref_g[[k for k, _ in d.named_parameters()].index(name.replace(".bias", ".weight"))].abs().max())
            assert float(a.abs().max()) == 0.0 and float(b.abs().max()) <= 1e-3 * wpeak, name
            continue
        cos = float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
        rel = float((a - b).norm() / b.norm())
        assert cos > 0.999 and rel < 3e-2, (name, cos, rel)
    _, _, _, gx_ref = run(False, True)
    _, _, pg, gx_nat = run(True, True)
    assert all(v is None for v in pg)                      # frozen critic: no parameter gradients are produced
    cos = float(torch.nn.functional.cosine_similarity(gx_nat.flatten(), gx_ref.flatten(), dim=0))
    rel = float((gx_nat - gx_ref).norm() / gx_ref.norm())
    print(f"{cfg} n={n}: logits err {err:.2e}, loss {nat_loss:.6f}/{ref_loss:.6f}, dL/dx cosine {cos:.5f} rel {rel:.4f}")
    assert cos > 0.999 and rel < 3e-2, (cos, rel)
    d.native = True


def test_sampler_at_full_hd_keyframes_matches_the_oracle_cut():
    """the C3 sampler at its real size (1080p keyframes, two guide sources, patch 80, batch 80): draws follow the oracle's
    bookkeeping (numpy RNG + k-th-unused selection) and every gathered patch equals the oracle's `_cut_patch` restatement on
    the resident images, including centres near the frame border"""
    from oracle import sampler_oracle as so
    from pbt_b200.sampler import StyleTransferDataset
    rng = np.random.RandomState(3)
    H, W, K = 1080, 1920, 2
    rgb = lambda: [rng.randint(0, 256, (H, W, 3)).astype(np.uint8) for _ in range(K)]  # noqa: E731
    pre, post, g1, g2 = rgb(), rgb(), rgb(), rgb()
    masks = []
    for i in range(K):
        m = np.zeros((H, W), np.uint8)
        m[:60, :90] = 255                      # touches the top-left corner: clamped, zero-padded patches
        m[H - 50:, W - 70:] = 200              # bottom-right corner
        m[400:700, 800:1300] = 255
        masks.append(m)
    ds = StyleTransferDataset.from_arrays(pre, post, masks, 80, additional={"gauss": g1, "flow": g2})
    # the oracle's draw bookkeeping on the same valid lists
    left = [list(range(len(v))) for v in ds._valid_np]
    idx = rng.randint(0, len(ds), 160).tolist()
    np.random.seed(11)
    batch_a = ds.sample_batch(idx[:80])
    batch_b = ds.sample_batch(idx[80:])
    np.random.seed(11)
    exp = []
    for i in idx:
        img = i % K
        if not left[img]:
            left[img] = list(range(len(ds._valid_np[img])))
        c = np.random.randint(0, len(left[img]))
        y, x = ds._valid_np[img][left[img].pop(c)]
        exp.append((img, int(y), int(x)))
    got = torch.cat([batch_a["positions"], batch_b["positions"]]).tolist()
    assert got == [list(e) for e in exp]
    assert any(y < 40 or x < 40 or y > H - 40 or x > W - 40 for _, y, x in exp), "no border patch in the draw"
    res = {k: [t.cpu().numpy() for t in v] for k, v in (("pre", ds.images_pre), ("post", ds.images_post),
                                                          ("gauss", ds.additional_channel_data["gauss"]),
                                                          ("flow", ds.additional_channel_data["flow"]))}
    for b, batch in enumerate((batch_a, batch_b)):
        comb, pst = batch["combined_input"].cpu().numpy(), batch["post"].cpu().numpy()
        for j in range(0, 80, 7):
            img, y, x = exp[b * 80 + j]
            ref = np.concatenate([so.cut_patch(res[k][img], y, x, 80) for k in ("pre", "gauss", "flow")], 0)
            assert np.array_equal(comb[j], ref) and np.array_equal(pst[j], so.cut_patch(res["post"][img], y, x, 80)), (b, j)


def test_fused_finalize_switch_gives_identical_results():
    """`_Engine.fuse_finalize` (InstanceNorm scale / shift computed inside the consuming norm_apply launch) is the same
    arithmetic: outputs and gradients must not change"""
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    x, t = torch.from_numpy(vec["x"][:8]).cuda(), torch.from_numpy(vec["target"][:8]).cuda()
    res = []
    for fused in (False, True):
        g = _small_gen().train()
        y0 = g(x[:1])                      # builds the engine
        g._engine.fuse_finalize = fused
        g.zero_grad(set_to_none=True)
        y = g(x)
        (torch.nn.functional.l1_loss(y, t) * 4.0).backward()
        res.append((y.detach().clone(), [p.grad.detach().clone() for p in g.parameters()]))
    assert float((res[0][0] - res[1][0]).abs().max()) <= 1e-5          # (double sums in a different association order)
    for a, b in zip(res[0][1], res[1][1]):
        peak = float(a.abs().max())
        assert float((a - b).abs().max()) <= 4e-3 * peak + 1e-12      # wgrad atomics reorder fp32 sums run to run


def test_forty_training_steps_track_the_oracle():
    """config C1 shape (batch 40 x 32x32, Cin 3): forty G-only steps (L1*4, clip 0.5, Adam 4e-4 / wd 1e-5) on two alternating real
    patch batches through the graph-replayed native step against the fp32 oracle's trajectory on the host cores"""
    from oracle import generator_oracle as go
    from pbt_b200.graphs import GraphedGeneratorStep
    from pbt_b200.optim import FusedClipAdam
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    z = np.load(os.path.join(GOLD, "gen_c3_trained.npz"))
    sd = {k: torch.from_numpy(z[k]).clone() for k in z.files}
    x, t = torch.from_numpy(vec["x"]), torch.from_numpy(vec["target"])
    batches = [(x, t), (x.flip(0).flip(3).contiguous(), t.flip(0).flip(3).contiguous())]
    torch.set_num_threads(os.cpu_count() or 1)
    st = go.AdamState([k for k, v in sd.items() if v.is_floating_point() and "running_" not in k])
    ref = [float(go.g_only_train_step(sd, st, *batches[i % 2])) for i in range(40)]
    g = _small_gen().train()
    opt = FusedClipAdam(g.parameters(), lr=4e-4, betas=(0.9, 0.999), weight_decay=1e-5, max_grad_norm=0.5)
    step = GraphedGeneratorStep(g, opt, tuple(x.shape), clip=0.5)
    got = [float(step(batches[i % 2][0].cuda(), batches[i % 2][1].cuda())) for i in range(40)]
    worst = max(abs(a - b) / b for a, b in zip(got, ref))
    print(f"native {got[0]:.4f} -> {got[-1]:.4f}, oracle {ref[0]:.4f} -> {ref[-1]:.4f}, worst relative gap {worst:.4f}, skipped {opt.skipped_steps}")
    assert got[-1] < 0.8 * got[0] and worst < 0.05 and opt.skipped_steps == 0
    # the trained weights stay close to the oracle's (same trajectory, 16-bit operand noise only)
    for k, p in g.named_parameters():
        if k.endswith(".weight") and p.dim() == 4:
            a, b = p.detach().cpu(), sd[k]
            assert float((a - b).norm() / b.norm()) < 0.05, k
