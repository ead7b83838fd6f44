"""The fused residual trunk (csrc/res_trunk.cu) against a float64 restatement of the reference's ResNetBlock chain
(src/models/generator.py:18-58) evaluated on the same 16-bit inputs and weights, every saved tensor checked."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _run(n, h, w, nb, seed, dt_name="fp16"):
    from pbt_b200 import ops
    from pbt_b200._native import BF16, FP16, P8, torch_dtype
    dt = {"fp16": FP16, "bf16": BF16}[dt_name]
    tdt = torch_dtype(dt)
    g = torch.Generator(device="cuda").manual_seed(seed)
    C = 128
    r0 = torch.randn((n, C, h, w), generator=g, device="cuda")
    ws = [(torch.randn((C, C, 3, 3), generator=g, device="cuda") * 0.03).to(tdt).float() for _ in range(2 * nb)]
    a0 = torch.relu(r0).to(tdt).float()
    E = lambda: P8.empty(n, C, h, w, dt)  # noqa: E731
    a = [P8.from_nchw(a0, dt)] + [E() for _ in range(nb - 1)]
    raw_a, hmid, raw_b = [E() for _ in range(nb)], [E() for _ in range(nb)], [E() for _ in range(nb)]
    stats = [[(torch.empty(n, C, device="cuda"), torch.empty(n, C, device="cuda")) for _ in range(nb)] for _ in range(2)]
    res = r0.reshape(n, C // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous()
    wide = P8.empty(n, 2 * C, h, w, dt, zero=True)          # last16 is a channel view of a wider tensor, like c2cat
    wide.t.fill_(3.0)
    last16 = wide.view(0, C)
    packed = [ops.pack_conv_weight(wt, C, 32, dt) for wt in ws]
    ops.res_trunk_fwd(a, raw_a, hmid, raw_b, packed[0::2], packed[1::2], stats[0], stats[1], res, last16, dt)
    torch.cuda.synchronize()
    # float64 chain with the same roundings: conv inputs and saved tensors are 16 bit, statistics are taken from the rounded raw output
    rnd = lambda t: t.to(tdt).double()  # noqa: E731
    worst = {}

    def cmp(name, got, want, tol):
        err = float((got.double() - want).abs().max()) / max(1.0, float(want.abs().max()))
        worst[name] = max(worst.get(name, 0.0), err)
        assert err <= tol, (name, err)

    tol16 = 6e-3 if dt == FP16 else 3e-2
    r = r0.double()
    x = a0.double()
    for b in range(nb):
        for half, (raw_t, wt) in enumerate(((raw_a[b], ws[2 * b]), (raw_b[b], ws[2 * b + 1]))):
            conv = F.conv2d(x, wt.double(), padding=1)
            cmp("raw", raw_t.to_nchw(), conv, tol16)
            rr = raw_t.to_nchw().double()                     # continue from the kernel's own rounded raw output
            mean, var = rr.mean((2, 3)), rr.var((2, 3), unbiased=False)
            rstd = torch.rsqrt(var + 1e-5)
            sc, sh = stats[half][b]
            cmp("scale", sc, rstd, 2e-4)
            cmp("shift", sh, -mean * rstd, 2e-4)
            y = rr * rstd[:, :, None, None] - (mean * rstd)[:, :, None, None]
            if half == 0:
                cmp("hmid", hmid[b].to_nchw(), torch.relu(y), tol16)
                x = hmid[b].to_nchw().double()
            else:
                r = r + y
                if b + 1 < nb:
                    got_r = res.permute(0, 1, 4, 2, 3).reshape(n, C, h, w)
                    cmp("a_next", a[b + 1].to_nchw(), torch.relu(r), tol16)
                    x = a[b + 1].to_nchw().double()
                else:
                    cmp("last16", last16.to_nchw(), r, tol16)
    if nb > 1:   # the fp32 residual stream holds r_{nb-1}
        r_chk = r0.double()
        # recompute r_{nb-1} from the kernel's own saved tensors
        for b in range(nb - 1):
            rr = raw_b[b].to_nchw().double()
            sc, sh = stats[1][b]
            r_chk = r_chk + rr * sc.double()[:, :, None, None] + sh.double()[:, :, None, None]
        cmp("residual32", res.permute(0, 1, 4, 2, 3).reshape(n, C, h, w), r_chk, 1e-5)
    assert bool((wide.to_nchw()[:, C:] == 3.0).all()), "last16 clobbered the neighbouring channels"
    return worst


@pytest.mark.parametrize("n,h,w,nb", [(3, 20, 20, 2), (2, 8, 8, 1), (5, 16, 16, 3), (2, 12, 40, 2), (1, 20, 20, 7), (2, 4, 4, 2),
                                      (160, 20, 20, 7), (3, 1, 1, 1), (2, 2, 6, 2), (1, 3, 100, 1)])
def test_fused_trunk_matches_the_block_chain(n, h, w, nb):
    worst = _run(n, h, w, nb, seed=n * 1000 + h * 10 + nb)
    print(f"n={n} {h}x{w} blocks={nb}: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()))


def test_fused_trunk_bf16_and_bad_arguments():
    from pbt_b200 import ops
    _run(2, 20, 20, 2, seed=77, dt_name="bf16")
    assert ops.res_trunk_supported(128, 20, 20) and ops.res_trunk_supported(128, 8, 8)
    assert not ops.res_trunk_supported(128, 24, 24) and not ops.res_trunk_supported(64, 20, 20)
    assert not ops.res_trunk_supported(128, 3, 126)        # 384 rows, but the 128-pixel pitch makes the halo rows too large for shared memory
    from pbt_b200._native import FP16, P8
    E = lambda c, h, w: P8.empty(1, c, h, w, FP16)  # noqa: E731
    st = [(torch.empty(1, 128, device="cuda"), torch.empty(1, 128, device="cuda"))]
    wp = [torch.zeros(9 * 128 * 128, dtype=torch.float16, device="cuda")]
    with pytest.raises(RuntimeError, match="res_trunk"):      # 24x24 does not fit
        ops.res_trunk_fwd([E(128, 24, 24)], [E(128, 24, 24)], [E(128, 24, 24)], [E(128, 24, 24)], wp, wp, st, st,
                          torch.zeros(1, 16, 24, 24, 8, device="cuda"), E(128, 24, 24), FP16)
    with pytest.raises(RuntimeError, match="res_trunk"):      # mismatching tensor
        ops.res_trunk_fwd([E(128, 8, 8)], [E(128, 8, 8)], [E(128, 8, 4)], [E(128, 8, 8)], wp, wp, st, st,
                          torch.zeros(1, 16, 8, 8, 8, device="cuda"), E(128, 8, 8), FP16)


def test_generator_step_is_the_same_with_and_without_the_fused_trunk():
    """whole GeneratorJ training pass (forward, loss, backward) with the trunk as one launch against the layer-by-layer launches:
    both are 16-bit-operand evaluations of the same arithmetic, so they agree far inside the tolerance against the fp32 oracle"""
    import os
    import numpy as np
    from pbt_b200.generator import GeneratorJ, _Engine
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(gold, "gen_cin9_trained.npz")).items()}
    vec = np.load(os.path.join(gold, "gen_cin9_vectors.npz"))
    x, t = torch.from_numpy(vec["p80_x"]).cuda(), torch.from_numpy(vec["p80_target"]).cuda()
    out = {}
    for fused in (True, False):
        g = GeneratorJ(input_channels=9, use_bias=True)
        g.load_state_dict(sd, strict=True)
        g = g.cuda().train()
        g._engine = _Engine(g)
        g._engine.fused_trunk = fused
        from pbt_b200._native import LAUNCHES
        l0 = LAUNCHES[0]
        y = g(x)
        fwd_launches = LAUNCHES[0] - l0
        loss = torch.nn.functional.l1_loss(y, t) * 4.0
        loss.backward()
        out[fused] = (y.detach(), float(loss.detach()), {k: p.grad.detach().clone() for k, p in g.named_parameters()}, fwd_launches)
    yf, lf, gf, nf = out[True]
    yl, ll, gl, nl = out[False]
    assert nl - nf == 6 * 7 - 1                       # 7 blocks x (2 convs + 2 finalize + 2 apply) -> one launch
    assert float((yf - yl).abs().max()) < 4e-3 and abs(lf - ll) < 2e-3 * abs(ll)
    for k in gf:
        if float(gl[k].abs().max()) < 1e-6:           # biases in front of an InstanceNorm: zero gradient
            continue
        cos = float(torch.nn.functional.cosine_similarity(gf[k].flatten().double(), gl[k].flatten().double(), dim=0))
        assert cos > 0.995, (k, cos)
