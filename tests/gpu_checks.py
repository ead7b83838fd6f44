"""Shared GPU parity checks (used by the pytest -m gpu suite and by tools/bringup.py).

Every check compares one native op (called through the C-ABI) with a float64 PyTorch restatement of the
same arithmetic on identical inputs.  Integer-valued test data makes most comparisons exact.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from pbt_b200 import ops
from pbt_b200._native import ACT_LEAKY, ACT_NONE, ACT_RELU, BF16, FP16, P8, torch_dtype


def _ints(shape, lo, hi, gen, device="cuda"):
    return torch.randint(lo, hi + 1, shape, generator=gen, device=device).float()


def ref_conv(x, w, pad_t, pad_l):
    kh, kw = w.shape[2], w.shape[3]
    xp = F.pad(x.double(), (pad_l, kw - 1 - pad_l, pad_t, kh - 1 - pad_t))
    return F.conv2d(xp, w.double())


def act_ref(v, act):
    if act == ACT_RELU:
        return v.clamp_min(0)
    if act == ACT_LEAKY:
        return torch.where(v > 0, v, 0.2 * v)
    return v


def check_conv(n=1, cin=16, cout=16, h=16, w=8, kh=1, kw=1, pad_t=0, pad_l=0, T=1, blk_c=16, dt=BF16, seed=0,
               debug_flags=0, in_off=0, in_extra=0, out_off=0, out_extra=0, bias=False, act=ACT_NONE, affine=False,
               mask=False, addend=False, out32=False, stats=False, head=False, integer=True, store16=True, cps=0, pair=False, bt=False, tp=0):
    """returns (ok, max_abs_err, message); cps = ctas_per_sm configuration of the kernel"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    if integer:
        x = _ints((n, cin, h, w), -3, 3, g)
        wt = _ints((cout, cin, kh, kw), -2, 2, g)
    else:
        x = torch.randn((n, cin, h, w), generator=g, device="cuda")
        wt = torch.randn((cout, cin, kh, kw), generator=g, device="cuda") * 0.05
    tdt = torch_dtype(dt)
    x = x.to(tdt).float()
    wt = wt.to(tdt).float()
    if tp:   # tap-pair mode (first layer): only the first `tp` (<= 8) of the 16 input channels carry weights
        wt[:, tp:] = 0
    # input lives inside a wider tensor (channel-offset view)
    xfull = P8.empty(n, in_off + cin + in_extra, h, w, dt, zero=True)
    xfull.t.copy_(P8.from_nchw(torch.cat([torch.full((n, in_off, h, w), 7.0, device="cuda"), x,
                                          torch.full((n, in_extra, h, w), -5.0, device="cuda")], 1), dt).t)
    xin = xfull.view(in_off, cin)
    wp = ops.pack_conv_weight_tap_pairs(wt[:, :tp], dt) if tp else ops.pack_conv_weight(wt, cin, blk_c, dt, pair=pair)
    kwargs = {"cta_pair": pair, "batch_tiles": bt, "tap_pairs": bool(tp)}
    exp = ref_conv(x, wt, pad_t, pad_l)
    if bias:
        b = _ints((cout,), -4, 4, g)
        kwargs["bias"] = b
        exp = exp + b.double().view(1, -1, 1, 1)
    exp = act_ref(exp, act)
    kwargs["act"] = act
    if affine:
        ps, pb = _ints((cout,), 1, 3, g), _ints((cout,), -2, 2, g)
        kwargs["post_scale"], kwargs["post_shift"] = ps, pb
        exp = exp * ps.double().view(1, -1, 1, 1) + pb.double().view(1, -1, 1, 1)
    if mask:
        m = _ints((n, cout, h, w), -1, 1, g)
        kwargs["mask"] = P8.from_nchw(m, dt)
        exp = torch.where(m > 0, exp, torch.zeros_like(exp))
    if addend:
        a = _ints((n, cout, h, w), -3, 3, g)
        kwargs["addend32"] = a.reshape(n, cout // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous()
        exp = exp + a.double()
    o32 = None
    if out32:
        o32 = torch.full((n, cout // 8, h, w, 8), float("nan"), device="cuda")
        kwargs["out32"] = o32
    ofull = None
    if store16:
        ofull = P8.empty(n, out_off + cout + out_extra, h, w, dt, zero=True)
        ofull.t.fill_(9.0)
        kwargs["out"] = ofull.view(out_off, cout)
    tiles = ops.conv_num_tiles(h, w, 1 if bt else T)
    part = None
    if stats:
        part = torch.full((n, tiles, 2, cout), float("nan"), device="cuda")
        kwargs["stats_partial"] = part
    hout = None
    if head:
        hw_ = torch.randn((3, cout), generator=g, device="cuda") * 0.1
        hb = torch.randn((3,), generator=g, device="cuda") * 0.1
        hout = torch.full((n, 3, h, w), float("nan"), device="cuda")
        kwargs.update(head_w=hw_, head_b=hb, head_out=hout, head_tanh=True)
    ops.conv_fwd(xin, wp, cout, kh, kw, pad_t, pad_l, dt, blk_c=blk_c, tiles_per_cta=T, debug_flags=debug_flags, ctas_per_sm=cps,
                 **kwargs)
    torch.cuda.synchronize()
    msgs, ok, worst = [], True, 0.0
    tol16 = 0.0 if integer else (2e-2 if dt == BF16 else 3e-3)
    stored = None
    if store16:
        got = ofull.view(out_off, cout).to_nchw().double()
        stored = got
        e16 = exp.to(tdt).double() if integer else exp
        err = (got - e16).abs().max().item()
        rel = err / max(1.0, exp.abs().max().item())
        worst = max(worst, err)
        if not (rel <= tol16 if not integer else err == 0.0):
            ok = False
            bad = (got - e16).abs() > (tol16 * max(1.0, exp.abs().max().item()))
            idx = bad.nonzero()[:5].tolist()
            msgs.append(f"out16 max_abs_err={err:.4g} nbad={int(bad.sum())}/{bad.numel()} first={idx}")
        # neighbouring channels must be untouched
        full = ofull.to_nchw()
        if out_off and not bool((full[:, :out_off] == 9.0).all()):
            ok = False
            msgs.append("out16 clobbered channels below the view")
        if out_extra and not bool((full[:, out_off + cout:] == 9.0).all()):
            ok = False
            msgs.append("out16 clobbered channels above the view")
    if out32:
        got = o32.permute(0, 1, 4, 2, 3).reshape(n, cout, h, w).double()
        err = (got - exp).abs().max().item()
        worst = max(worst, err)
        if not (err <= (1e-9 if integer else 1e-3 * max(1.0, exp.abs().max().item()))):
            ok = False
            msgs.append(f"out32 max_abs_err={err:.4g}")
    if stats:
        base = stored if stored is not None else exp
        s_exp = torch.stack([base.sum((2, 3)), (base * base).sum((2, 3))], 1)  # [n,2,c]
        s_got = part.double().sum(1)
        err = ((s_got - s_exp).abs() / (1.0 + s_exp.abs())).max().item()
        if not (err < 1e-4):
            ok = False
            msgs.append(f"stats rel_err={err:.4g}")
    if head:
        base = stored if stored is not None else exp
        h_exp = torch.tanh(torch.einsum("nchw,jc->njhw", base, hw_.double()) + hb.double().view(1, 3, 1, 1))
        err = (hout.double() - h_exp).abs().max().item()
        if not (err < 1e-4):
            ok = False
            msgs.append(f"head max_abs_err={err:.4g}")
    return ok, worst, "; ".join(msgs)


def check_wgrad(n=1, cin=16, cout=16, h=16, w=8, kh=1, kw=1, pad_t=0, pad_l=0, dt=BF16, seed=0, debug_flags=0,
                integer=True, inv_scale=None):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if integer:
        x = _ints((n, cin, h, w), -2, 2, g)
        dy = _ints((n, cout, h, w), -2, 2, g)
    else:
        x = torch.randn((n, cin, h, w), generator=g, device="cuda")
        dy = torch.randn((n, cout, h, w), generator=g, device="cuda")
    tdt = torch_dtype(dt)
    x, dy = x.to(tdt).float(), dy.to(tdt).float()
    wt = torch.zeros((cout, cin, kh, kw), device="cuda", dtype=torch.float64, requires_grad=True)
    out = ref_conv(x, wt, pad_t, pad_l)
    (gw,) = torch.autograd.grad(out, wt, dy.double())
    exp = gw.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout)  # [tap][ci][co]
    dw = torch.zeros((kh * kw, cin, cout), device="cuda")
    inv = None
    if inv_scale is not None:
        inv = torch.tensor([inv_scale], device="cuda")
        exp = exp * inv_scale
    ops.conv_wgrad(P8.from_nchw(x, dt), P8.from_nchw(dy, dt), kh, kw, pad_t, pad_l, dt, dw, inv_scale=inv,
                   debug_flags=debug_flags)
    torch.cuda.synchronize()
    err = (dw.double() - exp).abs().max().item()
    scale = max(1.0, exp.abs().max().item())
    ok = err == 0.0 if integer and (n * h * w * 4 < 2 ** 24) else err / scale < 2e-3
    msg = ""
    if not ok:
        bad = ((dw.double() - exp).abs() > 1e-3 * scale)
        msg = f"max_abs_err={err:.4g} (scale {scale:.4g}) nbad={int(bad.sum())}/{bad.numel()} first={bad.nonzero()[:5].tolist()}"
    return ok, err, msg


# ----------------------------------------------------------------------------- elementwise
def check_layout_roundtrip(dt=BF16):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = _ints((2, 5, 12, 20), -8, 8, g)
    out = P8.empty(2, 16, 12, 20, dt)
    out.t.fill_(3.0)
    ops.nchw_to_p8(x, out, dt)
    ok = bool((out.to_nchw()[:, :5] == x).all()) and bool((out.to_nchw()[:, 5:] == 0).all())
    xh = x.half()
    out2 = P8.empty(2, 8, 12, 20, dt)
    ops.nchw_to_p8(xh, out2, dt)
    ok &= bool((out2.to_nchw()[:, :5] == x).all())
    back = torch.empty((2, 5, 12, 20), device="cuda")
    ops.p8_to_nchw(out, 5, back, dt, 0.5)
    ok &= bool((back == x * 0.5).all())
    x32 = torch.randn((2, 2, 12, 20, 8), generator=g, device="cuda")
    back2 = torch.empty((2, 11, 12, 20), device="cuda")
    ops.p8f_to_nchw(x32, 11, back2)
    ok &= bool((back2 == x32.permute(0, 1, 4, 2, 3).reshape(2, 16, 12, 20)[:, :11]).all())
    return ok, 0.0, ""


def check_u8(dt=FP16):
    g = torch.Generator(device="cuda").manual_seed(2)
    img = torch.randint(0, 256, (2, 10, 14, 3), generator=g, device="cuda", dtype=torch.uint8)
    out = P8.empty(2, 16, 10, 14, dt)
    ops.u8hwc_to_p8(img, out, dt)
    # reference arithmetic on the CPU (true IEEE division, as ToTensor/Normalize do in the reference's
    # dataset; torch's CUDA kernels divide by a scalar through a reciprocal multiply and differ in the last bit)
    ref = (((img.cpu().permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5).cuda()
    msgs = []
    if not (bool((out.to_nchw()[:, :3] == ref.to(torch_dtype(dt)).float()).all()) and bool((out.to_nchw()[:, 3:] == 0).all())):
        msgs.append("u8hwc_to_p8 mismatch")
    chw = torch.empty((3, 10, 14), device="cuda")
    ops.u8hwc_to_norm_chw(img[0].contiguous(), chw)
    if not bool((chw == ref[0]).all()):  # bit exact vs ToTensor+Normalize arithmetic
        msgs.append(f"u8hwc_to_norm_chw mismatch max={float((chw - ref[0]).abs().max()):.3g}")
    y = torch.randn((2, 3, 10, 14), generator=g, device="cuda") * 0.8
    u8 = torch.empty((2, 10, 14, 3), device="cuda", dtype=torch.uint8)
    ops.nchw_to_u8hwc(y, u8)
    yc = y.cpu()
    exp = ((yc.clamp(-1, 1) + 1) * 127.5).clamp(0, 255).permute(0, 2, 3, 1).round().to(torch.uint8).cuda()
    if not bool((u8 == exp).all()):
        msgs.append(f"nchw_to_u8hwc mismatch n={int((u8 != exp).sum())}")
    m = (torch.rand((37, 53), generator=g, device="cuda") > 0.97).to(torch.uint8) * 255
    d = torch.empty_like(m)
    ops.mask_dilate7(m, d)
    expd = (F.conv2d((m.float() / 255.0)[None, None], torch.ones((1, 1, 7, 7), device="cuda"), padding=3)[0, 0] != 0)
    if not bool((d.bool() == expd).all()):
        msgs.append("mask_dilate7 mismatch")
    return not msgs, 0.0, "; ".join(msgs)


def check_norm(dt=BF16, batch_mode=False):
    """conv stats -> finalize -> apply (+act, residual, s2d, relu copy) against torch instance/batch norm"""
    g = torch.Generator(device="cuda").manual_seed(3)
    n, c, h, w = 3, 32, 24, 40
    tdt = torch_dtype(dt)
    x = (torch.randn((n, c, h, w), generator=g, device="cuda") * 2 + 0.5).to(tdt).float()
    # statistics through a 1x1 identity conv so that the epilogue path is the one exercised
    eye = torch.eye(c, device="cuda").reshape(c, c, 1, 1)
    wp = ops.pack_conv_weight(eye, c, 32, dt)
    T = 2
    tiles = ops.conv_num_tiles(h, w, T)
    part = torch.empty((n, tiles, 2, c), device="cuda")
    raw = P8.empty(n, c, h, w, dt)
    ops.conv_fwd(P8.from_nchw(x, dt), wp, c, 1, 1, 0, 0, dt, blk_c=32, tiles_per_cta=T, out=raw, stats_partial=part)
    scale = torch.empty((n, c), device="cuda")
    shift = torch.empty((n, c), device="cuda")
    mean = torch.empty((n, c), device="cuda")
    rstd = torch.empty((n, c), device="cuda")
    gamma = torch.rand((c,), generator=g, device="cuda") + 0.5
    beta = torch.randn((c,), generator=g, device="cuda")
    rm, rv = torch.zeros((c,), device="cuda"), torch.ones((c,), device="cuda")
    ops.norm_finalize(part, n, tiles, c, h * w, scale, shift, batch_mode=batch_mode, gamma=gamma if batch_mode else None,
                      beta=beta if batch_mode else None, running_mean=rm if batch_mode else None,
                      running_var=rv if batch_mode else None, mean_out=mean, rstd_out=rstd)
    res = torch.randn((n, c // 8, h, w, 8), generator=g, device="cuda")
    out = P8.empty(n, c, h, w, dt)
    out_relu = P8.empty(n, c, h, w, dt)
    out_s2d = P8.empty(n, 4 * c, h // 2, w // 2, dt)
    o32 = torch.empty((n, c // 8, h, w, 8), device="cuda")
    ops.norm_apply(raw, dt, scale=scale, shift=shift, act=ACT_LEAKY, residual32=res, out=out, out_relu=out_relu, out32=o32,
                   out_s2d=out_s2d)
    torch.cuda.synchronize()
    xd = x.double()
    if batch_mode:
        bn = torch.nn.BatchNorm2d(c).cuda().double()
        bn.weight.data.copy_(gamma)
        bn.bias.data.copy_(beta)
        bn.train()
        normed = bn(xd)
        ok = torch.allclose(rm.double(), bn.running_mean, atol=1e-5) and torch.allclose(rv.double(), bn.running_var, atol=1e-4)
    else:
        normed = F.instance_norm(xd, eps=1e-5)
        ok = True
    exp = act_ref(normed, ACT_LEAKY) + res.permute(0, 1, 4, 2, 3).reshape(n, c, h, w).double()
    got32 = o32.permute(0, 1, 4, 2, 3).reshape(n, c, h, w).double()
    e32 = (got32 - exp).abs().max().item()
    ok &= e32 < 2e-4
    tol = 4e-2 if dt == BF16 else 5e-3
    e16 = (out.to_nchw().double() - exp).abs().max().item()
    ok &= e16 < tol
    ok &= bool((out_relu.to_nchw() == got32.float().to(tdt).float().clamp_min(0)).all())
    s2d = out_s2d.to_nchw().reshape(n, 2, 2, c, h // 2, w // 2)
    o = out.to_nchw()
    for py in range(2):
        for px in range(2):
            ok &= bool((s2d[:, py, px] == o[:, :, py::2, px::2]).all())
    return ok, max(e32, e16), f"e32={e32:.3g} e16={e16:.3g}"


def check_upsample(dt=BF16):
    g = torch.Generator(device="cuda").manual_seed(4)
    n, c, h, w = 2, 24, 10, 20
    tdt = torch_dtype(dt)
    x = torch.randn((n, c, h, w), generator=g, device="cuda").to(tdt).float()
    big = P8.empty(n, 40, 2 * h, 2 * w, dt, zero=True)
    ops.upsample2x(P8.from_nchw(x, dt), big.view(16, c), dt)
    exp = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    got = big.view(16, c).to_nchw()
    e = (got - exp).abs().max().item()
    # fp16 runs in packed half2 arithmetic: <= 1 ulp at the test's magnitude (|x| up to ~4 -> ulp 3.9e-3)
    ok = e < (2e-2 if dt == BF16 else 4e-3) and bool((big.to_nchw()[:, :16] == 0).all())
    # fused producer: upsample(leaky(x*scale+shift)) with per-(n,c) scale/shift
    sc = torch.rand((n, c), generator=g, device="cuda") + 0.5
    sh = torch.randn((n, c), generator=g, device="cuda")
    fused = P8.empty(n, c, 2 * h, 2 * w, dt)
    ops.upsample2x(P8.from_nchw(x, dt), fused, dt, scale=sc, shift=sh, act=ACT_LEAKY)
    expf = F.interpolate(act_ref(x * sc[:, :, None, None] + sh[:, :, None, None], ACT_LEAKY), scale_factor=2,
                         mode="bilinear", align_corners=True)
    ef = (fused.to_nchw() - expf).abs().max().item()
    ok &= ef < (4e-2 if dt == BF16 else 8e-3)
    # transpose
    gy = torch.randn((n, c, 2 * h, 2 * w), generator=g, device="cuda").to(tdt).float()
    xr = x.double().requires_grad_(True)
    F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True).backward(gy.double())
    g32 = torch.empty((n, c // 8, h, w, 8), device="cuda")
    g16 = P8.empty(n, c, h, w, dt)
    ops.upsample2x_bwd(P8.from_nchw(gy, dt), dt, gin16=g16, gin32=g32)
    torch.cuda.synchronize()
    got32 = g32.permute(0, 1, 4, 2, 3).reshape(n, c, h, w).double()
    eb = (got32 - xr.grad).abs().max().item()
    ok &= eb < (1e-4 if dt == BF16 else 6e-3)   # fp16: the horizontal taps are blended in packed half2 arithmetic
    ok &= (g16.to_nchw().double() - xr.grad).abs().max().item() < (4e-2 if dt == BF16 else 8e-3)
    return ok, max(e, eb), f"fwd={e:.3g} bwd={eb:.3g}"


def check_norm_bwd(dt=BF16, batch_mode=False, s2d=False, n=2, c=16):
    """n * c/8 >= 128 with per-image statistics exercises the fused one-launch kernel"""
    g = torch.Generator(device="cuda").manual_seed(5)
    h, w = 12, 16
    tdt = torch_dtype(dt)
    x = (torch.randn((n, c, h, w), generator=g, device="cuda") * 1.5 + 0.3).to(tdt).float()
    ga = torch.randn((n, c, h, w), generator=g, device="cuda").to(tdt).float()
    gb = torch.randn((n, c, h, w), generator=g, device="cuda").to(tdt).float()
    gc = torch.randn((n, c, h, w), generator=g, device="cuda")
    xd = x.double().requires_grad_(True)
    gamma = (torch.rand((c,), generator=g, device="cuda") + 0.5)
    if batch_mode:
        mean = xd.mean((0, 2, 3), keepdim=True)
        var = xd.var((0, 2, 3), unbiased=False, keepdim=True)
    else:
        mean = xd.mean((2, 3), keepdim=True)
        var = xd.var((2, 3), unbiased=False, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    xhat = (xd - mean) * rstd
    y = act_ref(xhat, ACT_LEAKY) if not batch_mode else xhat * gamma.double().view(1, -1, 1, 1)
    gtot = ga.double() + gb.double() + gc.double()
    y.backward(gtot)
    rs = rstd.detach().float().reshape(-1, c) if not batch_mode else rstd.detach().float().reshape(c)
    mn = mean.detach().float().reshape(-1, c) if not batch_mode else mean.detach().float().reshape(c)
    scale = rs.contiguous()
    shift = (-mn * rs).contiguous()
    kmul = (rs * gamma).contiguous() if batch_mode else scale
    sums = torch.zeros((2, c) if batch_mode else (n, 2, c), device="cuda")
    dx = P8.empty(n, c, h, w, dt)
    if s2d:
        ga_p = P8.from_nchw(
            torch.cat([ga[:, :, py::2, px::2] for py in range(2) for px in range(2)], 1), dt)
    else:
        ga_p = P8.from_nchw(ga, dt)
    gc32 = gc.reshape(n, c // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous()
    ops.norm_bwd(P8.from_nchw(x, dt), dt, scale=scale, shift=shift, per_channel=batch_mode,
                 act=ACT_NONE if batch_mode else ACT_LEAKY, ga=ga_p, ga_is_s2d=s2d, gb16=P8.from_nchw(gb, dt), gb32=gc32,
                 sums=sums, kmul=kmul, count=(n * h * w if batch_mode else h * w), batch_mode=batch_mode, dx=dx)
    torch.cuda.synchronize()
    e = (dx.to_nchw().double() - xd.grad).abs().max().item()
    ok = e < (6e-2 if dt == BF16 else 8e-3)
    return ok, e, f"err={e:.3g}"


def check_head_bwd(dt=BF16):
    g = torch.Generator(device="cuda").manual_seed(6)
    n, c, h, w = 2, 64, 12, 20
    tdt = torch_dtype(dt)
    s = torch.randn((n, c, h, w), generator=g, device="cuda").clamp_min(0).to(tdt).float()
    hw_ = torch.randn((3, c), generator=g, device="cuda") * 0.2
    hb = torch.randn((3,), generator=g, device="cuda") * 0.1
    gy = torch.randn((n, 3, h, w), generator=g, device="cuda")
    sd = s.double().requires_grad_(True)
    wd = hw_.double().requires_grad_(True)
    bd = hb.double().requires_grad_(True)
    pre = sd  # s is already post-ReLU; gradient is masked where s == 0
    y = torch.tanh(torch.einsum("nchw,jc->njhw", pre, wd) + bd.view(1, 3, 1, 1))
    y.backward(gy.double())
    gs_exp = sd.grad * (s > 0)
    dw = torch.zeros((3, c), device="cuda")
    db = torch.zeros((3,), device="cuda")
    dbp = torch.zeros((c,), device="cuda")
    gs = P8.empty(n, c, h, w, dt)
    ops.head_bwd(gy, y.detach().float().contiguous(), P8.from_nchw(s, dt), hw_, dt, dw=dw, db=db, gs=gs, dbias_prev=dbp)
    torch.cuda.synchronize()
    e1 = (dw.double() - wd.grad).abs().max().item() / max(1.0, wd.grad.abs().max().item())
    e2 = (db.double() - bd.grad).abs().max().item() / max(1.0, bd.grad.abs().max().item())
    e3 = (gs.to_nchw().double() - gs_exp).abs().max().item()
    e4 = (dbp.double() - gs.to_nchw().double().sum((0, 2, 3))).abs().max().item()
    ok = e1 < 1e-4 and e2 < 1e-4 and e3 < (2e-2 if dt == BF16 else 2e-3) and e4 < 1e-2
    cs = torch.zeros((c,), device="cuda")
    ops.channel_sum(gs, cs, dt)
    torch.cuda.synchronize()
    ok &= (cs.double() - gs.to_nchw().double().sum((0, 2, 3))).abs().max().item() < 1e-2
    return ok, max(e1, e2, e3), f"dw={e1:.3g} db={e2:.3g} gs={e3:.3g} dbp={e4:.3g}"


def check_grad_scale():
    x = torch.randn(100000, device="cuda") * 3e-5
    amax = torch.zeros(1, device="cuda")
    s2 = torch.zeros(2, device="cuda")
    ops.absmax(x, amax)
    ops.make_grad_scale(amax, 64.0, s2)
    torch.cuda.synchronize()
    a = x.abs().max().item()
    ok = abs(amax.item() - a) == 0.0
    s = s2[0].item()
    ok &= (a * s <= 64.0) and (a * s > 32.0) and abs(s2[1].item() * s - 1.0) < 1e-6
    return ok, 0.0, f"amax={a:.3g} scale={s:.3g}"


def ref_cut_patch(t: torch.Tensor, y: int, x: int, size: int) -> torch.Tensor:
    """restatement of StyleTransferDataset._cut_patch (reference src/data/dataset.py:209-232)"""
    half = size // 2
    hn, hx = max(0, y - half), min(y + half, t.size(1) - 1)
    xn, xx = max(0, x - half), min(x + half, t.size(2) - 1)
    patch = t[:, hn:hx, xn:xx]
    if patch.size(1) != size or patch.size(2) != size:
        res = torch.zeros((t.size(0), size, size), device=t.device)
        res[:, :patch.size(1), :patch.size(2)] = patch
        patch = res
    return patch


def check_gather(patch=32, n_patches=37, seed=7):
    g = torch.Generator(device="cuda").manual_seed(seed)
    sizes = [(70, 90), (64, 48), (101, 77)]
    n_src = 3
    imgs = [[torch.randn((3, h, w), generator=g, device="cuda") for (h, w) in sizes] for _ in range(n_src)]
    table = torch.tensor([[im.data_ptr() for im in src] for src in imgs], dtype=torch.int64, device="cuda")
    hw = torch.tensor(sizes, dtype=torch.int32, device="cuda")
    pos = []
    cpu_g = torch.Generator().manual_seed(seed)
    for b in range(n_patches):
        i = int(torch.randint(0, 3, (1,), generator=cpu_g))
        y = int(torch.randint(0, sizes[i][0], (1,), generator=cpu_g))
        x = int(torch.randint(0, sizes[i][1], (1,), generator=cpu_g))
        pos.append((i, y, x))
    # force the reference's edge cases
    pos[0] = (0, 0, 0)
    pos[1] = (0, sizes[0][0] - 1, sizes[0][1] - 1)
    pos[2] = (1, 5, 5)
    pos[3] = (2, sizes[2][0] - 10, sizes[2][1] - 20)
    post = torch.tensor(pos, dtype=torch.int32, device="cuda")
    comb = torch.full((n_patches, 6, patch, patch), float("nan"), device="cuda")
    other = torch.full((n_patches, 3, patch, patch), float("nan"), device="cuda")
    ops.patch_gather(table, n_src, 3, 3, hw, post, patch, [comb, comb, other], [0, 3, 0], [6, 6, 3])
    torch.cuda.synchronize()
    ok = True
    for b, (i, y, x) in enumerate(pos):
        ok &= torch.equal(comb[b, 0:3], ref_cut_patch(imgs[0][i], y, x, patch))
        ok &= torch.equal(comb[b, 3:6], ref_cut_patch(imgs[1][i], y, x, patch))
        ok &= torch.equal(other[b], ref_cut_patch(imgs[2][i], y, x, patch))
    return bool(ok), 0.0, ""


def check_conv_norm_on_load(n=2, cpre=64, cin=32, cout=32, h=37, w=45, T=2, blk_c=32, dt=FP16, act="relu", seed=13, pair=False, cps=0):
    """conv3x3(pad 1) over cat(act(pre*scale+shift), x): the normalisation + activation of `pre` applied in-kernel"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    tdt = torch_dtype(dt)
    pre = torch.randn((n, cpre, h, w), generator=g, device="cuda").to(tdt).float()
    sc = torch.rand((n, cpre), generator=g, device="cuda") + 0.5
    sh = torch.randn((n, cpre), generator=g, device="cuda") * 0.3
    y = pre * sc[:, :, None, None] + sh[:, :, None, None]
    y = {"relu": F.relu, "leaky": lambda t: F.leaky_relu(t, 0.2), "none": lambda t: t}[act](y).to(tdt).float()
    xs = [y]
    x = None
    if cin:
        x = torch.randn((n, cin, h, w), generator=g, device="cuda").to(tdt).float()
        xs.append(x)
    wt = (torch.randn((cout, cpre + cin, 3, 3), generator=g, device="cuda") * 0.05).to(tdt).float()
    exp = ref_conv(torch.cat(xs, 1), wt, 1, 1)
    out = P8.empty(n, cout, h, w, dt)
    ops.conv_fwd(P8.from_nchw(x, dt) if cin else None, ops.pack_conv_weight(wt, cpre + cin, blk_c, dt, pair=pair), cout, 3, 3, 1, 1,
                 dt, blk_c=blk_c, tiles_per_cta=T, out=out, cta_pair=pair, ctas_per_sm=cps, pre=P8.from_nchw(pre, dt), pre_scale=sc.contiguous(),
                 pre_shift=sh.contiguous(), pre_act={"relu": ACT_RELU, "leaky": ACT_LEAKY, "none": ACT_NONE}[act])
    torch.cuda.synchronize()
    err = (out.to_nchw().double() - exp).abs().max().item()
    scale = max(1.0, exp.abs().max().item())
    ok = err / scale < (3e-2 if dt == BF16 else 4e-3)
    return ok, err, f"err={err:.4g}"


def check_conv_upsample(n=2, cin=64, cout=32, h=20, w=28, T=2, blk_c=32, dt=FP16, seed=11, stats=True, pair=False, raw_c=0):
    """conv3x3(pad 1) over the bilinear x2 (align_corners=True) upsample of a low-res input, interpolated in-kernel"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    tdt = torch_dtype(dt)
    x = torch.randn((n, cin, h, w), generator=g, device="cuda").to(tdt).float()
    wt = (torch.randn((cout, cin, 3, 3), generator=g, device="cuda") * 0.05).to(tdt).float()
    kw_raw = {}
    xa = x
    if raw_c:   # the first raw_c channels are raw conv outputs: relu(x*scale + shift) is applied to the low-res tile on load
        sc = torch.rand((n, raw_c), generator=g, device="cuda") + 0.5
        sh = torch.randn((n, raw_c), generator=g, device="cuda") * 0.3
        xa = x.clone()
        xa[:, :raw_c] = torch.relu(x[:, :raw_c] * sc[:, :, None, None] + sh[:, :, None, None]).to(tdt).float()
        kw_raw = dict(pre_scale=sc, pre_shift=sh, pre_act=ACT_RELU, up_raw_channels=raw_c)
    up = F.interpolate(xa, scale_factor=2, mode="bilinear", align_corners=True).to(tdt).float()
    exp = ref_conv(up, wt, 1, 1)
    out = P8.empty(n, cout, 2 * h, 2 * w, dt)
    tiles = ops.conv_num_tiles(2 * h, 2 * w, T)
    part = torch.full((n, tiles, 2, cout), float("nan"), device="cuda") if stats else None
    ops.conv_fwd(P8.from_nchw(x, dt), ops.pack_conv_weight(wt, cin, blk_c, dt, pair=pair), cout, 3, 3, 1, 1, dt, blk_c=blk_c,
                 tiles_per_cta=T, out=out, stats_partial=part, upsample2x=True, cta_pair=pair, **kw_raw)
    torch.cuda.synchronize()
    got = out.to_nchw().double()
    err = (got - exp).abs().max().item()
    scale = max(1.0, exp.abs().max().item())
    ok = err / scale < (3e-2 if dt == BF16 else 4e-3)
    msg = f"err={err:.4g}"
    if stats:
        s_exp = torch.stack([got.sum((2, 3)), (got * got).sum((2, 3))], 1)
        serr = ((part.double().sum(1) - s_exp).abs() / (1.0 + s_exp.abs())).max().item()
        ok &= serr < 1e-4
        msg += f" stats={serr:.3g}"
    return ok, err, msg


def check_fused_clip_adam(steps=5, clip=0.5, wd=1e-5, seed=3):
    """FusedClipAdam vs clip_grad_norm_ + torch.optim.Adam on the same gradients (reference lightning_model.py:245-250)"""
    from pbt_b200.optim import FusedClipAdam
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = [(32, 3, 7, 7), (32,), (64, 32, 3, 3), (128, 128, 3, 3), (3, 64, 1, 1), (3,), (5,)]
    pa = [torch.nn.Parameter(torch.randn(s, generator=g, device="cuda") * 0.05) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = torch.optim.Adam(pa, lr=4e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    ob = FusedClipAdam(pb, lr=4e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd, max_grad_norm=clip)
    worst, norm_err = 0.0, 0.0
    for it in range(steps):
        scale = 10.0 if it % 2 == 0 else 0.01      # alternate between clipped and un-clipped steps
        grads = [torch.randn(s, generator=g, device="cuda") * scale for s in shapes]
        for p, q, gr in zip(pa, pb, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        tn = torch.nn.utils.clip_grad_norm_(pa, clip) if clip else None
        oa.step()
        ob.step()
        torch.cuda.synchronize()
        if tn is not None:
            norm_err = max(norm_err, abs(float(tn) - float(ob.last_grad_norm)) / float(tn))
        for p, q in zip(pa, pb):
            worst = max(worst, ((p - q).abs().max() / (p.abs().max() + 1e-12)).item())
    sa, sb = oa.state[pa[3]], ob.state[pb[3]]
    m_err = ((sa["exp_avg"] - sb["exp_avg"]).abs().max() / (sa["exp_avg"].abs().max() + 1e-20)).item()
    v_err = ((sa["exp_avg_sq"] - sb["exp_avg_sq"]).abs().max() / (sa["exp_avg_sq"].abs().max() + 1e-30)).item()
    ok = worst < 2e-6 and norm_err < 1e-5 and m_err < 2e-6 and v_err < 2e-6 and float(sb["step"]) == steps
    return ok, worst, f"param rel err={worst:.3g} norm rel err={norm_err:.3g} m={m_err:.3g} v={v_err:.3g} step={float(sb['step'])}"


def check_fused_l1(seed=4, weight=4.0):
    from pbt_b200.optim import fused_l1_loss
    g = torch.Generator(device="cuda").manual_seed(seed)
    y = (torch.rand((7, 3, 20, 24), generator=g, device="cuda") * 2 - 1).requires_grad_(True)
    t = torch.rand((7, 3, 20, 24), generator=g, device="cuda") * 2 - 1
    t[0, 0, 0, :4] = y.detach()[0, 0, 0, :4]          # exact ties: sign(0) = 0 like torch
    y2 = y.detach().clone().requires_grad_(True)
    la = fused_l1_loss(y, t, weight)
    (la * 1.5).backward()
    lb = torch.nn.functional.l1_loss(y2, t) * weight
    (lb * 1.5).backward()
    e1 = abs(float(la) - float(lb)) / abs(float(lb))
    e2 = (y.grad - y2.grad).abs().max().item() / y2.grad.abs().max().item()
    return e1 < 1e-5 and e2 < 1e-6, e1, f"loss rel err={e1:.3g} grad rel err={e2:.3g}"
