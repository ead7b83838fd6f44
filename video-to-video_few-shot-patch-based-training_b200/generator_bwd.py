"""Backward pass of the native GeneratorJ (autograd of reference src/models/generator.py:210-239, which
fires inside ``manual_backward`` at lightning_model.py:241).

Hand-scheduled reverse sweep over the buffers the forward pass saved:
  per conv   : wgrad (tcgen05, K = pixels) + dgrad (the forward implicit-GEMM kernel with tap-flipped,
               channel-transposed weights; ReLU masks / residual fan-in fused into its epilogue)
  per norm   : InstanceNorm/BatchNorm backward with the activation derivative, skip-connection fan-in and the
               space-to-depth indexing of the stride-2 layers folded in: one fused launch per norm on patch-sized
               maps, reduce + apply kernels for batch statistics and large maps
  upsample   : gather form of the transposed bilinear x2
Weight gradients run on a side stream next to the data-gradient chain; parameter gradients are published at three
join points (all-reduce groups).  All 16-bit gradient tensors carry a power-of-two scale S chosen from max|dL/dy|
(fp16 operands only, with a device-side overflow back-off); the fp32 parameter gradients are unscaled inside the
producing kernels.
"""
from __future__ import annotations

from typing import List

import torch

from . import ops
from ._native import ACT_LEAKY, ACT_NONE, ACT_RELU, P8



def _wgrad_to_param(dw: torch.Tensor, dst: torch.Tensor) -> None:
    """wgrad layout [taps, cin_pad, cout] -> parameter layout [cout, cin, kh, kw], written into `dst` (a bucket view)"""
    cout, cin = dst.shape[0], dst.shape[1]
    dst.view(cout, cin, -1).copy_(dw[:, :cin].permute(2, 1, 0))


def _s2d_wgrad_to_param(dw: torch.Tensor, dst: torch.Tensor) -> None:
    """wgrad of a stride-2 conv run as 2x2 conv over the space-to-depth input: [4, 4*cin, cout] -> [cout, cin, 3, 3]
    (inverse of ops.s2d_weight; the nine (tap, phase) pairs cover the nine 3x3 taps exactly once)"""
    cout, cin = dst.shape[0], dst.shape[1]
    dw2 = dw.permute(2, 1, 0).reshape(cout, 4 * cin, 2, 2)
    for ty, py, dy in ops._S2D_TAPS:
        for tx, px, dx in ops._S2D_TAPS:
            ph = py * 2 + px
            dst[:, :, dy, dx].copy_(dw2[:, ph * cin:(ph + 1) * cin, ty, tx])


def generator_backward(eng, gy: torch.Tensor, y: torch.Tensor, want_input_grad: bool = False):
    g, dt, dev = eng.gen, eng.dt, eng.device
    ws, W = eng._saved
    f = g.filters
    cp = eng.cin_p
    n, h, w = ws.n, ws.h, ws.w
    h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
    nb = len(g.resnet_blocks)
    if nb == 0:
        raise NotImplementedError("native backward needs at least one residual block")
    E = lambda c, hh, ww: P8.empty(n, c, hh, ww, dt, device=dev)  # noqa: E731

    # every accumulate-into buffer of the sweep (wgrad outputs, reduction sums, bias gradients) is carved out of ONE
    # zero-filled allocation: a single memset instead of ~70 fill launches per step
    pool = torch.zeros(eng.zero_pool_floats(n), device=dev)
    pool_off = [0]

    def Z(*shape):
        cnt = 1
        for d_ in shape:
            cnt *= d_
        cnt_al = (cnt + 63) // 64 * 64          # keep every carve 256-byte aligned (vector reductions need 16 B)
        lo = pool_off[0]
        if lo + cnt_al > pool.numel():
            return torch.zeros(shape, device=dev)
        pool_off[0] = lo + cnt_al
        return pool[lo:lo + cnt].view(shape)

    # ---- gradient scale (fp16 only): S = 2^k with max|gy|*S in (target/2, target]
    scale2 = None
    gscale = inv = None
    if eng.grad_scale_target > 0:
        amax = torch.empty(1, device=dev)
        scale2 = torch.empty(2, device=dev)
        ops.absmax(gy, amax)
        ops.make_grad_scale(amax, eng.grad_scale_target, scale2, adjust=eng.grad_scale_adjust())
        gscale, inv = scale2[0:1], scale2[1:2]

    def T_of(pref, ww):
        return eng._T(pref, ww)

    def dgrad(name, gin: P8, cout, k, pad, out=None, T_pref=2, **kw):
        """conv of the output gradient with the transposed/flipped kernel -> gradient of the conv input"""
        acc_stride = (cout + 31) // 32 * 32
        T = T_of(T_pref, gin.w)
        # two co-resident CTAs per SM need <= 256 tensor-memory columns each (allocations are powers of two)
        while T > 1 and T * acc_stride > 256:
            T -= 1
        if kw.get("cta_pair"):
            T = 1
        bt = 0 if kw.get("cta_pair") else eng._bt(gin.n, gin.h, gin.w, cout)
        if bt:
            T = bt
        # small maps / narrow outputs get the small-footprint configuration (the kernel falls back when it does not fit)
        ops.conv_fwd(gin, W[name + ".d"], cout, k, k, k - 1 - pad, k - 1 - pad, dt, blk_c=eng._blk(gin.c), tiles_per_cta=T,
                     out=out, ctas_per_sm=0 if (kw.get("cta_pair") or bt) else 4, concurrent=True, batch_tiles=bool(bt), **kw)

    # Weight gradients run on a side stream: a conv's wgrad and dgrad are independent, and on patch-sized maps neither
    # fills the GPU (a 128->128 3x3 wgrad is ~180 CTAs, its dgrad ~480 of 592 slots).  Their parameter gradients are
    # published (grad hook -> all-reduce bucket) at the join points, on the main stream.
    main = torch.cuda.current_stream(dev)
    if getattr(eng, "_pack_bwd_done", None) is not None:
        main.wait_event(eng._pack_bwd_done)      # the data-gradient weights were packed on a side stream during the forward pass
        eng._pack_bwd_done = None
    sides = eng.side_streams()
    n_side = [0]
    keep = []
    last_dw = [None]

    def wgrad_side(name, x: P8, dy: P8, k, pad, convert=_wgrad_to_param):
        """enqueue wgrad(x, dy) + `convert(dw, bucket view)` on the side stream after everything enqueued so far on the
        main stream and publish the gradient from there (a data-parallel exchange that this gradient completes is then
        ordered behind the side stream without holding up the data-gradient chain); returns the event that marks the side
        job's completion (wait for it before overwriting x or dy)"""
        dw = Z(k * k, x.c, dy.c)
        side = sides[n_side[0] % len(sides)]
        n_side[0] += 1
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            ops.conv_wgrad(x, dy, k, k, pad, pad, dt, dw, inv_scale=inv)
            convert(dw, bucket.views[name])
            last_dw[0] = dw
            grads[name] = bucket.views[name]
            done = torch.cuda.Event()
            done.record(side)
        keep.append((x, dy))
        return done

    def join():
        """main stream waits for the side streams"""
        for side in sides:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)
        keep.clear()

    bn_mode, no_norm = eng.norm_mode() == "batch", eng.norm_mode() == "none"
    norm_mods, norm_names, conv_names = eng.norm_modules(), eng.norm_param_names(), eng.conv_param_names()

    def in_bwd(x: P8, st, act, dx: P8, count, name=None, **kw):
        if no_norm:
            # conv (+bias) -> activation: dx = g * act'(x).  The norm-backward kernels do exactly that with identity tables
            # and zero sums in the apply; the pooled reduce returns S1 = sum(dx), the gradient of the conv bias.
            one, zero = torch.ones(x.c, device=dev), torch.zeros(x.c, device=dev)
            sums = Z(2, x.c)

            def bias_grad(sums_):
                if norm_mods[name][0].bias is not None:
                    s1 = sums_[0].clone()
                    put(conv_names[name] + ".bias", s1, inv)
                sums_.zero_()

            ops.norm_bwd(x, dt, scale=one, shift=zero, per_channel=True, act=act, sums=sums, kmul=one, count=n * count,
                         batch_mode=True, dx=dx, between=bias_grad, **kw)
            return
        if not bn_mode:
            sums = Z(n, 2, x.c)
            ops.norm_bwd(x, dt, scale=st["scale"], shift=st["shift"], act=act, sums=sums, kmul=st["scale"], count=count, dx=dx,
                         **kw)
            return
        # BatchNorm2d (batch statistics) + activation.  The kernels evaluate the activation derivative at x*scale+shift,
        # so they get the post-affine value y = gamma*xhat+beta there; the reduce then returns S1 = sum(gact) and
        # S2 = sum(gact*y), from which A = sum(gact*xhat) = (S2 - beta*S1)/gamma gives d(gamma) = A, d(beta) = S1, and the
        # apply, which computes k*(gact - s1/cnt - y*s2/cnt), yields the BatchNorm input gradient
        # gamma*rstd*(gact - S1/cnt - xhat*A/cnt) when it is handed s1 = S1 - beta*A/gamma, s2 = A/gamma, k = gamma*rstd.
        bn = norm_mods[name][1]
        gamma, beta = bn.weight.detach().float(), bn.bias.detach().float()
        gsafe = torch.where(gamma.abs() < 1e-12, torch.full_like(gamma, 1e-12), gamma)
        sc = (gamma * st["rstd"]).contiguous()
        sh = (beta - st["mean"] * sc).contiguous()
        sums = Z(2, x.c)

        def fix(sums_):
            s1, s2 = sums_[0].clone(), sums_[1].clone()
            a = (s2 - beta * s1) / gsafe
            put(norm_names[name] + ".weight", a, inv)
            put(norm_names[name] + ".bias", s1, inv)
            sums_[0].copy_(s1 - beta * a / gsafe)
            sums_[1].copy_(a / gsafe)

        ops.norm_bwd(x, dt, scale=sc, shift=sh, per_channel=True, act=act, sums=sums, kmul=sc, count=n * count,
                     batch_mode=True, dx=dx, between=fix, **kw)

    hook = getattr(eng, "grad_hook", None)   # data-parallel training: parallel.GradAllReduce.grad_ready

    known = {k for k, _ in g.named_parameters()}      # use_bias=False: the bias sums below have no parameter to go to

    # Every parameter gradient is written straight into its slice of the engine's flat fp32 bucket (parallel.GradBucket):
    # autograd adopts the views as p.grad (no copy), the data-parallel all-reduce and the fused optimiser read them in
    # place, and their addresses never change.  If p.grad still aliases the bucket (gradient accumulation: no zero_grad
    # since the last sweep) this sweep goes to a fresh one, which autograd then adds to p.grad.
    bucket = eng.grad_bucket()
    if bucket.aliased_by_param_grads():
        from .parallel import GradBucket
        bucket = GradBucket(list(g.named_parameters()))

    class _Grads(dict):
        def __setitem__(self, k, v):
            if k not in known:
                return
            dst = bucket.views[k]
            if v is not dst:
                dst.copy_(v.reshape(dst.shape))
            super().__setitem__(k, dst)
            if hook is not None:
                hook(k, dst)

    grads = _Grads()

    def put(name, src, scale=None):
        """publish `src` (optionally times the device scalar `scale`) as the gradient of parameter `name`"""
        if name not in known:
            return
        dst = bucket.views[name]
        if scale is not None:
            torch.mul(src.reshape(dst.shape), scale, out=dst)
            grads[name] = dst
        else:
            grads[name] = src

    # a bias in front of an affine-less InstanceNorm has exactly zero gradient: publish those first
    # (the same holds in front of a train-mode BatchNorm; the norm's own beta is not one of these).  Their bucket slices
    # are zero from the allocation on and never written.
    silent = set() if no_norm else {id(conv.bias) for conv, _ in norm_mods.values() if conv.bias is not None}
    for name, p in g.named_parameters():
        if id(p) in silent:
            grads[name] = bucket.views[name]

    if g.append_smoothers:
        # ---- head (1x1 + tanh) and the ReLU of smoothers.3
        dw_out, db_out, db_s3 = Z(3, f[5]), Z(3), Z(f[5])
        g_s3 = E(f[5], h, w)
        ops.head_bwd(gy, y, ws.s3, W["head_w"], dt, gscale=gscale, head_tanh=g.use_tanh, dw=dw_out, db=db_out, gs=g_s3,
                     dbias_prev=db_s3)
        put("output.0.weight", dw_out, inv)
        put("output.0.bias", db_out, inv)
        # ---- smoothers.3
        wgrad_side("smoothers.3.weight", ws.s0n, g_s3, 3, 1)
        put("smoothers.3.bias", db_s3, inv)
        g_s0n = E(f[5], h, w)
        dgrad("smooth3", g_s3, f[5], 3, 1, out=g_s0n, T_pref=3)
        # ---- BatchNorm (batch statistics) + the ReLU in front of it
        st = ws.stats["bn"]
        bn = g.smoothers[2]
        bn_rstd, bn_mean = st["rstd"], st["mean"]
        bn_sums = Z(2, f[5])
        g_s0 = E(f[5], h, w)
        ops.norm_bwd(ws.s0, dt, scale=bn_rstd, shift=(-bn_mean * bn_rstd).contiguous(), per_channel=True, act=ACT_NONE, ga=g_s0n,
                     sums=bn_sums, kmul=(bn.weight.detach().float() * bn_rstd).contiguous(), count=n * h * w, batch_mode=True,
                     dx=g_s0, relu_mask_x=True)
        put("smoothers.2.weight", bn_sums[1], inv)
        put("smoothers.2.bias", bn_sums[0], inv)
        db_s0 = Z(f[5])
        ops.channel_sum(g_s0, db_s0, dt, inv_scale=inv)
        # ---- smoothers.0
        wgrad_side("smoothers.0.weight", ws.c11, g_s0, 3, 1)
        put("smoothers.0.bias", db_s0)
        g_c11 = E(f[5], h, w)
        dgrad("smooth0", g_s0, f[5], 3, 1, out=g_c11, T_pref=3, mask=ws.c11)
        db_11 = Z(f[5])
        ops.channel_sum(g_c11, db_11, dt, inv_scale=inv)
    else:
        # ---- head (1x1 + tanh) directly on conv11's ReLU output (append_smoothers=False)
        dw_out, db_out, db_11 = Z(3, f[5]), Z(3), Z(f[5])
        g_c11 = E(f[5], h, w)
        ops.head_bwd(gy, y, ws.c11, W["head_w"], dt, gscale=gscale, head_tanh=g.use_tanh, dw=dw_out, db=db_out, gs=g_c11,
                     dbias_prev=db_11)
        put("output.0.weight", dw_out, inv)
        put("output.0.bias", db_out, inv)
        if inv is not None:
            db_11 = db_11 * inv
    # ---- conv11 (input = cat11 = [up1 | conv0 | x])
    wgrad_side("conv11.0.weight", ws.cat11, g_c11, 7, 3)
    put("conv11.0.bias", db_11)
    if want_input_grad:
        # the same data gradient over ALL of cat11's channels: [up1 | conv0 | x]; the x slot is one source of dL/dx
        g_cat = E(eng.cat11x_channels(), h, w)
        dgrad("conv11x", g_c11, g_cat.c, 7, 3, out=g_cat, T_pref=3)
    else:
        g_cat = E(f[4] + f[0], h, w)
        dgrad("conv11", g_c11, f[4] + f[0], 7, 3, out=g_cat, T_pref=3, cta_pair=eng.pair11_dgrad())
    join()   # tail group complete (output head, smoothers, conv11)
    del g_c11
    if g.append_smoothers:
        del g_s0, g_s0n, g_s3
    # ---- upsample1 block
    g_rawU1 = E(f[4], h, w)
    in_bwd(ws.rawU1, ws.stats["up1"], ACT_RELU, g_rawU1, h * w, name="up1", ga=g_cat.view(0, f[4]))
    wgrad_side("upsample1.1.weight", ws.u1in, g_rawU1, 3, 1)
    g_u1in = E(f[4] + f[1], h, w)
    dgrad("up1", g_rawU1, f[4] + f[1], 3, 1, out=g_u1in)
    del g_rawU1
    g_c1cat = E(f[4] + f[1], h2, w2)
    ops.upsample2x_bwd(g_u1in, dt, gin16=g_c1cat)
    del g_u1in
    # ---- upsample2 block
    g_rawU2 = E(f[4], h2, w2)
    in_bwd(ws.rawU2, ws.stats["up2"], ACT_RELU, g_rawU2, h2 * w2, name="up2", ga=g_c1cat.view(0, f[4]))
    wgrad_side("upsample2.1.weight", ws.u2in, g_rawU2, 3, 1)
    g_u2in = E(2 * f[2], h2, w2)
    dgrad("up2", g_rawU2, 2 * f[2], 3, 1, out=g_u2in)
    g_r = torch.empty((n, f[2] // 8, h4, w4, 8), device=dev)          # fp32 gradient of the residual stream
    g_c2skip = E(f[2], h4, w4)
    ops.upsample2x_bwd(g_u2in.view(0, f[2]), dt, gin32=g_r)
    ops.upsample2x_bwd(g_u2in.view(f[2], f[2]), dt, gin16=g_c2skip)
    join()   # decoder group complete
    del g_u2in, g_rawU2
    # ---- residual blocks, last to first.  Two gradient buffers (one per conv of a block): the side-stream wgrad of
    # block b still reads g_rawB / g_rawA while the main stream moves on; a buffer is rewritten only after the wgrad
    # that read it has finished (event wait - normally already satisfied).
    g_rawB, g_rawA = E(f[2], h4, w4), E(f[2], h4, w4)
    g_h = E(f[2], h4, w4)
    evB = evA = None
    for b in range(nb - 1, -1, -1):
        if evB is not None:
            main.wait_event(evB)
        in_bwd(ws.rawB[b], ws.stats[f"res{b}.b"], ACT_NONE, g_rawB, h4 * w4, name=f"res{b}.b", gb32=g_r)
        evB = wgrad_side(conv_names[f"res{b}.b"] + ".weight", ws.hmid[b], g_rawB, 3, 1)
        dgrad(f"res{b}.b", g_rawB, f[2], 3, 1, out=g_h)
        if evA is not None:
            main.wait_event(evA)
        in_bwd(ws.rawA[b], ws.stats[f"res{b}.a"], ACT_RELU, g_rawA, h4 * w4, name=f"res{b}.a", ga=g_h)
        evA = wgrad_side(conv_names[f"res{b}.a"] + ".weight", ws.a[b], g_rawA, 3, 1)
        # g_r <- g_r + relu'(r_b) * dgrad   (in place: every element is read then written by the same thread)
        dgrad(f"res{b}.a", g_rawA, f[2], 3, 1, out=None, mask=ws.a[b], addend32=g_r, out32=g_r)
    # ---- downsample2 (conv2 feeds the residual stream and the decoder skip)
    g_raw2 = E(f[2], h4, w4)
    in_bwd(ws.raw2, ws.stats["down2"], ACT_LEAKY, g_raw2, h4 * w4, name="down2", gb16=g_c2skip, gb32=g_r)
    wgrad_side("downsample2.0.weight", ws.s2d1, g_raw2, 2, 1, _s2d_wgrad_to_param)               # dw: [4, 4*f1, f2]
    g_s2d1 = E(4 * f[1], h4, w4)
    dgrad("down2", g_raw2, 4 * f[1], 2, 1, out=g_s2d1)
    # ---- downsample1
    g_raw1 = E(f[1], h2, w2)
    in_bwd(ws.raw1, ws.stats["down1"], ACT_LEAKY, g_raw1, h2 * w2, name="down1", ga=g_s2d1, ga_is_s2d=True, gb16=g_c1cat.view(f[4], f[1]))
    wgrad_side("downsample1.0.weight", ws.s2d0, g_raw1, 2, 1, _s2d_wgrad_to_param)
    g_s2d0 = E(4 * f[0], h2, w2)
    dgrad("down1", g_raw1, 4 * f[0], 2, 1, out=g_s2d0)
    # ---- initial conv
    g_raw0 = E(f[0], h, w)
    in_bwd(ws.raw0, ws.stats["initial"], ACT_LEAKY, g_raw0, h * w, name="initial", ga=g_s2d0, ga_is_s2d=True, gb16=g_cat.view(f[4], f[0]))
    wgrad_side("initial_conv.0.weight", ws.cat11.view(f[4] + f[0], cp), g_raw0, 7, 3)
    gx = None
    if want_input_grad:
        # dL/dx = (x slot of conv11's data gradient) + (data gradient of the initial conv); never used by the reference
        # training, so it is assembled with tensor-library glue instead of a fused epilogue
        g_x0 = E(cp, h, w)
        dgrad("initial", g_raw0, cp, 7, 3, out=g_x0)
        ci = g.input_channels
        gx = g_cat.view(f[4] + f[0], cp).to_nchw()[:, :ci] + g_x0.to_nchw()[:, :ci]
        if inv is not None:
            gx = gx * inv
    join()   # trunk group complete
    if eng.grad_scale_target > 0:
        # an fp16 overflow anywhere in the sweep ends up as inf/nan in this most-downstream gradient: shrink the scale
        # of the next sweep (the optimiser skips a non-finite step - FusedClipAdam does; AMP semantics)
        # (probed in its wgrad accumulator: the bucket slice may already be under the data-parallel all-reduce)
        ops.grad_scale_feedback(last_dw[0], eng.grad_scale_adjust())

    # ---- assemble in parameter order
    out: List[torch.Tensor] = []
    for name, p in g.named_parameters():
        gr = grads.get(name)
        if gr is None:
            raise RuntimeError(f"no gradient produced for {name}")
        # a fresh alias per call: autograd adopts a returned gradient as p.grad without a copy only when nothing else
        # references that tensor object (the bucket keeps its own views)
        out.append(gr.detach() if gr.dtype == p.dtype else gr.to(p.dtype))
    return out, gx
