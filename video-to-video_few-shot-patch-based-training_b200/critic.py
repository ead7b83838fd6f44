"""Native path of the PatchGAN critic ``DiscriminatorN_IN`` (reference src/models/discriminator.py:8-149), used by the
adversarial branch of ``training_step`` (reference lightning_model.py:224-236,277-283,294-319).

The critic is tiny (3->12->24->48->1 channels in the shipped configuration, 0.2 % of the generator's FLOPs): through the
tensor library it costs ~240 launches per step.  Here it runs on the generator's own kernels in ~12 launches per forward
pass: channels are zero-padded to multiples of 16 (padding rows / columns of the packed weights are zero, so padded
channels stay exactly zero), and

  * a 4x4 stride-2 pad-1 stage is a 3x3 stride-1 pad-1 implicit-GEMM conv over the space-to-depth copy of its input
    (output row r reads input rows 2r-1..2r+2 = block r-1 phase 1, block r phases 0/1, block r+1 phase 0; the zero
    (tap, phase) combinations are zero weights - `pbt_pack_weights` mode bit 3), the space-to-depth copy being written by
    the producing `pbt_norm_apply` launch;
  * a 4x4 stride-1 pad-1 stage (``pre_output``, ``output``) shrinks the map by one pixel: it runs on the fixed grid of
    its input with pad (1, 1), the conv's valid window zeroes the extra row / column and keeps it out of the InstanceNorm
    statistics, and that zero border is the next layer's padding;
  * InstanceNorm + LeakyReLU are the generator's `pbt_norm_finalize` / `pbt_norm_apply` / `pbt_norm_bwd`, weight gradients
    `pbt_conv_wgrad`, data gradients the conv kernel with tap-flipped weights.

fp16 operands, fp32 accumulation; the 16-bit gradient tensors carry a power-of-two scale chosen from max|dL/dlogits| (the
MSE gradient of a 160-patch batch is ~1e-5, below fp16's normal range).  No CPU path.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
from torch import Tensor

from . import ops
from ._native import ACT_LEAKY, ACT_NONE, BF16, FP16, P8
from .parallel import GradBucket


def _p16(c: int) -> int:
    return (c + 15) // 16 * 16


def _blk(cin: int) -> int:
    return 32 if cin % 32 == 0 or cin > 32 else 16


class _Stage:
    """one conv stage of the critic: geometry + parameter names"""

    def __init__(self, name: str, conv, cin: int, cout: int, stride: int, norm: bool, act: bool, bias_name: Optional[str]):
        self.name, self.conv, self.cin, self.cout, self.stride, self.norm, self.act = name, conv, cin, cout, stride, norm, act
        self.cin_p, self.cout_p = _p16(cin), _p16(cout)
        self.wname = name + ".0.weight"
        self.bname = bias_name


def supported(disc) -> Optional[str]:
    """None when the native path can run this module, else the reason"""
    convs = [disc.initial[0]] + [s[0] for s in disc.intermediate] + [disc.pre_output[0], disc.output[0]]
    if any(len(s) < 2 or not isinstance(s[1], torch.nn.InstanceNorm2d) for s in list(disc.intermediate) + [disc.pre_output]):
        return "norm_layer is not instance_norm"
    if any(4 * _p16(c.in_channels) > 256 or _p16(c.out_channels) > 256 for c in convs):
        return "a layer is wider than 256 (padded) channels"
    return None


class CriticEngine:
    def __init__(self, disc, operand_dtype: str = "fp16"):
        self.disc = disc
        self.dt = {"fp16": FP16, "bf16": BF16}[operand_dtype]
        self.device = next(disc.parameters()).device
        use_bias = disc.initial[0].bias is not None
        st: List[_Stage] = []
        c0 = disc.initial[0]
        st.append(_Stage("initial", c0, c0.in_channels, c0.out_channels, 2, False, True, "initial.0.bias" if use_bias else None))
        for i, blk in enumerate(disc.intermediate):
            c = blk[0]
            st.append(_Stage(f"intermediate.{i}", c, c.in_channels, c.out_channels, 2, True, True, f"intermediate.{i}.0.bias" if use_bias else None))
        c = disc.pre_output[0]
        st.append(_Stage("pre_output", c, c.in_channels, c.out_channels, 1, True, True, "pre_output.0.bias" if use_bias else None))
        c = disc.output[0]
        st.append(_Stage("output", c, c.in_channels, c.out_channels, 1, False, False, "output.0.bias" if use_bias else None))
        self.stages = st
        self._ws: Dict[Any, Dict[str, Any]] = {}
        self._packer = None
        self._pack_key = None
        self._saved = None
        self._stamp = 0
        self.bucket: Optional[GradBucket] = None
        self.grad_scale_target = 32.0 if self.dt == FP16 else 0.0

    # ------------------------------------------------------------------ operands
    def _weights(self):
        params = list(self.disc.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packer is None or self._packer["ptrs"] != tuple(p.data_ptr() for p in params):
            pk = ops.WeightPacker(self.device)
            for s in self.stages:
                w = s.conv.weight.detach()
                if s.stride == 2:
                    pk.add(s.name, w, s2d=True, s2d4_cpp=s.cin_p, k_pad=4 * s.cin_p, n_out=s.cout_p, n_keep=s.cout,
                           blk_c=_blk(4 * s.cin_p), dt=self.dt)
                    pk.add(s.name + ".d", w, s2d=True, dgrad=True, s2d4_cpp=s.cin_p, k_pad=s.cout_p, n_out=4 * s.cin_p,
                           n_keep=4 * s.cin_p, blk_c=_blk(s.cout_p), dt=self.dt)
                else:
                    pk.add(s.name, w, k_pad=s.cin_p, n_out=s.cout_p, n_keep=s.cout, blk_c=_blk(s.cin_p), dt=self.dt)
                    pk.add(s.name + ".d", w, dgrad=True, k_pad=s.cout_p, n_out=s.cin_p, n_keep=s.cin, blk_c=_blk(s.cout_p), dt=self.dt)
            bias = {s.name: torch.zeros(s.cout_p, device=self.device) for s in self.stages if s.bname and not s.norm}
            self._packer = {"pk": pk, "ptrs": tuple(p.data_ptr() for p in params), "bias": bias}
            self._pack_key = None
        # while a CUDA graph is being captured the packing is always recorded: a replayed graph must see the parameters as
        # they are at replay time (warm-up roll-back, load_state_dict between replays), not a host-side cache decision
        if self._pack_key != key or torch.cuda.is_current_stream_capturing():
            self._packer["pk"].run()
            for s in self.stages:        # biases in front of an InstanceNorm cancel; the others go into the conv epilogue
                if s.name in self._packer["bias"]:
                    self._packer["bias"][s.name][:s.cout].copy_(s.conv.bias.detach())
            self._pack_key = key
        return self._packer["pk"].out, self._packer["bias"]

    def _workspace(self, n: int, h: int, w: int) -> Dict[str, Any]:
        key = (n, h, w)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        if len(self._ws) >= 4:
            self._ws.clear()
        dev, dt = self.device, self.dt
        E = lambda c, hh, ww, zero=False: P8.empty(n, c, hh, ww, dt, device=dev, zero=zero)  # noqa: E731
        ws = {"n": n, "h": h, "w": w, "stat": {}}
        s0 = self.stages[0]
        ws["x_p8"] = E(s0.cin_p, h, w)
        hh, ww = h, w
        for s in self.stages:
            if s.stride == 2:
                ws[s.name + ".in"] = E(4 * s.cin_p, hh // 2, ww // 2)          # space-to-depth copy of the stage input
                hh, ww = hh // 2, ww // 2
            else:
                ws.setdefault(s.name + ".in", E(s.cin_p, hh, ww))
            ws[s.name + ".grid"] = (hh, ww)
            ws[s.name + ".raw"] = E(s.cout_p, hh, ww)
        ws["logits32"] = torch.empty((n, self.stages[-1].cout_p // 8, hh, ww, 8), device=dev)
        self._ws[key] = ws
        return ws

    def _stat(self, ws, s: _Stage, tiles: int):
        st = ws["stat"].get(s.name)
        if st is None or st["tiles"] != tiles:
            n, dev = ws["n"], self.device
            st = dict(tiles=tiles, partial=torch.empty((n, tiles, 2, s.cout_p), device=dev),
                      scale=torch.empty((n, s.cout_p), device=dev), shift=torch.empty((n, s.cout_p), device=dev))
            ws["stat"][s.name] = st
        return st

    @staticmethod
    def _bt(n: int, h: int, w: int, cout_p: int) -> int:
        return 2 if (n >= 2 and h * w <= 1600 and (cout_p + 31) // 32 * 32 * 2 <= 256) else 0

    def _conv(self, xin: P8, wpack, s_cout_p: int, k: int, pad: int, **kw):
        bt = self._bt(xin.n, xin.h, xin.w, s_cout_p)
        ops.conv_fwd(xin, wpack, s_cout_p, k, k, pad, pad, self.dt, blk_c=_blk(xin.c), tiles_per_cta=2, ctas_per_sm=0 if bt else 4,
                     batch_tiles=bool(bt), **kw)
        return 1 if bt else 2           # tiles-per-CTA that index stats_partial

    # ------------------------------------------------------------------ forward
    def forward(self, x: Tensor, save: bool) -> Tensor:
        n, cin, h, w = x.shape
        L2 = sum(1 for s in self.stages if s.stride == 2)
        if h % (1 << L2) or w % (1 << L2):
            raise ValueError(f"critic input {h}x{w} must be divisible by {1 << L2}")
        dt = self.dt
        W, B = self._weights()
        ws = self._workspace(n, h, w)
        x = x.contiguous()
        if x.dtype not in (torch.float32, torch.float16):
            x = x.float()
        ops.nchw_to_p8(x, ws["x_p8"], dt)
        cur: P8 = ws["x_p8"]          # activated output of the previous stage (standard layout)
        cur_is_input = True
        vh, vw = h, w                  # valid window of `cur` on its grid
        for i, s in enumerate(self.stages):
            raw = ws[s.name + ".raw"]
            gh, gw = ws[s.name + ".grid"]
            if s.stride == 2:
                xin = ws[s.name + ".in"]
                if cur_is_input:   # the first stage: identity + space-to-depth of the converted input
                    ops.norm_apply(cur, dt, act=ACT_NONE, out_s2d=xin)
                k, pad, ovh, ovw = 3, 1, gh, gw
            else:
                xin = ws[s.name + ".in"]
                k, pad, ovh, ovw = 4, 1, vh - 1, vw - 1
            window = None if (ovh == gh and ovw == gw) else (ovh, ovw)
            last = i == len(self.stages) - 1
            if s.norm:
                st = self._stat(ws, s, ops.conv_num_tiles(gh, gw, 1 if self._bt(n, gh, gw, s.cout_p) else 2))
                self._conv(xin, W[s.name], s.cout_p, k, pad, out=raw, stats_partial=st["partial"], valid_hw=window)
                ops.norm_finalize(st["partial"], n, st["tiles"], s.cout_p, ovh * ovw, st["scale"], st["shift"], eps=1e-5)
                sc, sh = st["scale"], st["shift"]
            elif last:
                self._conv(xin, W[s.name], s.cout_p, k, pad, bias=B.get(s.name), out32=ws["logits32"], valid_hw=window)
                sc = sh = None
            else:
                self._conv(xin, W[s.name], s.cout_p, k, pad, bias=B.get(s.name), out=raw, valid_hw=window)
                sc = sh = None
            vh, vw = ovh, ovw
            if last:
                break
            nxt = self.stages[i + 1]
            dst = ws[nxt.name + ".in"]
            if nxt.stride == 2:
                ops.norm_apply(raw, dt, scale=sc, shift=sh, act=ACT_LEAKY if s.act else ACT_NONE, out_s2d=dst)
            else:
                ops.norm_apply(raw, dt, scale=sc, shift=sh, act=ACT_LEAKY if s.act else ACT_NONE, out=dst)
                if window is not None:
                    ops.zero_border(dst, ovh, ovw)      # act(0*scale + shift) != 0: restore the zero padding border
            cur_is_input = False
        gh, gw = ws[self.stages[-1].name + ".grid"]
        full = torch.empty((n, 1, gh, gw), device=self.device)
        ops.p8f_to_nchw(ws["logits32"], 1, full)
        if save:
            self._saved = (ws, W, (vh, vw))
            self._stamp += 1
        return full[:, :, :vh, :vw]

    # ------------------------------------------------------------------ backward
    def grad_bucket(self) -> GradBucket:
        params = list(self.disc.named_parameters())
        if self.bucket is None or any(a is not b for a, (_, b) in zip(self.bucket.params, params)):
            self.bucket = GradBucket(params)
        return self.bucket

    def backward(self, glog: Tensor, want_params: bool, want_input: bool):
        """glog: dL/dlogits [n,1,vh,vw] fp32 -> ({param name: grad view}, dL/dx or None)"""
        ws, W, (vh, vw) = self._saved
        n, h, w, dt, dev = ws["n"], ws["h"], ws["w"], self.dt, self.device
        S = self.stages
        gh, gw = ws[S[-1].name + ".grid"]
        pool = torch.zeros(self._pool_floats(n), device=dev)
        off = [0]

        def Z(*shape):
            cnt = 1
            for d_ in shape:
                cnt *= d_
            lo = off[0]
            off[0] = lo + (cnt + 63) // 64 * 64
            return pool[lo:lo + cnt].view(shape)

        inv = gscale = None
        g_full = torch.nn.functional.pad(glog.float(), (0, gw - vw, 0, gh - vh))
        if self.grad_scale_target > 0:
            amax, scale2 = torch.empty(1, device=dev), torch.empty(2, device=dev)
            ops.absmax(g_full, amax)
            ops.make_grad_scale(amax, self.grad_scale_target, scale2)
            gscale, inv = scale2[0:1], scale2[1:2]
            g_full = g_full * gscale
        bucket = self.grad_bucket()
        if want_params and bucket.aliased_by_param_grads():
            bucket = GradBucket(list(self.disc.named_parameters()))
        grads: Dict[str, Tensor] = {}
        E = lambda c, hh, ww: P8.empty(n, c, hh, ww, dt, device=dev)  # noqa: E731
        g = E(S[-1].cout_p, gh, gw)                      # gradient w.r.t. the raw output of the current stage
        ops.nchw_to_p8(g_full.contiguous(), g, dt)
        for i in range(len(S) - 1, -1, -1):
            s = S[i]
            sgh, sgw = ws[s.name + ".grid"]
            xin = ws[s.name + ".in"]
            k, pad = (3, 1) if s.stride == 2 else (4, 1)
            if want_params:
                dw = Z(k * k, xin.c, s.cout_p)
                ops.conv_wgrad(xin, g, k, k, pad, pad, dt, dw, inv_scale=inv)
                dst = bucket.views[s.wname]
                if s.stride == 2:
                    _s2d4_wgrad_to_param(dw, dst, s.cin_p)
                else:
                    dst.view(s.cout, s.cin, -1).copy_(dw[:, :s.cin, :s.cout].permute(2, 1, 0))
                grads[s.wname] = dst
                if s.bname is not None and s.norm:
                    grads[s.bname] = bucket.views[s.bname]   # in front of an InstanceNorm: zero gradient (the slice stays zero)
                if s.bname is not None and not s.norm and i == len(S) - 1:
                    # the last stage (no activation behind it): bias gradient = channel sum of dL/dlogits; the other
                    # norm-free stage gets its bias gradient from the activation backward below
                    db = Z(s.cout_p)
                    ops.channel_sum(g, db, dt, inv_scale=inv)
                    bucket.views[s.bname].copy_(db[:s.cout])
                    grads[s.bname] = bucket.views[s.bname]
            if i == 0 and not want_input:
                break
            # data gradient: conv with the tap-flipped, channel-transposed kernel
            n_in = xin.c
            g_in = E(n_in, sgh, sgw)
            prev_window = None
            if s.stride == 1:
                # the input map's own valid window (one pixel larger than this stage's output window)
                pvh, pvw = self._in_window(ws, i)
                prev_window = None if (pvh == sgh and pvw == sgw) else (pvh, pvw)
            self._conv(g, W[s.name + ".d"], n_in, k, k - 1 - pad, out=g_in, valid_hw=prev_window)
            if i == 0:
                gx = torch.empty((n, s.cin, h, w), device=dev)
                ops.p8s2d_to_nchw(g_in, s.cin_p, s.cin, gx, dt, mul=inv)
                return grads, gx
            # through the activation (+ InstanceNorm) of the previous stage
            p = S[i - 1]
            praw = ws[p.name + ".raw"]
            pgh, pgw = ws[p.name + ".grid"]
            g_prev = E(p.cout_p, pgh, pgw)
            is_s2d = s.stride == 2
            if p.norm:
                st = ws["stat"][p.name]
                pvh, pvw = self._out_window(ws, i - 1)
                sums = Z(n, 2, p.cout_p)
                ops.norm_bwd(praw, dt, scale=st["scale"], shift=st["shift"], act=ACT_LEAKY if p.act else ACT_NONE, ga=g_in,
                             ga_is_s2d=is_s2d, sums=sums, kmul=st["scale"], count=pvh * pvw, dx=g_prev)
                if (pvh, pvw) != (pgh, pgw):
                    ops.zero_border(g_prev, pvh, pvw)
            else:
                # conv (+bias) -> LeakyReLU: dx = g * act'(x); the pooled reduce returns sum(dx) = the bias gradient
                one, zero = self._ident(p.cout_p)
                sums = Z(2, p.cout_p)

                def bias_grad(sums_, p=p):
                    if want_params and p.bname is not None:
                        dstb = bucket.views[p.bname]
                        if inv is not None:
                            torch.mul(sums_[0, :p.cout], inv, out=dstb)
                        else:
                            dstb.copy_(sums_[0, :p.cout])
                        grads[p.bname] = dstb
                    sums_.zero_()

                ops.norm_bwd(praw, dt, scale=one, shift=zero, per_channel=True, act=ACT_LEAKY if p.act else ACT_NONE, ga=g_in,
                             ga_is_s2d=is_s2d, sums=sums, kmul=one, count=n * pgh * pgw, batch_mode=True, dx=g_prev, between=bias_grad)
            g = g_prev
        return grads, None

    def _ident(self, c: int):
        if getattr(self, "_id", None) is None or self._id[0].numel() < c:
            self._id = (torch.ones(max(c, 256), device=self.device), torch.zeros(max(c, 256), device=self.device))
        return self._id[0][:c], self._id[1][:c]

    def _out_window(self, ws, i: int):
        """valid output window of stage i on its grid"""
        vh, vw = ws["h"], ws["w"]
        for s in self.stages[:i + 1]:
            vh, vw = (vh // 2, vw // 2) if s.stride == 2 else (vh - 1, vw - 1)
        return vh, vw

    def _in_window(self, ws, i: int):
        return self._out_window(ws, i - 1) if i > 0 else (ws["h"], ws["w"])

    def _pool_floats(self, n: int) -> int:
        tot = 4096
        for s in self.stages:
            k = 9 if s.stride == 2 else 16
            cin = 4 * s.cin_p if s.stride == 2 else s.cin_p
            tot += k * cin * s.cout_p + 64 + n * 2 * s.cout_p + 64 + 2 * s.cout_p + 64 + s.cout_p + 64
        return tot


_KY = ((0, 1), (1, 0), (1, 1), (2, 0))     # original 4x4 tap ky -> (3x3 tap over space-to-depth, phase)
_S2D4_IDX: Dict[Any, Any] = {}


def _s2d4_wgrad_to_param(dw: Tensor, dst: Tensor, cpp: int) -> None:
    """wgrad of a 4x4 stride-2 stage run as 3x3 conv over the space-to-depth input: [9, 4*cpp, cout_p] -> [cout, cin, 4, 4]
    (one gather + one strided copy)"""
    cout, cin = dst.shape[0], dst.shape[1]
    idx = _S2D4_IDX.get(dw.device)
    if idx is None:
        t = torch.tensor([a for a, _ in _KY], device=dw.device)
        ph = torch.tensor([b for _, b in _KY], device=dw.device)
        idx = _S2D4_IDX[dw.device] = (t[:, None], t[None, :], ph[:, None], ph[None, :])
    d6 = dw.view(3, 3, 2, 2, cpp, dw.shape[2])              # (ty, tx, py, px, c, co)
    g = d6[idx[0], idx[1], idx[2], idx[3]]                  # (ky, kx, c, co)
    dst.copy_(g[:, :, :cin, :cout].permute(3, 2, 0, 1))


class _CriticFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng, *params):
        ctx.eng = eng
        ctx.want_input = x.requires_grad
        ctx.want_params = any(p.requires_grad for p in params)
        ctx.x_dtype = x.dtype
        y = eng.forward(x, save=True)
        ctx.stamp = eng._stamp
        return y

    @staticmethod
    def backward(ctx, gy):
        eng = ctx.eng
        if ctx.stamp != eng._stamp:
            raise RuntimeError("native critic: backward of a forward pass whose saved activations were overwritten by a later "
                               "grad-enabled forward (one outstanding pass per module)")
        grads, gx = eng.backward(gy.contiguous(), ctx.want_params, ctx.want_input)
        out = []
        for name, p in eng.disc.named_parameters():
            gr = grads.get(name) if (ctx.want_params and p.requires_grad) else None
            out.append(None if gr is None else gr.detach())
        return (None if gx is None else gx.to(ctx.x_dtype), None, *out)


def critic_forward(disc, eng: CriticEngine, x: Tensor) -> Tensor:
    params = list(disc.parameters())
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        return _CriticFn.apply(x, eng, *params)
    return eng.forward(x, save=False)
