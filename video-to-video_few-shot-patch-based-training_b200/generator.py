"""Drop-in ``GeneratorJ`` (reference src/models/generator.py:60-239) running on hand-written sm_100a kernels.

Same constructor, same ``state_dict`` keys/shapes (the module tree is rebuilt in the reference's
registration order so that a given torch seed initialises the weights bit-for-bit identically), same
``forward(x[N,Cin,H,W]) -> [N,3,H,W]`` contract incl. autograd.  The nn.Conv2d / norm sub-modules are only
parameter containers: the arithmetic runs through libpbt.so (tcgen05 implicit-GEMM convolutions with fused
epilogues, HBM-bound elementwise kernels) in the P8 activation layout.  No PyTorch / CPU fallback exists:
on a non-CUDA tensor forward raises.

Data flow of one forward (reference :210-239; fused differently):
  x -> P8 into the tail channels of the conv11 input buffer `cat11` (torch.cat :230 is never executed)
  initial 7x7 conv (+IN stats) -> norm+lrelu -> cat11[conv0 slot] and a space-to-depth copy
  downsample convs as 2x2 stride-1 convs over the space-to-depth tensors (stride 2 without strided TMA)
  7 residual blocks: conv -> IN stats -> norm(+relu / +residual, fp32 residual stream)
  decoder: bilinear x2 of [out, skip] written straight into the next conv's input buffer (cat :228-229)
  conv11 7x7 (+bias, relu) -> smoothers (BN: batch stats in train, folded affine in eval) -> 1x1+tanh fused
  into the last conv's epilogue.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from ._native import ACT_LEAKY, ACT_NONE, ACT_RELU, BF16, FP16, P8

EPS = 1e-5


class UpsamplingLayer(nn.Module):
    """parameter-free x2 bilinear (align_corners=True) stage; kept so module indices match the reference"""

    def __init__(self, channels: int):
        super().__init__()
        self.layer = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)


class ResNetBlock(nn.Module):
    """container for `block.{1,4}` conv parameters (reference :18-58)"""

    def __init__(self, channels: int, norm_layer: Optional[str] = "instance_norm", use_bias: bool = False):
        super().__init__()
        norm = _norm_cls(norm_layer)
        seq: List[nn.Module] = []
        for _ in range(2):
            seq += [nn.ReLU(inplace=False), nn.Conv2d(channels, channels, 3, 1, 1, bias=use_bias)]
            if norm is not None:
                seq.append(norm(channels))
        self.block = nn.Sequential(*seq)


def _norm_cls(name):
    return {"batch_norm": nn.BatchNorm2d, "instance_norm": nn.InstanceNorm2d}.get(name)


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


class GeneratorJ(nn.Module):
    #: operand dtype of the tensor-core path: "bf16" | "fp16" (class default; per-instance attribute overrides)
    operand_dtype = "fp16"

    def __init__(self, input_channels: int = 3, additional_channels: Optional[Dict[str, Any]] = None,
                 filters: List[int] = [32, 64, 128, 128, 128, 64], norm_layer: str = "instance_norm",
                 use_bias: bool = False, resnet_blocks: int = 7, tanh: bool = True, append_smoothers: bool = True,
                 input_size: int = 256):
        super().__init__()
        filters = [int(f) for f in filters]
        self.input_size = input_size
        self.append_smoothers = append_smoothers
        self.input_channels = int(input_channels)
        self.filters = filters
        self.norm_layer = norm_layer
        self.use_tanh = bool(tanh)
        norm = _norm_cls(norm_layer)

        def conv_block(cin, cout, k, s, p, act):
            mods: List[nn.Module] = [nn.Conv2d(cin, cout, k, s, p, bias=use_bias)]
            if norm is not None:
                mods.append(norm(cout))
            mods.append(act)
            return nn.Sequential(*mods)

        def up_block(cin, cout):
            mods: List[nn.Module] = [UpsamplingLayer(cin), nn.Conv2d(cin, cout, 3, 1, 1, bias=use_bias)]
            if norm is not None:
                mods.append(norm(cout))
            mods.append(nn.ReLU(inplace=False))
            return nn.Sequential(*mods)

        f = filters
        self.initial_conv = conv_block(self.input_channels, f[0], 7, 1, 3, nn.LeakyReLU(0.2, inplace=False))
        self.downsample1 = conv_block(f[0], f[1], 3, 2, 1, nn.LeakyReLU(0.2, inplace=False))
        self.downsample2 = conv_block(f[1], f[2], 3, 2, 1, nn.LeakyReLU(0.2, inplace=False))
        self.resnet_blocks = nn.ModuleList([ResNetBlock(f[2], norm_layer, use_bias) for _ in range(int(resnet_blocks))])
        self.upsample2 = up_block(f[2] + f[2], f[4])
        self.upsample1 = up_block(f[4] + f[1], f[4])
        self.conv11 = nn.Sequential(nn.Conv2d(f[0] + f[4] + self.input_channels, f[5], 7, 1, 3, bias=use_bias),
                                    nn.ReLU(inplace=False))
        if append_smoothers:
            self.smoothers = nn.Sequential(nn.Conv2d(f[5], f[5], 3, padding=1, bias=use_bias), nn.ReLU(inplace=False),
                                           nn.BatchNorm2d(f[5]),
                                           nn.Conv2d(f[5], f[5], 3, padding=1, bias=use_bias), nn.ReLU(inplace=False))
        head: List[nn.Module] = [nn.Conv2d(f[5], 3, 1, bias=True)]
        if tanh:
            head.append(nn.Tanh())
        self.output = nn.Sequential(*head)
        self.apply(self._init_weights)
        self._engine: Optional[_Engine] = None

    @staticmethod
    def _init_weights(m: nn.Module):
        # reference :149-154 — N(0, 0.02) conv weights, zero biases
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.normal_(m.weight.data, 0.0, 0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)

    # ------------------------------------------------------------------ forward
    def _check_supported(self):
        f = self.filters
        if any(c % 16 for c in (f[0], f[1], f[2], f[4], f[5])) or max(f) > 256 or 2 * f[2] > 256:
            raise NotImplementedError(f"native GeneratorJ needs filter counts that are multiples of 16 and <= 256, got {f}")

    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("GeneratorJ (B200-native) has no CPU path: move the module and its input to a CUDA device")
        if x.dim() != 4 or x.shape[1] != self.input_channels:
            raise ValueError(f"expected input [N,{self.input_channels},H,W], got {tuple(x.shape)}")
        if x.shape[2] % 4 or x.shape[3] % 4:
            raise ValueError("H and W must be multiples of 4 (skip connections of the reference mismatch otherwise)")
        self._check_supported()
        if self._engine is None:
            self._engine = _Engine(self)
        params = _param_list(self)
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        if need_grad:
            y = _GeneratorFn.apply(x, self, *params)
        else:
            y = self._engine.forward(x, save=False)
        return y if y.dtype == x.dtype else y.to(x.dtype)


def _param_list(g: GeneratorJ) -> List[nn.Parameter]:
    return list(g.parameters())


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gen, *params):
        ctx.gen = gen
        ctx.x_needs_grad, ctx.x_dtype = x.requires_grad, x.dtype
        if x.requires_grad:
            gen._engine.input_grad = True      # also packs the two data-gradient kernels that reach the input
        y = gen._engine.forward(x, save=True)
        ctx.stamp = gen._engine._saved_stamp
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        if ctx.stamp != ctx.gen._engine._saved_stamp:
            # the activations of a forward pass live in ONE workspace per engine: a second grad-enabled forward has
            # overwritten them (gradient accumulation over several forwards, G(x1) and G(x2) in one loss)
            raise RuntimeError("native GeneratorJ: backward of a forward pass whose saved activations were overwritten by a "
                               "later grad-enabled forward; call backward() before the next forward (one outstanding pass per module)")
        grads, gx = ctx.gen._engine.backward(gy.contiguous().float(), y, want_input_grad=ctx.x_needs_grad)
        return (None if gx is None else gx.to(ctx.x_dtype), None, *grads)


class _Workspace:
    """all device buffers for one (N, H, W) problem size"""

    def __init__(self, eng: "_Engine", n: int, h: int, w: int, train: bool):
        g, dt = eng.gen, eng.dt
        f = g.filters
        cp = eng.cin_p
        dev = eng.device
        self.n, self.h, self.w, self.train = n, h, w, train
        h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
        nb = len(g.resnet_blocks)
        E = lambda c, hh, ww, zero=False: P8.empty(n, c, hh, ww, dt, device=dev, zero=zero)  # noqa: E731
        self.cat11 = E(f[4] + f[0] + cp, h, w, zero=True)
        self.raw0 = E(f[0], h, w)
        self.s2d0 = E(4 * f[0], h2, w2)
        self.raw1 = E(f[1], h2, w2)
        self.c1cat = E(f[4] + f[1], h2, w2)
        self.s2d1 = E(4 * f[1], h4, w4)
        self.raw2 = E(f[2], h4, w4)
        self.c2cat = E(2 * f[2], h4, w4)
        k = nb if train else 1
        self.rawA = [E(f[2], h4, w4) for _ in range(k)]
        self.rawB = [E(f[2], h4, w4) for _ in range(k)]
        self.hmid = [E(f[2], h4, w4) for _ in range(k)]
        self.a = [E(f[2], h4, w4) for _ in range((nb + 1) if train else 2)]
        # residual stream: fp32 while training (its gradient sums seven branches); the inference pass keeps it in 16 bits
        # like the reference's own `.half()` CUDA inference (generator.py:185): 8 instead of 12 bytes per element in the
        # seven HBM-bound residual updates of a frame
        self.r = [torch.empty((n, f[2] // 8, h4, w4, 8), device=dev) for _ in range(2)] if (train or not eng.residual16) else None
        self.r16 = None if (train or not eng.residual16) else [E(f[2], h4, w4) for _ in range(2)]
        self.u2in = E(2 * f[2], h2, w2) if train else None   # inference upsamples inside the conv kernel
        self.rawU2 = E(f[4], h2, w2)
        self.u1in = E(f[4] + f[1], h, w) if train else None
        self.rawU1 = E(f[4], h, w)
        self.c11 = E(f[5], h, w)
        self.s0 = E(f[5], h, w)
        self.s0n = E(f[5], h, w)
        self.s3 = E(f[5], h, w) if train else None
        # norm statistics: name -> (partial, scale, shift)
        self.stats: Dict[str, Any] = {}

    def stat(self, name: str, c: int, hh: int, ww: int, T: int, dev):
        if name not in self.stats:
            tiles = ops.conv_num_tiles(hh, ww, T)
            self.stats[name] = dict(tiles=tiles, partial=torch.empty((self.n, tiles, 2, c), device=dev),
                                    scale=torch.empty((self.n, c), device=dev), shift=torch.empty((self.n, c), device=dev))
        return self.stats[name]


class _Engine:
    """orchestrates the native kernels for one GeneratorJ instance (host glue only)"""

    def __init__(self, gen: GeneratorJ):
        self.gen = gen
        od = getattr(gen, "operand_dtype", "fp16")
        self.dt = {"bf16": BF16, "fp16": FP16}[od]
        self.cin_p = _pad16(gen.input_channels)
        self.device = next(gen.parameters()).device
        self._ws: Dict[Any, _Workspace] = {}
        self._wslots: Dict[bool, Any] = {}
        self.grad_scale_target = 32.0 if self.dt == FP16 else 0.0  # dynamic power-of-two gradient scaling (fp16 only), see backward()
        self.input_grad = False   # set once an input with requires_grad is seen: the backward sweep then also produces dL/dx
        self.grad_hook = None     # callable(name, grad) fired as each parameter gradient is produced (data parallel)
        # inference-time kernel configuration switches (tools/infer_knobs.py measures them): CTA-pair configuration for the
        # 128-channel layers and the smoothers
        env = lambda k, d: os.environ.get(k, d) == "1"  # noqa: E731
        self.pair_res, self.pair_up, self.pair_smooth = env("PBT_PAIR_RES", "0"), env("PBT_PAIR_UP", "0"), env("PBT_PAIR_SMOOTH", "0")
        # inference: up2's norm + ReLU applied inside up1's upsample-on-load (built and tested; measured 20.25 -> 20.36 ms per
        # 4-frame pass - up1's transform warps are on its critical path, the saved 0.2-ms normalise pass does not pay: off)
        self.fold_up = env("PBT_FOLD_UP", "0")
        # training: InstanceNorm finalize inside the consuming norm_apply launch (built and tested; measured at C3 5.69 -> 5.72 ms
        # per step - the per-CTA tile sum costs more than the launch that programmatic dependent launch already hides: off)
        self.fuse_finalize = env("PBT_FUSE_FINALIZE", "0")
        self.ws_up, self.ws_res = env("PBT_WS_UP", "0"), env("PBT_WS_RES", "0")   # weight-stationary MMA runs on the N = 128 layers
        self.use_tap_pairs = env("PBT_TAP_PAIRS", "1")
        self.residual16 = env("PBT_RESIDUAL16", "1")    # inference: 16-bit residual stream (see _Workspace.r16)
        self.batch_tiles = env("PBT_BATCH_TILES", "1")   # small maps: one CTA = the same tile of two images (see _bt)
        # training passes on patch-sized maps (128 channels, h/4 * (w/4 + 2) <= 512): the whole residual trunk is ONE launch, one
        # CTA per image with the running activation resident in shared memory (csrc/res_trunk.cu) instead of 6 launches per block
        self.fused_trunk = env("PBT_FUSED_TRUNK", "1")
        self._pack_bwd_done = None   # event: the side-stream packing of the data-gradient weights of this step has finished
        self._saved_stamp = 0     # counts grad-enabled forward passes: a backward must match the pass that saved its activations
        self.bucket = None        # parallel.GradBucket: flat fp32 storage the backward sweep writes the parameter gradients into
        self.kernel_timer = None  # bench.py: list collecting (start event, end event, frames) around the dominant kernel (conv11)

    # -------------------------------------------------------------- helpers
    def workspace(self, n, h, w, train) -> _Workspace:
        key = (n, h, w, train)
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) >= 4:
                self._ws.clear()
            ws = _Workspace(self, n, h, w, train)
            self._ws[key] = ws
        return ws

    @staticmethod
    def _T(pref: int, w: int) -> int:
        """tiles per CTA: the largest T <= pref that wastes the fewest padded columns (8T-pixel wide CTA tiles)"""
        best, best_pad = 1, None
        for t in range(1, pref + 1):
            pad = -(-w // (8 * t)) * 8 * t
            if best_pad is None or pad <= best_pad:
                best, best_pad = t, pad
        return best

    # The short-CTA layers (initial conv, first strided conv, smoothers: cout <= 64) request the kernel's small-footprint
    # configuration: four co-resident CTAs per SM hide the per-CTA load / epilogue phases (tools/conv_occ.py: 10-25 %).
    SMOOTH_BLK = 16
    SMOOTH_T = 2

    def _bt(self, n: int, oh: int, ow: int, cout: int) -> int:
        """batch-tile factor for a conv on [n, *, oh, ow] maps: on patch-sized maps (<= 40x40) two images share one CTA, so
        the packed weights are streamed from L2 once per two tiles (the 128->128 3x3 convs on 80 x 20x20 maps are bound by
        that stream: 141 MB per launch, tools/conv_timeline.py) and no 8-pixel tile columns are wasted.  0 = off."""
        if not self.batch_tiles or n < 2 or oh * ow > 40 * 40 or (cout + 31) // 32 * 32 * 2 > 256:
            return 0
        return 2

    def tap_pairs(self) -> bool:
        """first layer with <= 8 input channels (all real channels in one 8-channel plane): one K = 16 MMA covers that plane
        at two horizontally adjacent taps, 28 instead of 49 MMAs per tile and half the input bytes (env PBT_TAP_PAIRS=0: off)"""
        return self.use_tap_pairs and self.gen.input_channels <= 8 and self.cin_p == 16

    def _T11(self, w: int) -> int:
        t = self._T(3, w)
        return 2 if (t == 1 and self.pair11()) else t   # the pair configuration is built for 2 or 3 tiles per CTA

    def pair11(self) -> bool:
        """conv11 (the dominant layer, cout 64) runs in the CTA-pair configuration: one M=256 cta_group::2 MMA stream per
        two CTAs, each staging half of the weight columns (+3-5 % on that layer, tools/conv_occ.py)"""
        return self.gen.filters[5] % 32 == 0

    def pair11_dgrad(self) -> bool:
        """conv11's data gradient (64 -> 160 channels, 7x7): 160 accumulator columns per tile leave room for one tile per
        CTA at two CTAs per SM, and then every CTA streams the whole 1 MB of weights for 128 pixels (L2-bound).  In the
        CTA-pair configuration the two CTAs of a cluster share the weight stream: 624 -> 434 us at the C3 shape."""
        f = self.gen.filters
        return _pad16(f[4] + f[0]) % 32 == 0 and self._blk(f[5]) == 32

    def norm_mode(self) -> str:
        """'instance' | 'batch' | 'none' — reference src/models/generator.py:83-87 (any other string means no norm layers)"""
        return {"instance_norm": "instance", "batch_norm": "batch"}.get(self.gen.norm_layer, "none")

    def res_conv_index(self):
        """positions of the two convs inside ResNetBlock.block: (1, 4) with norm layers, (1, 3) without (reference :31-56)"""
        return (1, 4) if self.norm_mode() != "none" else (1, 3)

    def norm_modules(self):
        """engine layer name -> (conv module, norm module or None) for every conv of a (conv -> norm -> act) block"""
        g = self.gen
        has = self.norm_mode() != "none"
        ia, ib = self.res_conv_index()
        m = {"initial": (g.initial_conv[0], g.initial_conv[1] if has else None),
             "down1": (g.downsample1[0], g.downsample1[1] if has else None),
             "down2": (g.downsample2[0], g.downsample2[1] if has else None),
             "up2": (g.upsample2[1], g.upsample2[2] if has else None), "up1": (g.upsample1[1], g.upsample1[2] if has else None)}
        for i, blk in enumerate(g.resnet_blocks):
            m[f"res{i}.a"] = (blk.block[ia], blk.block[ia + 1] if has else None)
            m[f"res{i}.b"] = (blk.block[ib], blk.block[ib + 1] if has else None)
        return m

    def conv_param_names(self):
        """engine layer name -> state_dict prefix of its conv module ('initial_conv.0', 'resnet_blocks.3.block.4', ...)"""
        ia, ib = self.res_conv_index()
        m = {"initial": "initial_conv.0", "down1": "downsample1.0", "down2": "downsample2.0", "up2": "upsample2.1",
             "up1": "upsample1.1"}
        for i in range(len(self.gen.resnet_blocks)):
            m[f"res{i}.a"] = f"resnet_blocks.{i}.block.{ia}"
            m[f"res{i}.b"] = f"resnet_blocks.{i}.block.{ib}"
        return m

    def norm_param_names(self):
        """engine layer name -> state_dict prefix of its norm module ('initial_conv.1', 'resnet_blocks.3.block.5', ...)"""
        return {k: v.rsplit(".", 1)[0] + "." + str(int(v.rsplit(".", 1)[1]) + 1) for k, v in self.conv_param_names().items()}

    def cat11x_channels(self) -> int:
        """channel count of the conv11 data gradient when it also covers the x slot: [up1 | conv0 | x] padded to 32"""
        f = self.gen.filters
        return (f[4] + f[0] + self.cin_p + 31) // 32 * 32

    def _blk(self, cin: int) -> int:
        return 32 if cin % 32 == 0 or cin > 32 else 16

    def _pairs(self, train: bool):
        """(residual convs, decoder convs, smoothers) run in the CTA-pair configuration - inference passes only"""
        f = self.gen.filters
        ok128 = f[2] % 32 == 0 and f[4] % 32 == 0
        return (self.pair_res and ok128 and not train, self.pair_up and ok128 and not train,
                self.pair_smooth and f[5] % 32 == 0 and f[5] <= 64 and not train)

    def _weights(self, with_dgrad: bool):
        """packed 16-bit conv operands, rebuilt by ONE native launch whenever a parameter changed (version counters)"""
        g = self.gen
        key = tuple((p.data_ptr(), p._version) for p in g.parameters())
        slot_key = (with_dgrad, with_dgrad and self.input_grad, self._pairs(with_dgrad))
        slot = self._wslots.get(slot_key)
        if slot is None or slot["ptrs"] != tuple(p.data_ptr() for p in g.parameters()):
            slot = self._build_packer(with_dgrad)
            self._wslots[slot_key] = slot
        # (under CUDA-graph capture the packing is always recorded: a replayed graph must follow the parameters as they are
        # at replay time, not a host-side cache decision taken at capture time)
        if slot["key"] != key or torch.cuda.is_current_stream_capturing():
            slot["packer"].run()
            self._pack_bwd_done = None
            if slot["packer_bwd"] is not None:
                main = torch.cuda.current_stream(self.device)
                side = self.side_streams()[0]
                fork = torch.cuda.Event()
                fork.record(main)
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    slot["packer_bwd"].run()
                    self._pack_bwd_done = torch.cuda.Event()
                    self._pack_bwd_done.record(side)
            slot["key"] = key
        return slot["W"]

    def _build_packer(self, with_dgrad: bool):
        g, dt, cp = self.gen, self.dt, self.cin_p
        f = g.filters
        pk = ops.WeightPacker(self.device)
        # the data-gradient forms are first read by the backward sweep: they get their own launch, which runs on a side
        # stream underneath the forward pass instead of in front of it (see _weights)
        pkd = ops.WeightPacker(self.device) if with_dgrad else None

        def fwd(name, conv, k_pad, s2d=False, blk=None, pair=False):
            co = conv.weight.shape[0]
            pk.add(name, conv.weight.detach(), s2d=s2d, k_pad=k_pad, n_out=co, n_keep=co, blk_c=blk or self._blk(k_pad), dt=dt,
                   pair=pair)

        def dgr(name, conv, s2d=False, keep=None, pair=False, n_out=None):
            co, ci = conv.weight.shape[0], conv.weight.shape[1]
            vi = 4 * ci if s2d else ci
            n_keep = vi if keep is None else keep
            pkd.add(name + ".d", conv.weight.detach(), s2d=s2d, dgrad=True, k_pad=co, n_out=n_out or _pad16(n_keep), n_keep=n_keep,
                    blk_c=self._blk(co), dt=dt, pair=pair)

        pr, pu, ps = self._pairs(with_dgrad)
        if self.tap_pairs():
            pk.add("initial", g.initial_conv[0].weight.detach(), k_pad=16, n_out=f[0], n_keep=f[0], blk_c=16, dt=dt, tap_pairs=True)
        else:
            fwd("initial", g.initial_conv[0], cp)
        fwd("down1", g.downsample1[0], 4 * f[0], s2d=True)
        fwd("down2", g.downsample2[0], 4 * f[1], s2d=True)
        for i, blk in enumerate(g.resnet_blocks):
            ia, ib = self.res_conv_index()
            fwd(f"res{i}.a", blk.block[ia], f[2], pair=pr)
            fwd(f"res{i}.b", blk.block[ib], f[2], pair=pr)
            if with_dgrad:
                dgr(f"res{i}.a", blk.block[ia])
                dgr(f"res{i}.b", blk.block[ib])
        fwd("up2", g.upsample2[1], 2 * f[2], pair=pu)
        fwd("up1", g.upsample1[1], f[4] + f[1], pair=pu)
        # conv11 input order = [out(f4), conv0(f0), x(cin)] — identical to the reference cat (:230), zero padded
        fwd("conv11", g.conv11[0], f[4] + f[0] + cp, pair=self.pair11())
        if g.append_smoothers:
            fwd("smooth0", g.smoothers[0], f[5], blk=self.SMOOTH_BLK, pair=ps)
            fwd("smooth3", g.smoothers[3], f[5], blk=self.SMOOTH_BLK, pair=ps)
        if with_dgrad:
            dgr("down1", g.downsample1[0], s2d=True)
            dgr("down2", g.downsample2[0], s2d=True)
            dgr("up2", g.upsample2[1])
            dgr("up1", g.upsample1[1])
            dgr("conv11", g.conv11[0], keep=f[4] + f[0], pair=self.pair11_dgrad())
            if g.append_smoothers:
                dgr("smooth0", g.smoothers[0])
                dgr("smooth3", g.smoothers[3])
            if self.input_grad:
                # dL/dx has two sources: the x slot of conv11's input (all of cat11's channels, rows padded to a multiple
                # of 32) and the initial conv
                dgr("conv11x", g.conv11[0], n_out=self.cat11x_channels())
                dgr("initial", g.initial_conv[0])
        W: Dict[str, Any] = dict(pk.out)
        if pkd is not None:
            W.update(pkd.out)

        def f32(t):
            return None if t is None else (t.detach() if t.dtype == torch.float32 else t.detach().float())

        # fp32 parameters are used in place (views); a .half()-ed module gets fp32 copies refreshed with the packer key
        W["head_w"] = f32(g.output[0].weight).reshape(3, f[5])
        W["head_b"] = f32(g.output[0].bias)
        W["b11"] = f32(g.conv11[0].bias)
        W["bs0"], W["bs3"] = (f32(g.smoothers[0].bias), f32(g.smoothers[3].bias)) if g.append_smoothers else (None, None)
        return {"packer": pk, "packer_bwd": pkd if (pkd is not None and pkd.jobs) else None, "W": W, "key": None,
                "ptrs": tuple(p.data_ptr() for p in g.parameters())}

    # -------------------------------------------------------------- forward
    def forward(self, x: Tensor, save: bool, u8_hwc: bool = False) -> Tensor:
        """x: NCHW float (module contract) or, with u8_hwc, uint8 frames [N,H,W,Cin] normalised on load"""
        g, dt, dev = self.gen, self.dt, self.device
        f = g.filters
        if u8_hwc:
            n, h, w, _ = x.shape
        else:
            n, _, h, w = x.shape
        h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
        train_bn = g.training
        bn_mode, no_norm = self.norm_mode() == "batch", self.norm_mode() == "none"
        ws = self.workspace(n, h, w, save)
        W = self._weights(with_dgrad=save)
        cp = self.cin_p
        pr, pu, ps = self._pairs(save)
        x = x.contiguous()
        if not u8_hwc and x.dtype not in (torch.float32, torch.float16):
            x = x.float()

        def conv_in(name, xin, cout, k, pad, raw, T_pref, up=False, pre=None, pre_st=None, pre_act=ACT_NONE, cps=0, pair=False,
                    up_raw=0, defer=False):
            """conv (bias skipped: a constant per channel is removed by the following InstanceNorm) + IN statistics;
            up=True: xin is the low-res tensor, the conv consumes its bilinear x2 upsample (interpolated in-kernel);
            pre: raw output of the previous conv, its InstanceNorm (pre_st) + activation applied on load"""
            src = xin if xin is not None else pre
            oh, ow = (2 * src.h, 2 * src.w) if up else (src.h, src.w)
            T = self._T(T_pref, ow)
            # (training passes only: inference keeps a frame's result bit-independent of the frames that share its pass -
            # batch tiles change the tile geometry of the statistics partial sums)
            tp = name == "initial" and self.tap_pairs()
            bt = 0 if (up or pre is not None or pair or not save or tp) else self._bt(n, oh, ow, cout)
            if bt:
                T, cps = bt, 0
            st = ws.stat(name, cout, oh, ow, 1 if bt else T, dev)
            cin = (xin.c if xin is not None else 0) + (pre.c if pre is not None else 0)
            frozen = bn_mode and not train_bn      # eval-mode BatchNorm: running statistics, no reduction
            plain_bias = None
            if no_norm:                            # conv -> activation: the bias counts, the "norm" is the identity table
                cb = self.norm_modules()[name][0].bias
                plain_bias = None if cb is None else cb.detach().float()
            ops.conv_fwd(xin, W[name], cout, k, k, pad, pad, dt, blk_c=self._blk(cin), tiles_per_cta=T, out=raw,
                         stats_partial=None if (frozen or no_norm) else st["partial"], upsample2x=up, pre=pre,
                         pre_scale=None if pre_st is None else pre_st["scale"], pre_shift=None if pre_st is None else pre_st["shift"],
                         up_raw_channels=up_raw, debug_flags=32 if ((up and self.ws_up) or (name.startswith("res") and self.ws_res)) else 0,
                         pre_act=pre_act, ctas_per_sm=0 if pair else cps, bias=plain_bias, cta_pair=pair, batch_tiles=bool(bt),
                         tap_pairs=tp)
            if no_norm:
                if not st.get("identity"):
                    st["scale"].fill_(1.0)
                    st["shift"].zero_()
                    st["identity"] = True
                return st
            if not bn_mode:
                if defer and save and st["tiles"] <= 64 and self.fuse_finalize:
                    # the consuming norm_apply computes scale / shift from the partial sums itself (and stores them for the
                    # backward pass): one launch less per normalised layer on the critical path of the training step
                    st["deferred"] = dict(partial=st["partial"], tiles=st["tiles"], count=oh * ow, eps=EPS)
                    return st
                st.pop("deferred", None)
                ops.norm_finalize(st["partial"], n, st["tiles"], cout, oh * ow, st["scale"], st["shift"], eps=EPS)
                return st
            # norm_layer='batch_norm' (reference :83-87): the same scale/shift tables, filled from batch statistics pooled
            # over the images (train) or from the running statistics (eval).  The conv bias is still skipped by the kernel:
            # batch statistics remove it; the running mean and the eval-mode shift account for it on the host.
            conv, bn = self.norm_modules()[name]
            gamma, beta = bn.weight.detach().float(), bn.bias.detach().float()
            bias = None if conv.bias is None else conv.bias.detach().float()
            if frozen:
                sc = gamma * torch.rsqrt(bn.running_var.float() + bn.eps)
                sh = beta - bn.running_mean.float() * sc + (0 if bias is None else bias * sc)
                st["scale"].copy_(sc.expand(n, cout))
                st["shift"].copy_(sh.expand(n, cout))
                return st
            if "mean" not in st:
                st["mean"] = torch.empty((cout,), device=dev)
                st["rstd"] = torch.empty((cout,), device=dev)
            mom = bn.momentum if bn.momentum is not None else 0.1
            ops.norm_finalize(st["partial"], n, st["tiles"], cout, oh * ow, st["scale"], st["shift"], eps=bn.eps,
                              batch_mode=True, gamma=gamma, beta=beta, running_mean=bn.running_mean,
                              running_var=bn.running_var, momentum=mom, mean_out=st["mean"], rstd_out=st["rstd"])
            if bias is not None:
                bn.running_mean.add_(bias, alpha=mom)      # the module's statistics are those of conv output + bias
            bn.num_batches_tracked += 1
            return st

        def napply(x, st, **kw):
            """norm_apply with the layer's statistics; a deferred finalize (see conv_in) is computed inside this launch"""
            ops.norm_apply(x, dt, scale=st["scale"], shift=st["shift"], **(st.pop("deferred", None) or {}), **kw)

        # input -> tail channels of cat11 (pad channels zeroed every call)
        xin = ws.cat11.view(f[4] + f[0], cp)
        if u8_hwc:
            ops.u8hwc_to_p8(x, xin, dt)   # ToTensor + Normalize(0.5, 0.5) fused into the layout conversion
        else:
            ops.nchw_to_p8(x, xin, dt)
        # encoder
        st = conv_in("initial", xin, f[0], 7, 3, ws.raw0, 3, cps=4, defer=True)
        napply(ws.raw0, st, act=ACT_LEAKY, out=ws.cat11.view(f[4], f[0]),
                       out_s2d=ws.s2d0)
        st = conv_in("down1", ws.s2d0, f[1], 2, 1, ws.raw1, 2, cps=4, defer=True)
        napply(ws.raw1, st, act=ACT_LEAKY, out=ws.c1cat.view(f[4], f[1]),
                       out_s2d=ws.s2d1)
        st = conv_in("down2", ws.s2d1, f[2], 2, 1, ws.raw2, 2, cps=4, defer=True)
        nb = len(g.resnet_blocks)
        r16 = ws.r16 is not None                # inference: 16-bit residual stream; r_0 is the skip slot of c2cat itself
        if r16:
            r_cur, r_nxt = ws.c2cat.view(f[2], f[2]), ws.r16[0]
        else:
            r_cur, r_nxt = ws.r[0], ws.r[1]
        a_of = (lambda i: ws.a[i]) if save else (lambda i: ws.a[i % 2])
        last16 = ws.c2cat.view(0, f[2])
        if nb == 0:
            napply(ws.raw2, st, act=ACT_LEAKY, out=ws.c2cat.view(f[2], f[2]))
            napply(ws.raw2, st, act=ACT_LEAKY, out=last16)
        else:
            napply(ws.raw2, st, act=ACT_LEAKY, out=ws.c2cat.view(f[2], f[2]),
                           out32=None if r16 else r_cur, out_relu=a_of(0))
        # residual blocks: r_{b+1} = r_b + IN(convB(relu(IN(convA(relu(r_b))))))
        # inference: IN + ReLU of a conv's raw output are applied inside the NEXT conv's shared-memory tile
        # (normalise-on-load), so hmid and the up1 half of cat11 are never written; training keeps them for wgrad
        nol_res = not save and f[2] % self._blk(f[2]) == 0 and f[2] <= 256
        nol_11 = not save and f[4] % 32 == 0 and f[4] <= 256 and g.append_smoothers
        fused_trunk = (save and self.fused_trunk and 1 <= nb <= 16 and not (bn_mode or no_norm or pr or r16)
                       and ops.res_trunk_supported(f[2], h4, w4))
        if fused_trunk:
            sts = [[ws.stat(f"res{b}.{ab}", f[2], h4, w4, 2, dev) for b in range(nb)] for ab in "ab"]
            for st_ in sts[0] + sts[1]:
                st_.pop("deferred", None)
            ops.res_trunk_fwd(ws.a[:nb], ws.rawA[:nb], ws.hmid[:nb], ws.rawB[:nb], [W[f"res{b}.a"] for b in range(nb)],
                              [W[f"res{b}.b"] for b in range(nb)], [(s_["scale"], s_["shift"]) for s_ in sts[0]],
                              [(s_["scale"], s_["shift"]) for s_ in sts[1]], r_cur, last16, dt, eps=EPS)
        for b in range(0 if fused_trunk else nb):
            k = b if save else 0
            st = conv_in(f"res{b}.a", a_of(b), f[2], 3, 1, ws.rawA[k], 2, cps=4, pair=pr, defer=not nol_res)
            if nol_res:
                st = conv_in(f"res{b}.b", None, f[2], 3, 1, ws.rawB[k], 2, pre=ws.rawA[k], pre_st=st, pre_act=ACT_RELU, pair=pr)
            else:
                napply(ws.rawA[k], st, act=ACT_RELU, out=ws.hmid[k])
                st = conv_in(f"res{b}.b", ws.hmid[k], f[2], 3, 1, ws.rawB[k], 2, cps=4, pair=pr, defer=True)
            lastb = b == nb - 1
            if r16:
                napply(ws.rawB[k], st, act=ACT_NONE, residual16=r_cur,
                               out=last16 if lastb else r_nxt, out_relu=None if lastb else a_of(b + 1))
                r_cur, r_nxt = r_nxt, (ws.r16[1] if b == 0 else r_cur)      # never write into c2cat's skip slot
            else:
                napply(ws.rawB[k], st, act=ACT_NONE, residual32=r_cur,
                               out32=None if lastb else r_nxt, out=last16 if lastb else None,
                               out_relu=None if lastb else a_of(b + 1))
                r_cur, r_nxt = r_nxt, r_cur
        # decoder.  Inference: the bilinear x2 upsample is interpolated inside the conv kernel's producer
        # (upsample2x=True), the upsampled tensors are never written.  Training keeps them: wgrad reads them.
        if save:
            ops.upsample2x(ws.c2cat, ws.u2in, dt)
            st = conv_in("up2", ws.u2in, f[4], 3, 1, ws.rawU2, 2)
            # IN + ReLU of the up2 output are applied on load by the upsample kernel (never materialised at H/2)
            ops.upsample2x(ws.rawU2, ws.u1in.view(0, f[4]), dt, scale=st["scale"], shift=st["shift"], act=ACT_RELU)
            ops.upsample2x(ws.c1cat.view(f[4], f[1]), ws.u1in.view(f[4], f[1]), dt)
            st = conv_in("up1", ws.u1in, f[4], 3, 1, ws.rawU1, 2, defer=True)
        else:
            # the raw output of up2 goes straight into its slot of up1's input; up1 normalises + activates it inside the staged
            # low-res tile before interpolating (up_raw_channels), so that normalise pass never touches HBM either
            fold = self.fold_up and f[4] % self._blk(f[4] + f[1]) == 0 and f[4] <= 256
            st = conv_in("up2", ws.c2cat, f[4], 3, 1, ws.c1cat.view(0, f[4]) if fold else ws.rawU2, 2, up=True, pair=pu)
            if fold:
                st = conv_in("up1", ws.c1cat, f[4], 3, 1, ws.rawU1, 2, up=True, pair=pu, pre_st=st, pre_act=ACT_RELU, up_raw=f[4])
            else:
                ops.norm_apply(ws.rawU2, dt, scale=st["scale"], shift=st["shift"], act=ACT_RELU, out=ws.c1cat.view(0, f[4]))
                st = conv_in("up1", ws.c1cat, f[4], 3, 1, ws.rawU1, 2, up=True, pair=pu)
        if not nol_11:
            napply(ws.rawU1, st, act=ACT_RELU, out=ws.cat11.view(0, f[4]))
        # conv11 + smoothers + fused head
        ev = self.kernel_timer
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        y = torch.empty((n, 3, h, w), device=dev)  # fresh result tensor: callers may keep it across calls
        # append_smoothers=False (reference :233-237): the 1x1 head + tanh sits directly on conv11's ReLU output
        head = {} if g.append_smoothers else dict(head_w=W["head_w"], head_b=W["head_b"], head_out=y, head_tanh=g.use_tanh)
        if nol_11:
            ops.conv_fwd(ws.cat11.view(f[4], f[0] + cp), W["conv11"], f[5], 7, 7, 3, 3, dt, blk_c=32,
                         tiles_per_cta=self._T11(w), bias=W["b11"], act=ACT_RELU, out=ws.c11, pre=ws.rawU1,
                         pre_scale=st["scale"], pre_shift=st["shift"], pre_act=ACT_RELU, cta_pair=self.pair11())
        else:
            ops.conv_fwd(ws.cat11, W["conv11"], f[5], 7, 7, 3, 3, dt, blk_c=32, tiles_per_cta=self._T11(w), bias=W["b11"],
                         act=ACT_RELU, out=ws.c11, cta_pair=self.pair11(), **head)
        if ev is not None:
            e1.record()
            ev.append((e0, e1, n))
        if not g.append_smoothers:
            if save:
                if not train_bn:
                    raise RuntimeError("autograd through GeneratorJ in eval() mode is not supported by the native path")
                self._saved = (ws, W)
                self._saved_stamp += 1
            return y
        bn = g.smoothers[2]
        T3, sblk = self._T(self.SMOOTH_T, w), self.SMOOTH_BLK
        if ps:
            T3 = 2       # the small-footprint CTA-pair configuration is built for two tiles per CTA
        if train_bn:
            st = ws.stat("bn", f[5], h, w, T3, dev)
            ops.conv_fwd(ws.c11, W["smooth0"], f[5], 3, 3, 1, 1, dt, blk_c=sblk, tiles_per_cta=T3, bias=W["bs0"], act=ACT_RELU,
                         out=ws.s0, stats_partial=st["partial"], ctas_per_sm=4)
            if "mean" not in st:
                st["mean"] = torch.empty((f[5],), device=dev)
                st["rstd"] = torch.empty((f[5],), device=dev)
            ops.norm_finalize(st["partial"], n, st["tiles"], f[5], h * w, st["scale"], st["shift"], eps=bn.eps,
                              batch_mode=True, gamma=bn.weight.detach().float(), beta=bn.bias.detach().float(),
                              running_mean=bn.running_mean, running_var=bn.running_var,
                              momentum=bn.momentum if bn.momentum is not None else 0.1, mean_out=st["mean"],
                              rstd_out=st["rstd"])
            bn.num_batches_tracked += 1
            ops.norm_apply(ws.s0, dt, scale=st["scale"], shift=st["shift"], act=ACT_NONE, out=ws.s0n)
        else:
            # eval: fold BatchNorm's running statistics into the epilogue of smoothers.0 (after its ReLU)
            rstd = torch.rsqrt(bn.running_var.float() + bn.eps)
            sc = (bn.weight.detach().float() * rstd).contiguous()
            sh = (bn.bias.detach().float() - bn.running_mean.float() * sc).contiguous()
            ops.conv_fwd(ws.c11, W["smooth0"], f[5], 3, 3, 1, 1, dt, blk_c=sblk, tiles_per_cta=T3, bias=W["bs0"], act=ACT_RELU,
                         post_scale=sc, post_shift=sh, out=ws.s0n, ctas_per_sm=4, cta_pair=ps)
            if save:
                raise RuntimeError("autograd through GeneratorJ in eval() mode is not supported by the native path")
        ops.conv_fwd(ws.s0n, W["smooth3"], f[5], 3, 3, 1, 1, dt, blk_c=sblk, tiles_per_cta=T3, bias=W["bs3"], act=ACT_RELU,
                     out=ws.s3 if save else None, head_w=W["head_w"], head_b=W["head_b"], head_out=y, head_tanh=g.use_tanh,
                     ctas_per_sm=4, cta_pair=ps)
        if save:
            self._saved = (ws, W)
            self._saved_stamp += 1
        return y

    def grad_scale_adjust(self) -> Tensor:
        """device float[3]: multiplier of the fp16 gradient-scale target (overflow back-off), clean-sweep and overflow
        counters - maintained by pbt_grad_scale_feedback at the end of every backward sweep"""
        if getattr(self, "_gs_adjust", None) is None:
            self._gs_adjust = torch.tensor([1.0, 0.0, 0.0], device=self.device)
        return self._gs_adjust

    def grad_bucket(self):
        """flat gradient storage of this generator's parameters (shared with parallel.GradAllReduce when data parallel)"""
        params = list(self.gen.named_parameters())
        if self.bucket is None or len(self.bucket.params) != len(params) or \
                any(a is not b for a, (_, b) in zip(self.bucket.params, params)) or self.bucket.flat.device != params[0][1].device:
            from .parallel import GradBucket
            self.bucket = GradBucket(params)
        return self.bucket

    def side_streams(self):
        """streams of the backward sweep that carry the weight gradients next to the data-gradient chain.  Successive wgrad
        launches alternate between them: on patch-sized maps a wgrad is a partial wave of latency-bound CTAs (128->128 3x3 on
        80 x 20x20: 180 CTAs, 29 % SM throughput), so two of them side by side fill the GPU better (env PBT_SIDE_STREAMS)."""
        if getattr(self, "_sides", None) is None:
            n = max(1, int(os.environ.get("PBT_SIDE_STREAMS", "2")))
            self._sides = [torch.cuda.Stream(self.device) for _ in range(n)]
        return self._sides

    def zero_pool_floats(self, n: int) -> int:
        """upper bound (in floats) of all accumulate-into buffers of one backward sweep"""
        g = self.gen
        f, cp = g.filters, self.cin_p
        nb = len(g.resnet_blocks)
        wg = (49 * cp * f[0] + 4 * 4 * f[0] * f[1] + 4 * 4 * f[1] * f[2] + nb * 2 * 9 * f[2] * f[2] + 9 * 2 * f[2] * f[4]
              + 9 * (f[4] + f[1]) * f[4] + 49 * (f[4] + f[0] + cp) * f[5] + 2 * 9 * f[5] * f[5])
        sums = n * 2 * (f[0] + f[1] + f[2] * (1 + 2 * nb) + 2 * f[4]) + 16 * f[5] + 64
        return wg + sums + 64 * 64   # slack for the 64-float alignment of each carve

    # -------------------------------------------------------------- backward
    def backward(self, gy: Tensor, y: Tensor, want_input_grad: bool = False):
        """returns (gradients for `list(gen.parameters())` in order, dL/dx or None)"""
        from .generator_bwd import generator_backward
        return generator_backward(self, gy, y, want_input_grad)
