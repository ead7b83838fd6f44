"""Native path of the perceptual term of the generator loss: ``PerceptualVGG19`` feature taps + mean squared feature
difference (reference src/models/perception.py:93-143, consumed at lightning_model.py:270-275).

The reference runs the frozen VGG19 prefix twice per step (generated patches, target patches), flattens and concatenates
the tapped activations and takes ``((fake - target) ** 2).mean()``; autograd then walks back through the prefix to the
generator output.  Here

  * both passes are ONE batch of 2B images through the generator's implicit-GEMM conv kernel (`pbt_conv_fwd`, bias and
    ReLU in the epilogue; 16-bit operands, fp32 accumulation) and `pbt_maxpool2`;
  * the feature tensors are never flattened or concatenated: `pbt_feature_mse` adds each tap's squared difference to the
    loss (fixed summation order) and, in the same launch, forms the tap's gradient, adds the gradient arriving from the
    deeper layers and applies the ReLU mask;
  * the VGG weights are frozen, so the only gradient is the one w.r.t. the generated patches: it is computed by the same
    sweep (conv kernel with tap-flipped weights, `pbt_maxpool2_bwd`) before the loss is returned, and autograd's backward
    of this node is a single scaling of that tensor.

Tap semantics follow the reference exactly: torchvision's VGG uses in-place ReLUs and the reference taps a VIEW of the
running activation (perception.py:106-110), so a tapped conv whose ReLU is executed is rectified too.  Shipped
configuration ``feature_layers: [0, 3, 5]`` = relu(conv1_1), relu(conv1_2), conv2_1.  No CPU path.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import torch
from torch import Tensor

from . import ops
from ._native import ACT_NONE, ACT_RELU, BF16, FP16, P8


def _p16(c: int) -> int:
    return (c + 15) // 16 * 16


def _blk(cin: int) -> int:
    return 32 if cin % 32 == 0 or cin > 32 else 16


class _Node:
    """one tensor of the prefix: the output of a conv (+ its in-place ReLU) or of a 2x2 max pooling"""

    def __init__(self, kind: str, index: int, conv=None, relu: bool = False, taps: int = 0):
        self.kind, self.index, self.conv, self.relu, self.taps = kind, index, conv, relu, taps
        self.name = f"f{index}"
        if conv is not None:
            self.cin, self.cout = conv.in_channels, conv.out_channels
            self.cin_p, self.cout_p = _p16(self.cin), _p16(self.cout)


def plan(module) -> Union[List[_Node], str]:
    """the node list of `module`'s tapped prefix, or the reason the native path cannot run it"""
    feats = module.model.features
    layers = set(int(i) for i in module.feature_layers)
    if not layers or min(layers) < 0 or max(layers) >= len(feats):
        return "feature_layers outside the feature stack"
    last = max(layers)
    nodes: List[_Node] = []
    i = 0
    while i <= last:
        m = feats[i]
        if isinstance(m, torch.nn.Conv2d):
            if (m.kernel_size, m.stride, m.padding, m.dilation, m.groups) != ((3, 3), (1, 1), (1, 1), (1, 1), 1) or m.padding_mode != "zeros":
                return f"features[{i}] is not a 3x3 stride-1 pad-1 convolution"
            if _p16(m.out_channels) > 256 or _p16(m.in_channels) > 256:
                return f"features[{i}] is wider than 256 channels"
            relu = i + 1 <= last and isinstance(feats[i + 1], torch.nn.ReLU)
            if relu and not feats[i + 1].inplace and i in layers:
                return f"features[{i}] is tapped in front of an out-of-place ReLU"
            nodes.append(_Node("conv", i, m, relu, int(i in layers) + (int(i + 1 in layers) if relu else 0)))
            i += 2 if relu else 1
        elif isinstance(m, torch.nn.MaxPool2d):
            k = m.kernel_size if isinstance(m.kernel_size, tuple) else (m.kernel_size, m.kernel_size)
            s = m.stride if isinstance(m.stride, tuple) else (m.stride, m.stride)
            if k != (2, 2) or s != (2, 2) or m.padding not in (0, (0, 0)) or m.ceil_mode or m.dilation not in (1, (1, 1)):
                return f"features[{i}] is not a 2x2 stride-2 max pooling"
            if not nodes:
                return "the feature stack starts with a pooling layer"
            nodes.append(_Node("pool", i, taps=int(i in layers)))
            i += 1
        else:
            return f"features[{i}] ({type(m).__name__}) has no native kernel"
    if any(p.requires_grad for p in module.model.features.parameters()):
        return "the VGG weights are trainable (requires_grad=True)"
    return nodes


def supported(module, x: Tensor) -> Optional[str]:
    """None when the native path can evaluate `module` on `x`, else the reason"""
    if not x.is_cuda or x.dim() != 4:
        return "not a CUDA NCHW batch"
    key = (tuple(module.feature_layers), len(module.model.features))
    cached = getattr(module, "_native_plan", None)
    if cached is None or cached[0] != key:
        cached = (key, plan(module))
        object.__setattr__(module, "_native_plan", cached)
        object.__setattr__(module, "_native_engine", None)
    p = cached[1]
    if isinstance(p, str):
        if "trainable" in p and not any(q.requires_grad for q in module.model.features.parameters()):
            object.__setattr__(module, "_native_plan", None)      # frozen since the plan was made: plan again
            return supported(module, x)
        return p
    if any(q.requires_grad for q in module.model.features.parameters()):
        return "the VGG weights are trainable (requires_grad=True)"
    pools = sum(1 for nd in p if nd.kind == "pool")
    if x.shape[1] != p[0].cin:
        return "channel count does not match the first convolution"
    if module.use_normalization and x.shape[1] != 3:
        return "ImageNet normalisation needs 3 channels"
    if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        return "unsupported dtype"
    if x.shape[2] % (1 << pools) or x.shape[3] % (1 << pools) or min(x.shape[2], x.shape[3]) < (8 << pools):
        return f"patch size must be a multiple of {1 << pools} and at least {8 << pools}"
    return None


class PerceptualEngine:
    def __init__(self, module, operand_dtype: str = "fp16"):
        nodes = plan(module)
        if isinstance(nodes, str):
            raise RuntimeError("native perceptual loss: " + nodes)
        self.module, self.nodes = module, nodes
        self.dt = {"fp16": FP16, "bf16": BF16}[operand_dtype]
        self.device = next(module.model.features.parameters()).device
        #: the 16-bit gradient tensors hold dL/df * count / 2 * grad_scale (count = elements of the concatenated taps): the
        #: tap gradient is then (f - t) * grad_scale, O(1), instead of O(1e-8)
        self.grad_scale = 1.0
        self._ws: Dict[Any, Dict[str, Any]] = {}
        self._packer = None
        self._pack_key = None

    # ------------------------------------------------------------------ operands
    def _norm_affine(self):
        """(x + 1) / 2 - mean) / std as a per-channel scale and shift (reference perception.py:75-91)"""
        m = self.module
        a = (0.5 / m.std).reshape(-1).float()
        b = ((0.5 - m.mean) / m.std).reshape(-1).float()
        return a, b

    def _weights(self):
        convs = [nd.conv for nd in self.nodes if nd.kind == "conv"]
        params = [c.weight for c in convs] + [c.bias for c in convs if c.bias is not None]
        key = tuple((p.data_ptr(), p._version) for p in params) + (bool(self.module.use_normalization),)
        if self._packer is None or self._packer["ptrs"] != tuple(p.data_ptr() for p in params):
            pk = ops.WeightPacker(self.device)
            first = self.nodes[0]
            w0d = torch.empty_like(first.conv.weight, dtype=torch.float32).contiguous()
            for nd in self.nodes:
                if nd.kind != "conv":
                    continue
                w = nd.conv.weight.detach()
                pk.add(nd.name, w, k_pad=nd.cin_p, n_out=nd.cout_p, n_keep=nd.cout, blk_c=_blk(nd.cin_p), dt=self.dt)
                src = w0d if nd is first else w
                pk.add(nd.name + ".d", src, dgrad=True, k_pad=nd.cout_p, n_out=nd.cin_p, n_keep=nd.cin, blk_c=_blk(nd.cout_p), dt=self.dt)
            bias = {nd.name: torch.zeros(nd.cout_p, device=self.device) for nd in self.nodes if nd.kind == "conv"}
            c0 = first.cin_p
            self._packer = {"pk": pk, "ptrs": tuple(p.data_ptr() for p in params), "bias": bias, "w0d": w0d,
                            "a": torch.zeros(c0, device=self.device), "b": torch.zeros(c0, device=self.device)}
            self._pack_key = None
        if self._pack_key != key or torch.cuda.is_current_stream_capturing():
            pkd = self._packer
            first = self.nodes[0]
            w0 = first.conv.weight.detach().float()
            if self.module.use_normalization:
                a, b = self._norm_affine()
                pkd["a"][:first.cin].copy_(a)
                pkd["b"][:first.cin].copy_(b)
                torch.mul(w0, a.view(1, -1, 1, 1), out=pkd["w0d"])     # d(normalised input)/d(input) folded into conv 0's dgrad
            else:
                pkd["w0d"].copy_(w0)
            pkd["pk"].run()
            for nd in self.nodes:
                if nd.kind == "conv" and nd.conv.bias is not None:
                    pkd["bias"][nd.name][:nd.cout].copy_(nd.conv.bias.detach())
            self._pack_key = key
        return self._packer["pk"].out, self._packer["bias"]

    def _workspace(self, n: int, h: int, w: int) -> Dict[str, Any]:
        key = (n, h, w)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        if len(self._ws) >= 2:
            self._ws.clear()
        dev, dt = self.device, self.dt
        ws = {"n": n, "h": h, "w": w}
        c0 = self.nodes[0].cin_p
        ws["x"] = P8.empty(2 * n, c0, h, w, dt, device=dev, zero=True)
        ws["xn"] = P8.empty(2 * n, c0, h, w, dt, device=dev, zero=True)
        hh, ww, c, real = h, w, c0, 0
        count = 0
        for nd in self.nodes:
            if nd.kind == "conv":
                c, real = nd.cout_p, nd.cout           # padded / real channels of the running activation
            else:
                hh, ww = hh // 2, ww // 2
            ws[nd.name] = P8.empty(2 * n, c, hh, ww, dt, device=dev)
            ws[nd.name + ".g"] = P8.empty(n, c, hh, ww, dt, device=dev)
            count += nd.taps * real * hh * ww
        ws["gx"] = P8.empty(n, c0, h, w, dt, device=dev)
        ws["count"] = count * n                                    # elements of the concatenated feature matrix [n, sum]
        ws["partial"] = torch.empty(4096, device=dev)
        ws["counter"] = torch.zeros(1, dtype=torch.int32, device=dev)
        self._ws[key] = ws
        return ws

    def _conv(self, xin: P8, wpack, cout_p: int, **kw) -> None:
        bt = xin.n >= 2 and xin.h * xin.w <= 1600 and (cout_p + 31) // 32 * 32 * 2 <= 256
        ops.conv_fwd(xin, wpack, cout_p, 3, 3, 1, 1, self.dt, blk_c=_blk(xin.c), tiles_per_cta=2, ctas_per_sm=0 if bt else 4,
                     batch_tiles=bool(bt), **kw)

    # ------------------------------------------------------------------ value (+ gradient w.r.t. the generated patches)
    def loss_and_grad(self, y: Tensor, target: Tensor, want_grad: bool):
        """(mean squared feature difference as a 0-dim fp32 tensor, d loss / d y as fp32 NCHW or None)"""
        if y.shape != target.shape:
            raise ValueError(f"generated {tuple(y.shape)} and target {tuple(target.shape)} patches differ in shape")
        why = supported(self.module, y)
        if why is not None:
            raise RuntimeError("native perceptual loss: " + why)
        n, c, h, w = y.shape
        dt, dev = self.dt, self.device
        W, B = self._weights()
        ws = self._workspace(n, h, w)

        def as_src(t: Tensor) -> Tensor:
            t = t.detach().contiguous()
            return t if t.dtype in (torch.float32, torch.float16) else t.float()

        x = ws["x"]
        ops.nchw_to_p8(as_src(y), P8(x.t[:n]), dt)
        ops.nchw_to_p8(as_src(target), P8(x.t[n:]), dt)
        cur = x
        if self.module.use_normalization:
            ops.norm_apply(x, dt, scale=self._packer["a"], shift=self._packer["b"], per_channel=True, act=ACT_NONE, out=ws["xn"])
            cur = ws["xn"]
        for nd in self.nodes:
            out = ws[nd.name]
            if nd.kind == "conv":
                self._conv(cur, W[nd.name], nd.cout_p, bias=B[nd.name], act=ACT_RELU if nd.relu else ACT_NONE, out=out)
            else:
                ops.maxpool2(cur, out, dt)
            cur = out
        loss = torch.zeros((), device=dev)
        count = float(ws["count"])
        red = dict(partial=ws["partial"], counter=ws["counter"], loss=loss)
        if not want_grad:
            for nd in reversed(self.nodes):          # the order the gradient sweep adds the taps in: same rounding, same value
                if nd.taps:
                    ops.feature_mse(ws[nd.name], n, dt, tap=True, loss_mul=nd.taps / count, **red)
            return loss, None
        s = self.grad_scale
        for k in range(len(self.nodes) - 1, -1, -1):
            nd = self.nodes[k]
            g = ws[nd.name + ".g"]
            behind = k < len(self.nodes) - 1
            if nd.taps or nd.relu:
                ops.feature_mse(ws[nd.name], n, dt, g=g, grad_mul=s * nd.taps, accumulate=behind, relu=nd.relu, tap=nd.taps > 0,
                                loss_mul=nd.taps / count, **(red if nd.taps else {}))
            g_in = ws["gx"] if k == 0 else ws[self.nodes[k - 1].name + ".g"]
            if nd.kind == "conv":
                self._conv(g, W[nd.name + ".d"], nd.cin_p, out=g_in)
            else:
                ops.maxpool2_bwd(ws[self.nodes[k - 1].name], g, g_in, dt)
        gy = torch.empty((n, c, h, w), device=dev)
        ops.p8_to_nchw(ws["gx"], c, gy, dt, mul=2.0 / (count * s))
        return loss, gy


class _PerceptualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, target, eng):
        want = y.requires_grad
        loss, gy = eng.loss_and_grad(y, target, want)
        ctx.y_dtype = y.dtype
        if want:
            ctx.save_for_backward(gy)
        return loss

    @staticmethod
    def backward(ctx, gl):
        (gy,) = ctx.saved_tensors
        return (gy * gl).to(ctx.y_dtype), None, None


def vgg19_prefix(depth: int, seed: int = 5, bias_std: float = 0.05) -> torch.nn.Sequential:
    """``features[0:depth]`` of torchvision's VGG19 (configuration E: conv3x3 + in-place ReLU, 2x2 max pooling), weights drawn
    like torchvision's initialiser (kaiming-normal, fan-out) under `seed`, small non-zero biases.  For tests and benchmarks
    on boxes without the ImageNet checkpoint; training uses ``PerceptualVGG19(path=...)`` / the torchvision cache."""
    cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
    layers, cin = [], 3
    for v in cfg:
        if len(layers) >= depth:
            break
        if v == "M":
            layers.append(torch.nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [torch.nn.Conv2d(cin, v, 3, padding=1), torch.nn.ReLU(inplace=True)]
            cin = v
    seq = torch.nn.Sequential(*layers[:depth])
    g = torch.Generator().manual_seed(seed)
    for m in seq:
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data = torch.randn(m.weight.shape, generator=g) * (2.0 / (m.out_channels * 9)) ** 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=g) * bias_std
    return seq


def feature_mse(module, y: Tensor, target: Tensor) -> Tensor:
    """``((features(y) - features(target)) ** 2).mean()`` of a ``PerceptualVGG19`` on the native kernels, differentiable
    w.r.t. `y` (the target is a constant, as at reference lightning_model.py:273)"""
    eng = getattr(module, "_native_engine", None)
    if eng is None or eng.device != y.device:
        eng = PerceptualEngine(module)
        object.__setattr__(module, "_native_engine", eng)
    if torch.is_grad_enabled() and y.requires_grad:
        return _PerceptualFn.apply(y, target, eng)
    return eng.loss_and_grad(y, target, False)[0]
