"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the reference generator path, written functionally over a ``state_dict`` so it
shares no code with either the reference module tree or the CUDA implementation:

  generator_forward    <- GeneratorJ.forward                      reference src/models/generator.py:210-239
  generator_forward_bn <- the same with norm_layer='batch_norm'                                   :83-87
  generator_forward_plain <- the same with a norm_layer string that selects no norm layers        :83-87
  conv blocks          <- _make_conv_block / ResNetBlock / _make_upconv_block        :18-58,156-208
  g_only_train_step    <- StyleTransferModel.training_step (generator half) + _generator_step with
                          discriminator/perception disabled        reference lightning_model.py:239-250,260-292
                          and configure_optimizers                                     :323-329
  frame_to_uint8       <- StyleTransferInference.process_image post-processing   reference generator.py:643-647

Pinning: tests/test_oracle.py checks this file against tests/golden/*.npz, which were produced by importing the
UNMODIFIED reference modules in the build container (oracle/make_golden.py).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _inorm(x: Tensor, eps: float = 1e-5) -> Tensor:
    # nn.InstanceNorm2d(affine=False, track_running_stats=False): per-(n,c) biased variance, always instance stats
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def _conv(x: Tensor, sd: Dict[str, Tensor], key: str, stride: int, pad: int) -> Tensor:
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=pad)


def _up2(x: Tensor) -> Tensor:
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def _bnorm(x: Tensor, sd: Dict[str, Tensor], key: str, training: bool, running: Optional[Dict[str, Tensor]]) -> Tensor:
    """nn.BatchNorm2d(affine=True, eps 1e-5, momentum 0.1) at module path `key`; `running` (optional, train mode) maps
    '<key>.running_mean' / '.running_var' / '.num_batches_tracked' to tensors that are advanced in place"""
    g, b = sd[key + ".weight"].view(1, -1, 1, 1), sd[key + ".bias"].view(1, -1, 1, 1)
    if training:
        mean, var = x.mean(dim=(0, 2, 3)), x.var(dim=(0, 2, 3), unbiased=False)
        if running is not None:
            cnt = x.numel() / x.shape[1]
            with torch.no_grad():
                running[key + ".running_mean"].mul_(0.9).add_(0.1 * mean.detach())
                running[key + ".running_var"].mul_(0.9).add_(0.1 * var.detach() * cnt / max(cnt - 1, 1))
                running[key + ".num_batches_tracked"] += 1
    else:
        mean, var = sd[key + ".running_mean"], sd[key + ".running_var"]
    return (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-5) * g + b


def generator_forward_bn(sd: Dict[str, Tensor], x: Tensor, *, training: bool = False,
                         running: Optional[Dict[str, Tensor]] = None, tanh: bool = True) -> Tensor:
    """GeneratorJ(norm_layer='batch_norm') (reference src/models/generator.py:83-87 selects nn.BatchNorm2d for every norm
    slot): same graph as generator_forward with BatchNorm2d at `<block>.1` (conv blocks), `block.2` / `block.5` (residual
    blocks) and `<up block>.2`"""
    n_blocks = 1 + max((int(k.split(".")[1]) for k in sd if k.startswith("resnet_blocks.")), default=-1)
    bn = lambda t, key: _bnorm(t, sd, key, training, running)  # noqa: E731
    conv0 = F.leaky_relu(bn(_conv(x, sd, "initial_conv.0", 1, 3), "initial_conv.1"), 0.2)
    conv1 = F.leaky_relu(bn(_conv(conv0, sd, "downsample1.0", 2, 1), "downsample1.1"), 0.2)
    conv2 = F.leaky_relu(bn(_conv(conv1, sd, "downsample2.0", 2, 1), "downsample2.1"), 0.2)
    out = conv2
    for b in range(n_blocks):
        t = bn(_conv(F.relu(out), sd, f"resnet_blocks.{b}.block.1", 1, 1), f"resnet_blocks.{b}.block.2")
        t = bn(_conv(F.relu(t), sd, f"resnet_blocks.{b}.block.4", 1, 1), f"resnet_blocks.{b}.block.5")
        out = out + t
    out = F.relu(bn(_conv(_up2(torch.cat([out, conv2], 1)), sd, "upsample2.1", 1, 1), "upsample2.2"))
    out = F.relu(bn(_conv(_up2(torch.cat([out, conv1], 1)), sd, "upsample1.1", 1, 1), "upsample1.2"))
    out = F.relu(_conv(torch.cat([out, conv0, x], 1), sd, "conv11.0", 1, 3))
    out = F.relu(_conv(out, sd, "smoothers.0", 1, 1))
    out = F.relu(_conv(bn(out, "smoothers.2"), sd, "smoothers.3", 1, 1))
    out = _conv(out, sd, "output.0", 1, 0)
    return torch.tanh(out) if tanh else out


def generator_forward_plain(sd: Dict[str, Tensor], x: Tensor, *, training: bool = False,
                            running: Optional[Dict[str, Tensor]] = None, tanh: bool = True) -> Tensor:
    """GeneratorJ with a norm_layer string that selects no norm (reference src/models/generator.py:83-87 leaves norm = None):
    conv -> activation everywhere, the residual convs sit at block.1 / block.3 (:38-52); only smoothers.2 normalises"""
    n_blocks = 1 + max((int(k.split(".")[1]) for k in sd if k.startswith("resnet_blocks.")), default=-1)
    conv0 = F.leaky_relu(_conv(x, sd, "initial_conv.0", 1, 3), 0.2)
    conv1 = F.leaky_relu(_conv(conv0, sd, "downsample1.0", 2, 1), 0.2)
    conv2 = F.leaky_relu(_conv(conv1, sd, "downsample2.0", 2, 1), 0.2)
    out = conv2
    for b in range(n_blocks):
        t = _conv(F.relu(out), sd, f"resnet_blocks.{b}.block.1", 1, 1)
        out = out + _conv(F.relu(t), sd, f"resnet_blocks.{b}.block.3", 1, 1)
    out = F.relu(_conv(_up2(torch.cat([out, conv2], 1)), sd, "upsample2.1", 1, 1))
    out = F.relu(_conv(_up2(torch.cat([out, conv1], 1)), sd, "upsample1.1", 1, 1))
    out = F.relu(_conv(torch.cat([out, conv0, x], 1), sd, "conv11.0", 1, 3))
    out = F.relu(_conv(out, sd, "smoothers.0", 1, 1))
    out = F.relu(_conv(_bnorm(out, sd, "smoothers.2", training, running), sd, "smoothers.3", 1, 1))
    out = _conv(out, sd, "output.0", 1, 0)
    return torch.tanh(out) if tanh else out


def generator_forward(sd: Dict[str, Tensor], x: Tensor, *, training: bool = False, n_blocks: Optional[int] = None,
                      bn_state: Optional[Dict[str, Tensor]] = None, tanh: bool = True) -> Tensor:
    """x [N,Cin,H,W] fp32 -> [N,3,H,W].  `training` selects batch statistics for the single BatchNorm
    (smoothers.2); when `bn_state` is given its running_mean/var/num_batches_tracked are updated like
    nn.BatchNorm2d(momentum=0.1) does."""
    if n_blocks is None:
        n_blocks = 1 + max((int(k.split(".")[1]) for k in sd if k.startswith("resnet_blocks.")), default=-1)
    conv0 = F.leaky_relu(_inorm(_conv(x, sd, "initial_conv.0", 1, 3)), 0.2)
    conv1 = F.leaky_relu(_inorm(_conv(conv0, sd, "downsample1.0", 2, 1)), 0.2)
    conv2 = F.leaky_relu(_inorm(_conv(conv1, sd, "downsample2.0", 2, 1)), 0.2)
    out = conv2
    for b in range(n_blocks):
        t = _inorm(_conv(F.relu(out), sd, f"resnet_blocks.{b}.block.1", 1, 1))
        t = _inorm(_conv(F.relu(t), sd, f"resnet_blocks.{b}.block.4", 1, 1))
        out = out + t
    out = F.relu(_inorm(_conv(_up2(torch.cat([out, conv2], 1)), sd, "upsample2.1", 1, 1)))
    out = F.relu(_inorm(_conv(_up2(torch.cat([out, conv1], 1)), sd, "upsample1.1", 1, 1)))
    out = F.relu(_conv(torch.cat([out, conv0, x], 1), sd, "conv11.0", 1, 3))
    if "smoothers.0.weight" in sd:
        out = F.relu(_conv(out, sd, "smoothers.0", 1, 1))
        g, b = sd["smoothers.2.weight"], sd["smoothers.2.bias"]
        if training:
            mean = out.mean(dim=(0, 2, 3))
            var = out.var(dim=(0, 2, 3), unbiased=False)
            if bn_state is not None:
                cnt = out.numel() / out.shape[1]
                with torch.no_grad():
                    bn_state["running_mean"].mul_(0.9).add_(0.1 * mean.detach())
                    bn_state["running_var"].mul_(0.9).add_(0.1 * var.detach() * cnt / max(cnt - 1, 1))
                    bn_state["num_batches_tracked"] += 1
        else:
            mean, var = sd["smoothers.2.running_mean"], sd["smoothers.2.running_var"]
        out = (out - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-5) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        out = F.relu(_conv(out, sd, "smoothers.3", 1, 1))
    out = _conv(out, sd, "output.0", 1, 0)
    return torch.tanh(out) if tanh else out


def loss_and_grads(sd: Dict[str, Tensor], x: Tensor, target: Tensor, weight: float = 4.0
                   ) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """one-step parameter gradients of  L1(G(x), target) * weight  in train() mode (reference
    lightning_model.py:267-268 with reconstruction_weight 4.0). Returns (output, loss, grads)."""
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    y = generator_forward(full, x, training=True)
    loss = (y - target).abs().mean() * weight
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return y.detach(), loss.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}


class AdamState:
    """torch.optim.Adam(lr, betas, eps=1e-8, weight_decay) restated (L2 decay added to the gradient)."""

    def __init__(self, names, lr=4e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.t = 0
        self.m = {k: None for k in names}
        self.v = {k: None for k in names}

    def step(self, sd: Dict[str, Tensor], grads: Dict[str, Tensor]) -> None:
        self.t += 1
        b1, b2 = self.betas
        for k, g in grads.items():
            p = sd[k]
            g = g + self.wd * p
            if self.m[k] is None:
                self.m[k], self.v[k] = torch.zeros_like(p), torch.zeros_like(p)
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
            denom = (self.v[k].sqrt() / (bc2 ** 0.5)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def clip_grad_norm(grads: Dict[str, Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ (reference lightning_model.py:245-248)"""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads.values():
        g.mul_(coef)
    return total


def g_only_train_step(sd: Dict[str, Tensor], opt: AdamState, x: Tensor, target: Tensor, clip: float = 0.5,
                      weight: float = 4.0) -> Tensor:
    """generator half of training_step with the GAN / VGG branches off; updates `sd` in place, returns the loss"""
    bn_state = {k.split(".")[-1]: sd[k] for k in ("smoothers.2.running_mean", "smoothers.2.running_var",
                                                  "smoothers.2.num_batches_tracked")}
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    y = generator_forward(full, x, training=True, bn_state=bn_state)
    loss = (y - target).abs().mean() * weight
    gl = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, gl)}
    clip_grad_norm(grads, clip)
    with torch.no_grad():
        opt.step(sd, grads)
    return loss.detach()


def frame_to_uint8(y: Tensor) -> Tensor:
    """[N,3,H,W] in [-1,1] -> uint8 [N,H,W,3]   (reference generator.py:643-647)"""
    v = ((y.float().clamp(-1, 1) + 1) * 127.5).clamp(0, 255)
    return v.permute(0, 2, 3, 1).round().to(torch.uint8)
