"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the adversarial half of the reference training step, functional over state dicts:

  discriminator_forward  <- DiscriminatorN_IN.forward / _make_block    reference src/models/discriminator.py:105-155
  gan_train_step         <- StyleTransferModel.training_step           reference lightning_model.py:201-258
                            _discriminator_step                                                    :294-319
                            _generator_step (image + adversarial terms)                            :260-292
                            configure_optimizers (two Adams)                                       :323-341

Order of one step, as in the reference: (1) no-grad generator pass in train() mode (its BatchNorm running statistics
move), D loss = 0.5 * (crit(D(post), 1) + crit(D(G(x)), 0)), clip, Adam on D; (2) generator pass with gradients,
loss = w_rec * L1(G(x), post) + w_adv * crit(D(G(x)), 1) against the UPDATED critic, clip, Adam on G.

Pinning: tests/test_oracle.py checks this file against tests/golden/gan_step.npz, produced by
oracle/make_golden_gan.py from the UNMODIFIED reference GeneratorJ / DiscriminatorN_IN modules and torch.optim.Adam.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch
import torch.nn.functional as F

from .generator_oracle import AdamState, _inorm, clip_grad_norm, generator_forward

Tensor = torch.Tensor


def discriminator_forward(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """instance-norm PatchGAN: 4x4 convs, pad 1; stride 2 for `initial` and `intermediate.*`, stride 1 after;
    LeakyReLU(0.2) everywhere except the last conv; InstanceNorm everywhere except the first and last conv"""
    def conv(h, key, stride):
        return F.conv2d(h, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=1)
    h = F.leaky_relu(conv(x, "initial.0", 2), 0.2)
    i = 0
    while f"intermediate.{i}.0.weight" in sd:
        h = F.leaky_relu(_inorm(conv(h, f"intermediate.{i}.0", 2)), 0.2)
        i += 1
    h = F.leaky_relu(_inorm(conv(h, "pre_output.0", 1)), 0.2)
    return conv(h, "output.0", 1)


def _leaves(sd: Dict[str, Tensor]):
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    full = dict(sd)
    full.update(leaves)
    return names, leaves, full


def _bn_state(g_sd):
    return {k.split(".")[-1]: g_sd[k] for k in ("smoothers.2.running_mean", "smoothers.2.running_var",
                                                "smoothers.2.num_batches_tracked")}


def gan_train_step(g_sd: Dict[str, Tensor], d_sd: Dict[str, Tensor], opt_g: AdamState, opt_d: AdamState, x: Tensor,
                   post: Tensor, *, clip: float = 0.5, w_rec: float = 4.0, w_adv: float = 0.5,
                   crit: Callable[[Tensor, Tensor], Tensor] = F.mse_loss) -> Tuple[Dict[str, Tensor], Dict[str, Dict[str, Tensor]]]:
    """one full reference training step; updates both state dicts in place.  Returns (losses, un-clipped gradients)"""
    out: Dict[str, Tensor] = {}
    # ---- critic
    with torch.no_grad():
        fake = generator_forward(g_sd, x, training=True, bn_state=_bn_state(g_sd))
    names, leaves, full = _leaves(d_sd)
    real_l = discriminator_forward(full, post)
    fake_l = discriminator_forward(full, fake)
    out["d_real_loss"] = crit(real_l, torch.ones_like(real_l))
    out["d_fake_loss"] = crit(fake_l, torch.zeros_like(fake_l))
    out["d_total_loss"] = (out["d_real_loss"] + out["d_fake_loss"]) * 0.5
    gd = dict(zip(names, torch.autograd.grad(out["d_total_loss"], [leaves[k] for k in names])))
    raw_d = {k: v.clone() for k, v in gd.items()}
    clip_grad_norm(gd, clip)
    with torch.no_grad():
        opt_d.step(d_sd, gd)
    # ---- generator, against the updated critic
    names, leaves, full = _leaves(g_sd)
    y = generator_forward(full, x, training=True, bn_state=_bn_state(g_sd))
    out["margin_loss"] = (y - post).abs().mean() * w_rec
    labels = discriminator_forward(d_sd, y)
    out["g_adversarial_loss"] = crit(labels, torch.ones_like(labels)) * w_adv
    out["g_total_loss"] = out["margin_loss"] + out["g_adversarial_loss"]
    gl = torch.autograd.grad(out["g_total_loss"], [leaves[k] for k in names], allow_unused=True)
    gg = {k: (g if g is not None else torch.zeros_like(g_sd[k])) for k, g in zip(names, gl)}
    raw_g = {k: v.clone() for k, v in gg.items()}
    clip_grad_norm(gg, clip)
    with torch.no_grad():
        opt_g.step(g_sd, gg)
    return {k: v.detach() for k, v in out.items()}, {"d": raw_d, "g": raw_g}
