"""torch fp32 restatement of GeneratorJ with 16-bit ROUNDING at the points where the native path stores or
feeds 16-bit tensors (conv operands, conv outputs, upsample in/out) and an exact fp32 autograd backward.

Purpose: separates the two sources of gradient deviation from the fp32 reference
  (1) the operand precision of the forward pass (ReLU/LeakyReLU masks flip for pre-activations within
      rounding distance of zero; inherent to ANY 16-bit-operand implementation, including the reference's
      own `precision: 16` mode), from
  (2) errors of the hand-written backward kernels — which must agree with this emulation tightly.
"""
import torch
import torch.nn.functional as F


def emulated_loss_and_grads(sd, x, target, dtype=torch.float16, weight=4.0):
    class _R(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return t.to(dtype).float()

        @staticmethod
        def backward(ctx, g):
            return g

    R = _R.apply
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    p = dict(sd)
    p.update(leaves)

    def conv(t, k, s, pad, bias=True):
        b = p.get(k + ".bias") if bias else None
        return F.conv2d(R(t), R(p[k + ".weight"]), b, stride=s, padding=pad)

    def inorm(t):
        t = R(t)  # the native path stores the raw conv output in 16 bit and normalises that
        m = t.mean((2, 3), keepdim=True)
        v = t.var((2, 3), unbiased=False, keepdim=True)
        return (t - m) / torch.sqrt(v + 1e-5)

    def up(t):
        return R(F.interpolate(R(t), scale_factor=2, mode="bilinear", align_corners=True))

    c0 = F.leaky_relu(inorm(conv(x, "initial_conv.0", 1, 3, bias=False)), 0.2)
    c1 = F.leaky_relu(inorm(conv(c0, "downsample1.0", 2, 1, bias=False)), 0.2)
    c2 = F.leaky_relu(inorm(conv(c1, "downsample2.0", 2, 1, bias=False)), 0.2)
    out = c2
    nb = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("resnet_blocks."))
    for b in range(nb):
        t = inorm(conv(F.relu(out), f"resnet_blocks.{b}.block.1", 1, 1, bias=False))
        t = inorm(conv(F.relu(t), f"resnet_blocks.{b}.block.4", 1, 1, bias=False))
        out = out + t
    out = F.relu(inorm(conv(up(torch.cat([out, c2], 1)), "upsample2.1", 1, 1, bias=False)))
    out = F.relu(inorm(conv(up(torch.cat([out, c1], 1)), "upsample1.1", 1, 1, bias=False)))
    out = R(F.relu(conv(torch.cat([out, c0, x], 1), "conv11.0", 1, 3)))
    out = R(F.relu(conv(out, "smoothers.0", 1, 1)))
    m = out.mean((0, 2, 3))
    v = out.var((0, 2, 3), unbiased=False)
    out = (out - m.view(1, -1, 1, 1)) / torch.sqrt(v.view(1, -1, 1, 1) + 1e-5) * p["smoothers.2.weight"].view(1, -1, 1, 1) \
        + p["smoothers.2.bias"].view(1, -1, 1, 1)
    out = R(F.relu(conv(out, "smoothers.3", 1, 1)))
    y = torch.tanh(F.conv2d(out, p["output.0.weight"], p["output.0.bias"]))
    loss = (y - target).abs().mean() * weight
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return y.detach(), loss.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}
