"""``DiscriminatorN_IN`` under the reference's import path (reference src/models/discriminator.py:8-155): the PatchGAN
critic of the adversarial branch of ``training_step`` (reference lightning_model.py:224-236,279-283,294-319).

Same constructor, same ``state_dict`` keys and shapes (``initial.0.*``, ``intermediate.<i>.0.*``, ``pre_output.0.*``,
``output.0.*``), same ``forward(x) -> (logits, None)``, same N(0, 0.02) / zero-bias initialisation in the same module
registration order, so a torch seed reproduces the reference init and reference checkpoints load with strict=True.

On a CUDA device the arithmetic runs on the library's own kernels (pbt_b200/critic.py: the generator's tcgen05 implicit-GEMM
conv, wgrad and norm kernels over zero-padded channels; 4x4 stride-2 stages as 3x3 convs over a space-to-depth copy) - the
nn.Conv2d / norm sub-modules are parameter containers there.  The tensor-library expression below is what runs for CPU
tensors (host-side tests, the gloo data-parallel tests) and for configurations the native path does not cover
(`norm_layer` other than instance_norm, layers wider than 256 channels); `native = False` forces it.
"""
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

_NORMS = {"batch_norm": nn.BatchNorm2d, "instance_norm": nn.InstanceNorm2d}


def _stage(cin: int, cout: int, stride: int, bias: bool, norm, leaky: bool) -> nn.Sequential:
    """4x4 conv, padding 1 -> optional norm -> optional LeakyReLU(0.2) (reference discriminator.py:105-133)"""
    mods = [nn.Conv2d(cin, cout, kernel_size=4, stride=stride, padding=1, bias=bias)]
    if norm is not None:
        mods.append(norm(cout))
    if leaky:
        mods.append(nn.LeakyReLU(0.2, inplace=True))
    return nn.Sequential(*mods)


class DiscriminatorN_IN(nn.Module):
    #: run CUDA inputs through the native engine (pbt_b200/critic.py) when the configuration is supported
    native = True

    def __init__(self, input_channels: int = 3, additional_channels: Optional[Dict[str, Any]] = None, num_filters: int = 64,
                 n_layers: int = 3, use_noise: bool = False, noise_sigma: float = 0.2, norm_layer: str = "instance_norm",
                 use_bias: bool = True):
        super().__init__()
        self.use_noise, self.noise_sigma = bool(use_noise), float(noise_sigma)
        norm = _NORMS.get(norm_layer)
        cap = num_filters * 8
        # stride-2 pyramid: the first stage has no norm, widths double up to 8 * num_filters
        self.initial = _stage(input_channels, num_filters, 2, use_bias, None, True)
        self.intermediate = nn.ModuleList()
        width = num_filters
        for _ in range(1, n_layers):
            nxt = min(2 * width, cap)
            self.intermediate.append(_stage(width, nxt, 2, use_bias, norm, True))
            width = nxt
        # two stride-1 stages: each 4x4 / pad 1 conv shrinks the map by one pixel
        nxt = min(2 * width, cap)
        self.pre_output = _stage(width, nxt, 1, use_bias, norm, True)
        self.output = _stage(nxt, 1, 1, use_bias, None, False)
        self._engine = None
        self._native_reason = None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight.data, 0.0, 0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias.data, 0.0)

    def forward(self, x: Tensor) -> Tuple[Tensor, None]:
        if self.use_noise and self.training:
            x = x + torch.randn_like(x) * self.noise_sigma
        if x.is_cuda and self.native and self._native_reason is None:
            from pbt_b200 import critic
            if self._engine is None:
                self._native_reason = critic.supported(self)
                if self._native_reason is None:
                    self._engine = critic.CriticEngine(self)
            if self._engine is not None:
                return critic.critic_forward(self, self._engine, x), None
        h = self.initial(x)
        for stage in self.intermediate:
            h = stage(h)
        return self.output(self.pre_output(h)), None
