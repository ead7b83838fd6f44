"""``PerceptualVGG19`` under the reference's import path (reference src/models/perception.py:9-143): frozen VGG19
feature taps for the perceptual term of the generator loss (reference lightning_model.py:270-275).

``forward`` / ``get_features`` return the flattened taps like the reference (tensor-library ops: a consumer of raw feature
matrices gets exactly the reference's tensors).  The training step does not need the matrices, only
``((features(generated) - features(target)) ** 2).mean()`` and its gradient: ``feature_mse`` computes that on the native
kernels (``pbt_b200/perceptual.py``: both passes as one batch through ``pbt_conv_fwd``, taps reduced in place) whenever the
module is on a CUDA device with frozen weights, and ``perceptual_loss`` routes through it.

The ImageNet weights cannot be downloaded on an offline box, so ``path=None`` raises a clear error unless torchvision finds
them in its local cache; ``path=<file>`` loads a custom checkpoint exactly like the reference does (8x8 classifier head with
``num_classes`` outputs); ``from_features`` wraps an already built feature stack.
"""
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor


class PerceptualVGG19(nn.Module):
    def __init__(self, feature_layers: List[int], use_normalization: bool = True, path: Optional[str] = None,
                 num_classes: int = 40, requires_grad: bool = False):
        super().__init__()
        from torchvision import models
        if path is None:
            # the reference downloads the ImageNet weights here (perception.py:50); never start a download from inside a
            # training job on a box that may have no route out - use torchvision's local cache or fail at once
            w = models.VGG19_Weights.IMAGENET1K_V1
            cached = os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(w.url))
            if not os.path.exists(cached) and os.environ.get("PBT_ALLOW_DOWNLOAD", "0") != "1":
                raise RuntimeError(f"PerceptualVGG19: the ImageNet VGG19 weights are not available offline ({cached} is "
                                   "missing); pass perception_model.args.path=<state_dict file>, set PBT_ALLOW_DOWNLOAD=1, "
                                   "or disable training.use_perception_loss")
            net = models.vgg19(weights=w)
        else:
            net = models.vgg19(weights=None)
            head = [nn.Linear(512 * 8 * 8, 4096), nn.ReLU(True), nn.Dropout(), nn.Linear(4096, 4096), nn.ReLU(True),
                    nn.Dropout(), nn.Linear(4096, num_classes)]
            net.classifier = nn.Sequential(*head)
            net.load_state_dict(torch.load(path, map_location="cpu"))
        self.model = net.float().eval()
        self.feature_layers = sorted(int(i) for i in feature_layers)
        self.use_normalization = bool(use_normalization)
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    @classmethod
    def from_features(cls, features: nn.Sequential, feature_layers: List[int], use_normalization: bool = True,
                      requires_grad: bool = False) -> "PerceptualVGG19":
        """the same module around an existing VGG ``features`` stack (weights already loaded by the caller)"""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        holder = nn.Module()
        holder.features = features
        self.model = holder.float().eval()
        self.feature_layers = sorted(int(i) for i in feature_layers)
        self.use_normalization = bool(use_normalization)
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False
        return self

    def normalize(self, x: Tensor) -> Tensor:
        """[-1, 1] images -> ImageNet-normalised (identity when use_normalization is off)"""
        return (0.5 * (x + 1.0) - self.mean) / self.std if self.use_normalization else x

    def get_features(self, x: Tensor) -> Tensor:
        """flattened taps after the listed ``features`` indices, concatenated per sample.  The taps are VIEWS of the
        running activation, as in the reference (perception.py:106-110): an in-place ReLU that follows a tapped conv
        therefore also rectifies the tap."""
        taps, h = [], x
        for i, layer in enumerate(self.model.features[: self.feature_layers[-1] + 1]):
            h = layer(h)
            if i in self.feature_layers:
                taps.append(h.view(h.size(0), -1))
        return torch.cat(taps, dim=1)

    def forward(self, x: Tensor) -> Tuple[None, Tensor]:
        return None, self.get_features(self.normalize(x))

    def native_unsupported(self, x: Tensor):
        """None when ``feature_mse`` runs on the native kernels for `x`, else the reason it takes the tensor-library route"""
        if not x.is_cuda:
            return "not on a CUDA device"
        from pbt_b200 import perceptual
        return perceptual.supported(self, x)

    def feature_mse(self, generated: Tensor, target: Tensor) -> Tensor:
        """``((features(generated) - features(target)) ** 2).mean()`` (reference lightning_model.py:272-274), the target a
        constant.  Native kernels on CUDA; configurations they do not cover (trainable VGG weights, taps deeper than the
        256-channel stages, odd patch sizes) evaluate the reference expression."""
        if self.native_unsupported(generated) is None:
            from pbt_b200 import perceptual
            return perceptual.feature_mse(self, generated, target.detach())
        return ((self(generated)[1] - self(target.detach())[1]) ** 2).mean()

    def perceptual_loss(self, y_pred: Tensor, y_true: Tensor) -> Tensor:
        if self.native_unsupported(y_pred) is None and not y_true.requires_grad:
            return self.feature_mse(y_pred, y_true)
        return F.mse_loss(self(y_pred)[1], self(y_true)[1])
