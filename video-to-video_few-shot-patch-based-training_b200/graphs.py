"""CUDA-graph capture of the whole G-only training step (forward, loss, backward, clip, Adam).

The step issues ~160 native launches plus a few dozen small tensor-library kernels; at patch-training sizes the
GPU work per launch is a few microseconds, so the host would be the bottleneck (measured: 10 ms/step of pure
launch overhead at 40 x 32x32 patches).  Capturing the step once and replaying it removes that overhead.
Everything the step touches is static: parameter storage, the engine's workspace, and the two input buffers
that the sampler's gather kernel fills (reference lightning_model.py:239-250,260-292).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .optim import FusedClipAdam, fused_l1_loss


class GraphedGeneratorStep:
    def __init__(self, generator, optimizer: torch.optim.Optimizer, batch_shape, *, reconstruction_weight: float = 4.0,
                 clip: Optional[float] = 0.5, criterion: Optional[Callable] = None, grad_sync=None, warmup: int = 3):
        """batch_shape = (N, Cin, P, P).  `optimizer` must be created with capturable=True."""
        dev = next(generator.parameters()).device
        n, cin, ph, pw = batch_shape
        self.gen, self.opt, self.clip, self.weight = generator, optimizer, clip, reconstruction_weight
        self.criterion = criterion or torch.nn.functional.l1_loss
        self._l1 = criterion is None or criterion is torch.nn.functional.l1_loss or (
            isinstance(criterion, torch.nn.L1Loss) and criterion.reduction == "mean")
        self.grad_sync = grad_sync
        self.x = torch.zeros((n, cin, ph, pw), device=dev)
        self.target = torch.zeros((n, 3, ph, pw), device=dev)
        self.loss = torch.zeros((), device=dev)
        for grp in optimizer.param_groups:
            if not grp.get("capturable", False):
                raise ValueError("GraphedGeneratorStep needs an optimizer created with capturable=True")
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):        # allocator warm-up, lazy state (Adam moments), kernel attributes
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()

    def _step(self):
        self.opt.zero_grad(set_to_none=True)
        y = self.gen(self.x)
        if self._l1:      # value + gradient of L1 * weight in one native launch
            loss = fused_l1_loss(y, self.target, self.weight)
        else:
            loss = self.criterion(y, self.target) * self.weight
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        if isinstance(self.opt, FusedClipAdam):       # clip + Adam in two native launches (pbt_clip_adam_step)
            self.opt.step(max_grad_norm=self.clip)
        else:
            if self.clip is not None:
                torch.nn.utils.clip_grad_norm_(self.gen.parameters(), self.clip)
            self.opt.step()
        self.loss.copy_(loss.detach())

    def __call__(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.graph.replay()
        return self.loss
