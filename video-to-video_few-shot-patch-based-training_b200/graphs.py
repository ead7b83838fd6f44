"""CUDA-graph capture of the whole training step: G-only (forward, loss, backward, clip, Adam) and the two-network
adversarial step (critic update, then generator update against the updated critic).

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


class _StateSnapshot:
    """Training state (module tensors and optimiser moments / step counters) saved before the warm-up passes of a
    capture and written back IN PLACE afterwards — captured graphs hold the addresses — so that the warm-up, which has
    to run real steps to populate lazy state and allocator pools, leaves no trace in the training trajectory."""

    def __init__(self, modules, optimizers):
        self.modules = [(m, {k: v.detach().clone() for k, v in m.state_dict().items()}) for m in modules]
        self.opts = [(o, {p: {k: v.detach().clone() for k, v in st.items() if torch.is_tensor(v)}
                          for p, st in o.state.items()}) for o in optimizers]

    def restore(self) -> None:
        with torch.no_grad():
            for m, saved in self.modules:
                for k, v in m.state_dict().items():
                    v.copy_(saved[k])
            for o, saved in self.opts:
                for p, st in o.state.items():
                    for k, v in st.items():
                        if torch.is_tensor(v):      # state created lazily during the warm-up restarts from zero
                            v.copy_(saved[p][k]) if p in saved and k in saved[p] else v.zero_()


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
        snap = _StateSnapshot([generator], [optimizer])
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
        snap.restore()                     # the warm-up steps (on the zero buffers) are not part of the training run

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


class GraphedGanStep:
    """The reference's full training_step (lightning_model.py:224-250) as ONE CUDA graph: no-grad generator pass +
    critic forward/backward + clip + Adam, then generator forward, reconstruction / adversarial / perceptual losses,
    backward through the critic into the native generator sweep, clip + Adam.  `model` is the StyleTransferModel whose
    ``full_step`` (host-synchronisation free) is captured; both optimisers must be capturable.  Calls return the SAME
    dict of static device scalars every time (overwritten by the next replay)."""

    def __init__(self, model, batch_shape, warmup: int = 3):
        gen = model.generator
        dev = next(gen.parameters()).device
        n, cin, ph, pw = batch_shape
        self.model = model
        for opt in model.optimizers():
            for grp in opt.param_groups:
                if not grp.get("capturable", False):
                    raise ValueError("GraphedGanStep needs optimizers created with capturable=True")
        self.x = torch.zeros((n, cin, ph, pw), device=dev)
        self.target = torch.zeros((n, 3, ph, pw), device=dev)
        self.out = None
        # captured lazily by the first call, so that the warm-up passes see a real batch (the critic's InstanceNorm on
        # an all-zero batch would be degenerate)
        self._warmup, self.graph = warmup, None

    def _capture(self):
        dev = self.x.device
        mods = [m for m in (self.model.generator, self.model.discriminator) if m is not None]
        snap = _StateSnapshot(mods, self.model.optimizers())
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self.model.full_step(self.x, self.target)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = self.model.full_step(self.x, self.target)
            self.out = {k: v.detach().clone() for k, v in out.items()}
        snap.restore()

    def __call__(self, x: torch.Tensor, target: torch.Tensor):
        self.x.copy_(x, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        return self.out
