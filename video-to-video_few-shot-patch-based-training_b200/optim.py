"""Fused optimiser tail of the G-only training step.

``FusedClipAdam`` is a drop-in for the pair the reference runs after ``manual_backward``:
``torch.nn.utils.clip_grad_norm_(G.parameters(), gradient_clip_val)`` (lightning_model.py:245-248) followed by
``torch.optim.Adam(G.parameters(), lr, betas, eps, weight_decay)`` (lightning_model.py:326-329,
config/optimizer/default.yaml:2-10).  Both run in two native launches over a device table of the 48 parameter tensors
(``pbt_clip_adam_step``) instead of ~50 foreach / reduction launches; the step counter lives on the device, so the tail
is CUDA-graph capturable.  State is exposed in ``torch.optim.Adam``'s layout (``state[p] = {step, exp_avg,
exp_avg_sq}``), so checkpoints stay interchangeable with the reference's optimizer state.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch

from . import _native as nv
from ._native import check, lib, stream_ptr


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None, **unused):
        for k in ("amsgrad", "maximize"):
            if unused.get(k):
                raise NotImplementedError(f"FusedClipAdam: {k} is not supported")
        defaults = dict(lr=float(lr), betas=tuple(float(b) for b in betas), eps=float(eps), weight_decay=float(weight_decay),
                        capturable=True)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError("FusedClipAdam: one parameter group (the generator) is supported")
        self.max_grad_norm = max_grad_norm
        ps = [p for p in self.param_groups[0]["params"]]
        if not ps or not all(p.is_cuda and p.dtype == torch.float32 for p in ps):
            raise RuntimeError("FusedClipAdam needs fp32 CUDA parameters (no CPU path)")
        dev = ps[0].device
        self._dev = dev
        # device state: [0] sum of squares scratch, [1] step count; total norm of the last step
        # [squared norm, step count, skipped (non-finite) steps, per-block partial sums of the norm (32 per tensor)]
        self._state = torch.zeros(3 + 32 * len(ps), device=dev)
        self.last_grad_norm = torch.zeros((), device=dev)
        total = sum(p.numel() for p in ps)
        self._m = torch.zeros(total, device=dev)
        self._v = torch.zeros(total, device=dev)
        off = 0
        for p in ps:
            n = p.numel()
            self.state[p] = {"step": self._state[1], "exp_avg": self._m[off:off + n].view_as(p),
                             "exp_avg_sq": self._v[off:off + n].view_as(p)}
            off += n
        self._max = max(p.numel() for p in ps)
        self._tables = {}   # pointer set -> (pinned host table, device table, built during graph capture)

    @property
    def skipped_steps(self) -> int:
        """steps skipped because the gradients were not finite (fp16 overflow in the backward sweep)"""
        return int(self._state[2].item())

    # torch.optim.Optimizer.load_state_dict re-creates the per-parameter tensors: copy them back into the flat buffers
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        off = 0
        step = None
        for p in self.param_groups[0]["params"]:
            n = p.numel()
            st = self.state[p]
            self._m[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self._v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            step = float(st["step"]) if step is None else step
            st["exp_avg"], st["exp_avg_sq"], st["step"] = self._m[off:off + n].view_as(p), self._v[off:off + n].view_as(p), self._state[1]
            off += n
        if step is not None:
            self._state[1] = step

    def _job_table(self) -> torch.Tensor:
        """device table of (param, grad, m, v, n) per tensor for the current pointer set.  Tables are cached per pointer
        set together with the pinned host copy they were uploaded from: a table built while a CUDA graph is being
        captured is re-uploaded from that host buffer by EVERY replay, so captured entries are never evicted (an evicted
        pinned block would be recycled by the caching host allocator and later replays would upload garbage pointers)."""
        ps = self.param_groups[0]["params"]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        hit = self._tables.get(key)
        if hit is not None:
            return hit[1]
        jobs = (nv.OptimJob * len(ps))()
        for i, p in enumerate(ps):
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                raise RuntimeError("FusedClipAdam: gradients and parameters must be contiguous fp32")
            st = self.state[p]
            jobs[i] = nv.OptimJob(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
        host = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).pin_memory()
        capturing = torch.cuda.is_current_stream_capturing()
        table = host.to(self._dev, non_blocking=True)
        if not capturing:
            # eager pointer sets come and go (e.g. the ragged last batch of an epoch): keep the newest few; the upload
            # above is ordered before any later reuse of the pinned block only once it has completed
            torch.cuda.current_stream(self._dev).synchronize()
            stale = [k for k, v in self._tables.items() if not v[2]]
            for k in stale[:-7]:
                del self._tables[k]
        self._tables[key] = (host, table, capturing)
        return table

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = "default"):
        """clip (when max_grad_norm is set) + Adam.  Every parameter must have a gradient."""
        if closure is not None:
            raise NotImplementedError("FusedClipAdam: closures are not supported")
        grp = self.param_groups[0]
        ps = grp["params"]
        if any(p.grad is None for p in ps):
            raise RuntimeError("FusedClipAdam: every parameter needs a gradient (the native backward produces all of them)")
        clip = self.max_grad_norm if max_grad_norm == "default" else max_grad_norm
        table = self._job_table()
        b1, b2 = grp["betas"]
        check(lib().pbt_clip_adam_step(table.data_ptr(), len(ps), self._max, self._state.data_ptr(),
                                       float(clip) if clip else 0.0, float(grp["lr"]), float(b1), float(b2),
                                       float(grp["eps"]), float(grp["weight_decay"]), self.last_grad_norm.data_ptr(),
                                       stream_ptr()), "pbt_clip_adam_step")
        # the kernel wrote the parameters behind autograd's back: bump their version counters so that consumers keyed on
        # them (the generator's packed 16-bit weight cache) see the update
        torch.autograd.graph.increment_version(ps)


class _FusedL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, target, weight):
        if not (y.is_cuda and y.dtype == torch.float32 and target.dtype == torch.float32 and y.shape == target.shape):
            raise RuntimeError("fused_l1_loss needs fp32 CUDA tensors of equal shape")
        y, target = y.contiguous(), target.contiguous()
        loss = torch.empty((), device=y.device)
        gy = torch.empty_like(y)
        check(lib().pbt_l1_loss_fwd_bwd(y.data_ptr(), target.data_ptr(), y.numel(), float(weight), loss.data_ptr(), gy.data_ptr(),
                                        stream_ptr()), "pbt_l1_loss_fwd_bwd")
        ctx.save_for_backward(gy)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (gy,) = ctx.saved_tensors
        return gy * grad_out, None, None


def fused_l1_loss(y: torch.Tensor, target: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
    """``torch.nn.functional.l1_loss(y, target) * weight`` (reference lightning_model.py:267-268) with the value and the
    gradient w.r.t. ``y`` produced by one native launch"""
    return _FusedL1.apply(y, target, weight)
