"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous.

  * full-frame inference shards FRAMES: rank r of R owns the contiguous range [r*F/R, (r+1)*F/R); frames are
    independent (InstanceNorm is per frame, BatchNorm uses frozen statistics in eval) so there is NO collective
    on the data path.
  * patch training is synchronous data parallel: each rank samples its own batch, the generator gradients are
    summed with ONE exchange per step and divided by the world size — what Lightning's DDP does implicitly for
    the reference (train.py:93-94, fires inside manual_backward at lightning_model.py:241), before
    clip_grad_norm_.  Gradients are packed into a flat fp32 bucket in three groups (tail / decoder / trunk) in
    the order the backward sweep produces them; each group's all-reduce is launched asynchronously as soon as
    its last gradient lands, so NCCL runs on its own stream underneath the rest of the backward pass.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def dist_env() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly"""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_distributed(backend: str | None = None) -> Tuple[int, int, int]:
    rank, world, local = dist_env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [lo, hi) range of rank `rank`; sizes differ by at most one and cover 0..n_items exactly"""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# gradient groups in the order generator_backward produces them
def _group_of(name: str) -> int:
    if name.startswith(("output.", "smoothers.", "conv11.")):
        return 0
    if name.startswith(("upsample1.", "upsample2.")):
        return 1
    return 2


class GradAllReduce:
    """flat-bucket mean all-reduce of a parameter set's gradients, overlapped group by group"""

    def __init__(self, named_params: Sequence[Tuple[str, torch.nn.Parameter]], world: int | None = None):
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.names = [n for n, _ in named_params]
        self.params = [p for _, p in named_params]
        order = sorted(range(len(self.names)), key=lambda i: (_group_of(self.names[i]), i))
        self.slices: Dict[str, Tuple[int, int]] = {}
        self.group_bounds: List[Tuple[int, int]] = []
        off = 0
        cur_g, g_lo = None, 0
        for i in order:
            g = _group_of(self.names[i])
            if cur_g is None:
                cur_g = g
            if g != cur_g:
                self.group_bounds.append((g_lo, off))
                cur_g, g_lo = g, off
            n = self.params[i].numel()
            self.slices[self.names[i]] = (off, off + n)
            off += n
        self.group_bounds.append((g_lo, off))
        dev = self.params[0].device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self._pending: List = []
        self._filled = [0] * len(self.group_bounds)
        self._group_sizes = [hi - lo for lo, hi in self.group_bounds]
        self._group_idx = {n: next(k for k, (lo, hi) in enumerate(self.group_bounds) if lo <= self.slices[n][0] < hi)
                           for n in self.names}

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def grad_ready(self, name: str, grad: torch.Tensor) -> None:
        """called by the backward sweep the moment a parameter gradient exists"""
        lo, hi = self.slices[name]
        self.flat[lo:hi].copy_(grad.reshape(-1))
        k = self._group_idx[name]
        self._filled[k] += hi - lo
        if self._filled[k] == self._group_sizes[k] and self.world > 1:
            glo, ghi = self.group_bounds[k]
            self._pending.append(dist.all_reduce(self.flat[glo:ghi], op=dist.ReduceOp.SUM, async_op=True))

    def collect_from_params(self) -> None:
        """fallback entry: pack every p.grad after a finished backward and reduce in one go"""
        for n, p in zip(self.names, self.params):
            if self._filled[self._group_idx[n]] < self._group_sizes[self._group_idx[n]]:
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                self.grad_ready(n, g)

    def finish(self) -> None:
        """wait for the exchanges, average, and hand the reduced gradients back as views of the bucket"""
        for w in self._pending:
            w.wait()
        self._pending.clear()
        if self.world > 1:
            self.flat.div_(self.world)
        for n, p in zip(self.names, self.params):
            lo, hi = self.slices[n]
            p.grad = self.flat[lo:hi].view_as(p)
        self._filled = [0] * len(self.group_bounds)
