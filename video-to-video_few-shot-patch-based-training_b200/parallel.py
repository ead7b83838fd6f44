"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous.

  * full-frame inference shards FRAMES: rank r of R owns the contiguous range [r*F/R, (r+1)*F/R); frames are
    independent (InstanceNorm is per frame, BatchNorm uses frozen statistics in eval) so there is NO collective
    on the data path.
  * patch training is synchronous data parallel: each rank samples its own batch, the generator gradients are
    averaged with one exchange per gradient group and step — what Lightning's DDP does implicitly for the reference
    (train.py:93-94, fires inside manual_backward at lightning_model.py:241), before clip_grad_norm_.
    The backward sweep writes every parameter gradient straight into a flat fp32 bucket (`GradBucket`; `p.grad` are
    views of it, so there is no pack / unpack copy), ordered in the groups the sweep finishes them: tail, decoder, one
    group per residual block, encoder.  A group's all-reduce is enqueued the moment its last gradient is written, on the
    stream that wrote it, so NCCL runs underneath the rest of the backward pass and only the small encoder group
    (0.4 MB) is exposed at the end.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

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


def broadcast_module_state(module: torch.nn.Module, src: int = 0) -> None:
    """every parameter and buffer of `module` takes rank `src`'s value — what DistributedDataParallel does when the
    reference's Lightning trainer wraps the model (train.py:93-94).  Without it each rank would start from its own
    random initialisation and the averaged gradients would be applied to different weights."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)


def replicas_identical(module: torch.nn.Module, buffers: bool = True) -> bool:
    """True when every rank holds bit-identical parameters (and buffers): the raw bytes of the whole state (13 MB for the
    generator) are all-gathered and compared, so this is exact, not a tolerance.  buffers=False leaves out BatchNorm
    running statistics, which are per-rank by design in data-parallel training (batch statistics are not synchronised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return True
    tensors = list(module.parameters()) + (list(module.buffers()) if buffers else [])
    mine = torch.cat([t.detach().contiguous().reshape(-1).view(torch.uint8) for t in tensors])
    every = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(every, mine)
    return all(torch.equal(every[0], e) for e in every[1:])


# ---- gradient groups in the order generator_backward finishes them
_RES = re.compile(r"resnet_blocks\.(\d+)\.")


#: how finely the bucket is split into all-reduce groups: "block" = tail | decoder | one group per residual block | encoder
#: (10 groups), "pairs" = residual blocks two by two (6-7 groups), "coarse" = tail | decoder | trunk + encoder (3 groups, round 1),
#: "single" = one exchange at the end of the sweep.  Measured at N = 2 by tools/allreduce_groups.py.
GROUPING = os.environ.get("PBT_AR_GROUPS", "block")


def group_of(name: str, n_blocks: int = 0, grouping: str | None = None) -> int:
    """0 tail (head, smoothers, conv11) | 1 decoder | 2.. residual blocks, last block first | encoder last"""
    grouping = grouping or GROUPING
    if grouping == "single":
        return 0
    if name.startswith(("output.", "smoothers.", "conv11.")):
        return 0
    if name.startswith(("upsample1.", "upsample2.")):
        return 1
    if grouping == "coarse":
        return 2
    m = _RES.match(name)
    if m:
        k = n_blocks - 1 - int(m.group(1))
        return 2 + (k // 2 if grouping == "pairs" else k)
    return 2 + n_blocks


class GradBucket:
    """flat fp32 storage for the gradients of a parameter set, laid out group by group"""

    def __init__(self, named_params: Sequence[Tuple[str, torch.nn.Parameter]], grouping: str | None = None):
        self.names = [n for n, _ in named_params]
        self.params = [p for _, p in named_params]
        nb = 1 + max([int(m.group(1)) for m in map(_RES.match, self.names) if m] or [-1])
        grp = [group_of(n, nb, grouping) for n in self.names]
        order = sorted(range(len(self.names)), key=lambda i: (grp[i], i))
        self.slices: Dict[str, Tuple[int, int]] = {}
        self.group_of_name: Dict[str, int] = {}
        bounds: Dict[int, List[int]] = {}
        off = 0
        for i in order:
            n = self.params[i].numel()
            self.slices[self.names[i]] = (off, off + n)
            self.group_of_name[self.names[i]] = grp[i]
            b = bounds.setdefault(grp[i], [off, off])
            b[1] = off + n
            off += (n + 3) // 4 * 4           # 16-byte aligned slices (vector loads of the optimiser kernels)
        gids = sorted(bounds)
        self._gpos = {g: k for k, g in enumerate(gids)}
        self.group_bounds: List[Tuple[int, int]] = [tuple(bounds[g]) for g in gids]
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.params[0].device)
        self.views: Dict[str, torch.Tensor] = {n: self.flat[lo:hi].view_as(p) for (n, p), (lo, hi) in
                                               zip(named_params, (self.slices[n] for n in self.names))}

    def group_index(self, name: str) -> int:
        return self._gpos[self.group_of_name[name]]

    def aliased_by_param_grads(self) -> bool:
        """a parameter's .grad still lives in this bucket (no zero_grad since the last sweep): writing the next sweep into
        it would corrupt gradient accumulation"""
        return any(p.grad is not None and p.grad.data_ptr() == self.views[n].data_ptr() for n, p in zip(self.names, self.params))


class GradAllReduce:
    """mean all-reduce of a parameter set's gradients over a GradBucket, overlapped group by group"""

    def __init__(self, named_params: Sequence[Tuple[str, torch.nn.Parameter]], world: int | None = None,
                 bucket: Optional[GradBucket] = None, grouping: str | None = None):
        named_params = list(named_params)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.bucket = bucket if bucket is not None else GradBucket(named_params, grouping)
        self.names, self.params = self.bucket.names, self.bucket.params
        self.flat, self.slices, self.group_bounds = self.bucket.flat, self.bucket.slices, self.bucket.group_bounds
        self._pending: List = []
        self._filled = [0] * len(self.group_bounds)
        self._events: List[List] = [[] for _ in self.group_bounds]
        self._group_sizes = [sum(hi - lo for n, (lo, hi) in self.slices.items() if self.bucket.group_index(n) == k)
                             for k in range(len(self.group_bounds))]
        # NCCL averages inside the collective; gloo (CPU tests) sums and the division is a separate pass
        self._avg = dist.is_initialized() and dist.get_backend() == "nccl" and hasattr(dist.ReduceOp, "AVG")

    @property
    def nbytes(self) -> int:
        return sum(self._group_sizes) * 4

    def attach(self, generator) -> "GradAllReduce":
        """make the native backward sweep of `generator` write into this bucket and report each finished gradient"""
        from .generator import _Engine
        if generator._engine is None:
            generator._engine = _Engine(generator)
        generator._engine.bucket = self.bucket
        generator._engine.grad_hook = self.grad_ready
        return self

    def grad_ready(self, name: str, grad: Optional[torch.Tensor] = None) -> None:
        """called by the backward sweep the moment a parameter gradient exists (already inside the bucket when `grad`
        is the bucket's own view; copied in otherwise)"""
        lo, hi = self.slices[name]
        view = self.bucket.views[name]
        if grad is not None and grad.data_ptr() != view.data_ptr():
            view.copy_(grad.reshape(view.shape))
        k = self.bucket.group_index(name)
        self._filled[k] += hi - lo
        cuda = self.flat.is_cuda
        if self._filled[k] < self._group_sizes[k]:
            if cuda:        # the group's exchange must also wait for this stream's writes
                ev = torch.cuda.Event()
                ev.record()
                self._events[k].append((torch.cuda.current_stream(), ev))
            return
        if self.world > 1:
            if cuda:
                cur = torch.cuda.current_stream()
                for st, ev in self._events[k]:
                    if st != cur:
                        cur.wait_event(ev)
            glo, ghi = self.group_bounds[k]
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            self._pending.append(dist.all_reduce(self.flat[glo:ghi], op=op, async_op=True))
        self._events[k] = []

    def collect_from_params(self) -> None:
        """fallback entry: pack every p.grad after a finished backward and reduce"""
        for n, p in zip(self.names, self.params):
            k = self.bucket.group_index(n)
            if self._filled[k] < self._group_sizes[k]:
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                self.grad_ready(n, g)

    def finish(self) -> None:
        """wait for the exchanges, average, and hand the reduced gradients back as views of the bucket"""
        for w in self._pending:
            w.wait()
        self._pending.clear()
        if self.world > 1 and not self._avg:
            self.flat.div_(self.world)
        for n, p in zip(self.names, self.params):
            p.grad = self.bucket.views[n]
        self._filled = [0] * len(self.group_bounds)
        self._events = [[] for _ in self.group_bounds]
