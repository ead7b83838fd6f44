"""A ~150-line stand-in for pytorch_lightning.Trainer (not installed in this image) that drives the
reference-shaped ``StyleTransferModel``: setup -> train_dataloader -> training_step, manual optimisation,
`{'state_dict': ...}` checkpoints with the `generator.` key prefix the inference driver expects
(reference generator.py:115-118,180), top-k on g_total_loss + last.ckpt (reference train.py:22-31), early
stopping on the epoch mean (train.py:39-46).  With WORLD_SIZE > 1 (torchrun, one process per GPU) rank 0's initial
weights are broadcast to every rank (what DDP does when Lightning wraps the reference model, train.py:93-94), the
generator gradients are mean all-reduced over NCCL before clipping (parallel.GradAllReduce), and every decision that
changes control flow (epoch loss -> top-k / early stop) is taken on a value averaged over the ranks, so that no rank
can leave the loop while another waits in a collective.  Checkpoints carry the optimiser states (`optimizer_states`,
Lightning's key) and `fit(ckpt_path=...)` resumes from one."""
from __future__ import annotations

import os
import time
from typing import Any, Dict, List, Optional

import torch

from .parallel import GradAllReduce, broadcast_module_state, init_distributed


class Trainer:
    def __init__(self, max_epochs: int = 50, max_steps: Optional[int] = None, output_dir: str = "outputs",
                 log_every_n_steps: int = 10, save_top_k: int = 3, early_stopping_patience: Optional[int] = None,
                 steps_per_epoch: Optional[int] = None, devices: int = 1, **_ignored: Any):
        self.max_epochs, self.max_steps = max_epochs, max_steps
        self.ckpt_dir = os.path.join(output_dir, "checkpoints")
        self.log_every = log_every_n_steps
        self.save_top_k = save_top_k
        self.patience = early_stopping_patience
        self.steps_per_epoch = steps_per_epoch
        self.rank, self.world, self.local = init_distributed()
        if devices and devices > 1 and self.world == 1:
            print(f"[trainer] devices={devices} requested: launch with `torchrun --nproc-per-node {devices} train.py ...`; "
                  "running on one GPU")
        self.global_step = 0
        self.logged: Dict[str, float] = {}
        self._best: List[tuple] = []

    # Lightning-style logging sink used by StyleTransferModel.log / log_dict
    def log(self, name: str, value) -> None:
        self.logged[name] = float(value)

    def save_checkpoint(self, model, path: str) -> None:
        if self.rank == 0:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            torch.save({"state_dict": model.state_dict(), "global_step": self.global_step, "epoch": self.epoch,
                        "optimizer_states": [o.state_dict() for o in (model.optimizers() or [])]}, path)

    def _resume(self, model, ckpt_path: str) -> None:
        ckpt = torch.load(ckpt_path, map_location=torch.device("cuda", self.local))
        model.load_state_dict(ckpt["state_dict"], strict=True)
        for opt, st in zip(model.optimizers(), ckpt.get("optimizer_states", [])):
            opt.load_state_dict(st)
        self.global_step = int(ckpt.get("global_step", 0))
        self.epoch = int(ckpt.get("epoch", -1)) + 1

    def _mean_over_ranks(self, value: float) -> float:
        if self.world == 1:
            return value
        import torch.distributed as dist
        t = torch.tensor([value], dtype=torch.float64, device=torch.device("cuda", self.local) if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t)
        return float(t.item()) / self.world

    def fit(self, model, ckpt_path: Optional[str] = None) -> None:
        model.trainer = self
        model.to(torch.device("cuda", self.local))
        gen = model.generator
        if self.world > 1:
            # identical replicas from the first step on: rank 0's parameters and buffers (reference: DDP's initial broadcast)
            broadcast_module_state(gen)
            if getattr(model, "discriminator", None) is not None:
                broadcast_module_state(model.discriminator)
        model.setup("fit")
        opts = model.configure_optimizers()
        model._optimizers = opts
        self.epoch = 0
        if ckpt_path:
            self._resume(model, ckpt_path)
        if self.world > 1:
            model.grad_sync = GradAllReduce(list(gen.named_parameters()), world=self.world).attach(gen)
            if getattr(model, "discriminator", None) is not None:     # the critic's 98 kB of gradients: one exchange per step
                model.d_grad_sync = GradAllReduce(list(model.discriminator.named_parameters()), world=self.world)
        loader = model.train_dataloader()
        bad_epochs, best_epoch_loss = 0, float("inf")
        model.train()
        for epoch in range(self.epoch, self.max_epochs):
            self.epoch = epoch
            t0, losses = time.time(), []
            for batch_idx, batch in enumerate(loader):
                graphed = getattr(model, "use_cuda_graph", False) and hasattr(model, "graphed_training_step")
                out = model.graphed_training_step(batch, batch_idx) if graphed else model.training_step(batch, batch_idx)
                self.global_step += 1
                losses.append(float(out["loss"]) if self.global_step % self.log_every == 0 or batch_idx == 0 else None)
                if self.rank == 0 and self.global_step % self.log_every == 0:
                    msg = " ".join(f"{k}={v:.4f}" for k, v in sorted(self.logged.items()))
                    print(f"[epoch {epoch} step {self.global_step}] {msg}", flush=True)
                if (self.max_steps and self.global_step >= self.max_steps) or \
                        (self.steps_per_epoch and batch_idx + 1 >= self.steps_per_epoch):
                    break
            vals = [v for v in losses if v is not None]
            # one value for every rank: top-k, early stopping and the loop exit must not diverge between ranks
            epoch_loss = self._mean_over_ranks(sum(vals) / max(1, len(vals)))
            if self.rank == 0:
                print(f"[epoch {epoch}] g_total_loss={epoch_loss:.4f} ({time.time() - t0:.1f}s)", flush=True)
                name = os.path.join(self.ckpt_dir, f"style_transfer-epoch={epoch:02d}-g_total_loss={epoch_loss:.4f}.ckpt")
                self.save_checkpoint(model, name)
                self.save_checkpoint(model, os.path.join(self.ckpt_dir, "last.ckpt"))
                self._best = sorted(self._best + [(epoch_loss, name)])
                for _, stale in self._best[self.save_top_k:]:
                    if os.path.exists(stale):
                        os.remove(stale)
                self._best = self._best[:self.save_top_k]
            if epoch_loss < best_epoch_loss - 1e-12:
                best_epoch_loss, bad_epochs = epoch_loss, 0
            else:
                bad_epochs += 1
            if self.patience is not None and bad_epochs >= self.patience:   # Lightning: wait_count >= patience
                if self.rank == 0:
                    print(f"[trainer] early stop: g_total_loss did not improve for {bad_epochs} epochs")
                break
            if self.max_steps and self.global_step >= self.max_steps:
                break
