"""Drop-in for the reference's lightning_model.py: ``StyleTransferModel`` with the same constructor and hooks
(reference lightning_model.py:13-21,201-356), generator path on sm_100a kernels.

Scope: the generator half of ``training_step`` is the accelerated path — batched device
sampler -> native GeneratorJ forward/backward -> L1*reconstruction_weight -> (data-parallel mean all-reduce)
-> clip_grad_norm_ -> Adam.  The PatchGAN discriminator and the VGG19 perceptual branch are outside that path
(section 8f) and are not built here: configs that carry them load unchanged, the two loss terms are reported
as disabled once, and the step runs generator-only (the G-only step is also what BASELINE.md times).
Uses pytorch_lightning.LightningModule as base when it is importable, else a plain nn.Module driven by
pbt_b200.trainer.Trainer.
"""
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from pbt_b200.config import to_container
from src.data.dataset import StyleTransferDataset
from src.models.generator import GeneratorJ

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    pl = None
    _Base = nn.Module


class _IndexLoader:
    """DataLoader(shuffle=True) replacement: the same RandomSampler index stream (torch RNG), but every batch
    is cut by ONE gather launch on the device instead of per-item __getitem__ + collate + H2D."""

    def __init__(self, dataset: StyleTransferDataset, batch_size: int, rank: int = 0, world: int = 1):
        self.ds, self.bs, self.rank, self.world = dataset, batch_size, rank, world

    def __len__(self):
        return (len(self.ds) // self.world + self.bs - 1) // self.bs

    def __iter__(self):
        n = len(self.ds)
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        gen = torch.Generator()
        gen.manual_seed(seed)
        perm = torch.randperm(n, generator=gen)
        if self.world > 1:                      # DistributedSampler-style strided shard of the permutation
            perm = perm[self.rank::self.world]
        for i in range(0, len(perm), self.bs):
            yield self.ds.sample_batch(perm[i:i + self.bs].tolist())


class StyleTransferModel(_Base):
    def __init__(self, generator_config: Dict[str, Any], discriminator_config: Optional[Dict[str, Any]],
                 training_config: Dict[str, Any], optimizer_config: Dict[str, Any], data_config: Dict[str, Any],
                 perception_loss_config: Optional[Dict[str, Any]] = None):
        super().__init__()
        if pl is not None:
            self.automatic_optimization = False
        self.data_config = data_config
        self.additional_channels = dict(data_config.get("additional_channels", {}) or {})
        self.training_config = training_config
        self.optimizer_config = optimizer_config
        args = to_container(dict(generator_config.get("args", {})))
        if args.get("input_channels") in ("auto", None):
            # reference lightning_model.py:71-88,137-148: RGB + sum of the configured guide depths
            args["input_channels"] = 3 + sum(int(c.get("depth", 1)) if isinstance(c, dict) else 1
                                             for c in self.additional_channels.values()) \
                if args.get("input_channels") == "auto" else 3
            args["additional_channels"] = to_container(self.additional_channels)
        self.generator = GeneratorJ(**args)
        self.discriminator = None
        self.perception_loss_model = None
        if discriminator_config is not None or perception_loss_config:
            print("[StyleTransferModel] discriminator / perceptual branches are outside the B200 hot path "
                  "(SURVEY.md section 8f): running the generator-only step (L1 reconstruction loss)")
        self.reconstruction_criterion = getattr(nn, training_config["reconstruction_criterion"])()
        self.use_cuda_graph = bool(training_config.get("cuda_graph", True))
        self.grad_sync = None
        self._optimizers = None
        self._graphed = None

    # ------------------------------------------------------------------ Lightning-shaped hooks
    def configure_optimizers(self):
        oc = to_container(dict(self.optimizer_config["generator"]))
        oc["betas"] = tuple(oc.get("betas", (0.9, 0.999)))
        fused = bool(oc.pop("fused", True))
        if fused and next(self.generator.parameters()).is_cuda:
            # clip_grad_norm_ + Adam as two native launches over all 48 tensors (pbt_b200/optim.py); same state layout
            from pbt_b200.optim import FusedClipAdam
            return [FusedClipAdam(self.generator.parameters(), **oc)]
        if self.use_cuda_graph and next(self.generator.parameters()).is_cuda:
            oc["capturable"] = True   # the whole step is replayed as one CUDA graph (pbt_b200/graphs.py)
        return [torch.optim.Adam(self.generator.parameters(), **oc)]

    def setup(self, stage: Optional[str] = None):
        if stage in ("fit", None):
            self.train_dataset = StyleTransferDataset(**to_container(dict(self.data_config)))

    def train_dataloader(self):
        tr = getattr(self, "trainer", None)
        rank, world = (getattr(tr, "rank", 0), getattr(tr, "world", 1)) if tr is not None else (0, 1)
        return _IndexLoader(self.train_dataset, int(self.training_config["batch_size"]), rank, world)

    def optimizers(self):
        return self._optimizers

    def log_dict(self, d, **_):
        tr = getattr(self, "trainer", None)
        if tr is not None and hasattr(tr, "log"):
            for k, v in d.items():
                tr.log(k, v)

    def training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int):
        (opt_g,) = self.optimizers()[:1]
        if "combined_input" in batch:
            combined_input = batch["combined_input"]
        else:  # reference-shaped batch dict: concatenate pre + guides in config order (lightning_model.py:211-221)
            tensors = [batch["pre"]]
            for name in self.additional_channels:
                key = f"channel_{name}"
                if key not in batch:
                    raise ValueError(f"Channel {name} not found in batch")
                tensors.append(batch[key])
            combined_input = torch.cat(tensors, dim=1)
        opt_g.zero_grad(set_to_none=True)
        g_loss = self._generator_step(combined_input, batch)
        g_loss["loss"].backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()          # mean over ranks, before the clip (DDP semantics of the reference)
        clip = self.training_config["gradient_clip_val"] if self.training_config.get("use_gradient_clipping", False) else None
        if hasattr(opt_g, "last_grad_norm"):     # FusedClipAdam: the clip is part of the optimiser launch
            opt_g.step(max_grad_norm=clip)
        else:
            if clip is not None:
                torch.nn.utils.clip_grad_norm_(self.generator.parameters(), clip)
            opt_g.step()
        return g_loss

    def graphed_training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int):
        """the same generator step replayed as ONE CUDA graph (static shapes; falls back to training_step for the
        ragged last batch of an epoch)"""
        from pbt_b200.graphs import GraphedGeneratorStep
        x = batch["combined_input"]
        if self._graphed is None:
            self._graphed = GraphedGeneratorStep(
                self.generator, self.optimizers()[0], tuple(x.shape),
                reconstruction_weight=float(self.training_config["reconstruction_weight"]),
                clip=float(self.training_config["gradient_clip_val"]) if self.training_config.get("use_gradient_clipping", False) else None,
                criterion=self.reconstruction_criterion, grad_sync=self.grad_sync)
        if tuple(x.shape) != tuple(self._graphed.x.shape):
            return self.training_step(batch, batch_idx)
        loss = self._graphed(x, batch["post"])
        tr = getattr(self, "trainer", None)
        if tr is not None and getattr(tr, "global_step", 0) % max(1, int(self.training_config.get("log_every_n_steps", 10))) == 0:
            self.log_dict({"g_image_loss": float(loss), "g_total_loss": float(loss)})
        return {"loss": loss, "g_total_loss": loss}

    def _generator_step(self, combined_input, batch):
        generated = self.generator(combined_input)
        losses = {}
        if self.training_config["use_image_loss"]:
            losses["margin_loss"] = self.reconstruction_criterion(generated, batch["post"]) * \
                self.training_config["reconstruction_weight"]
        total = sum(losses.values())
        losses["g_total_loss"] = total
        tr = getattr(self, "trainer", None)
        if tr is not None and getattr(tr, "global_step", 0) % max(1, int(self.training_config.get("log_every_n_steps", 10))) == 0:
            self.log_dict({"g_image_loss": float(losses.get("margin_loss", 0.0)), "g_total_loss": float(total)})
        return {"loss": total, **losses}
