"""Drop-in for the reference's lightning_model.py: ``StyleTransferModel`` with the same constructor and hooks
(reference lightning_model.py:13-21,201-356), generator path on sm_100a kernels.

Scope: the generator half of ``training_step`` is the accelerated path — batched device
sampler -> native GeneratorJ forward/backward -> L1*reconstruction_weight -> (data-parallel mean all-reduce)
-> clip_grad_norm_ -> Adam; the G-only step is what BASELINE.md times.
The adversarial and perceptual branches (section 8f rank 4; reference lightning_model.py:224-236,270-283,294-319) are
available behind ``training.use_adversarial_loss`` (default on, as the reference trains) / ``training.use_perception_loss``
(default off: the ImageNet VGG19 weights cannot be downloaded here):
the PatchGAN critic (3->12->24->48->1 channels in the shipped config, 0.2 % of the generator's FLOPs) and the VGG taps
run on the tensor library, the generator passes inside them — the no-grad pass of the critic step, the forward and the
backward that receives dL/dy from both loss terms — run on the native kernels, and both Adam steps use the fused
clip+Adam launches.  The whole two-network step replays as one CUDA graph (pbt_b200/graphs.py::GraphedGanStep).
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
        return (self._per_rank() + self.bs - 1) // self.bs

    def _per_rank(self) -> int:
        return (len(self.ds) + self.world - 1) // self.world

    def __iter__(self):
        n = len(self.ds)
        seed = torch.empty((), dtype=torch.int64).random_()
        if self.world > 1:                      # one permutation for all ranks (DistributedSampler: shared seed + epoch)
            import torch.distributed as dist
            if dist.is_initialized():
                dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
                seed = seed.to(dev)
                dist.broadcast(seed, src=0)
        gen = torch.Generator()
        gen.manual_seed(int(seed.item()))
        perm = torch.randperm(n, generator=gen)
        if self.world > 1:
            # DistributedSampler semantics: pad the permutation by wrapping around so that it divides evenly, then take
            # the strided shard - every rank runs the same number of steps (unequal counts would leave a collective unmatched)
            total = self._per_rank() * self.world
            if total > n:
                perm = torch.cat([perm, perm[:total - n]])
            perm = perm[self.rank:total:self.world]
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
        self.perception_loss_weight = 0.0
        # The reference builds both branches whenever their config blocks exist (lightning_model.py:37-60).  Here the
        # shipped config keeps the blocks (schema compatibility) and two switches turn the branches on; passing the
        # blocks programmatically with the switches absent enables them, like the reference does.
        want_d = discriminator_config is not None and bool(training_config.get("use_adversarial_loss", True))
        want_p = bool(perception_loss_config) and bool(training_config.get("use_perception_loss", True))
        if want_d:
            from src.models.discriminator import DiscriminatorN_IN
            d_args = to_container(dict(discriminator_config.get("args", {})))
            if d_args.get("input_channels") in ("auto", None):
                d_args["input_channels"] = 3          # the critic sees RGB patches (post / generated) only
            self.discriminator = DiscriminatorN_IN(**d_args)
        if want_p:
            from src.models.perception import PerceptualVGG19
            pm = perception_loss_config["perception_model"]
            self.perception_loss_model = PerceptualVGG19(**to_container(dict(pm.get("args", {}))))
            self.perception_loss_weight = float(perception_loss_config["weight"])
        self.reconstruction_criterion = getattr(nn, training_config["reconstruction_criterion"])()
        self.adversarial_criterion = getattr(nn, training_config.get("adversarial_criterion", "MSELoss"))()
        self.d_grad_sync = None
        # one generator forward per adversarial step instead of the reference's two identical ones (see
        # _shared_generator_pass); training.share_generator_pass=false restores the literal two-pass order
        self.share_generator_pass = bool(training_config.get("share_generator_pass", True)) and \
            all(bn.momentum is not None for bn in self.generator.modules() if isinstance(bn, nn.BatchNorm2d))
        self.use_cuda_graph = bool(training_config.get("cuda_graph", True))
        self.grad_sync = None
        self._optimizers = None
        self._graphed = None

    # ------------------------------------------------------------------ Lightning-shaped hooks
    def _adam(self, module: nn.Module, key: str):
        oc = to_container(dict(self.optimizer_config[key]))
        oc["betas"] = tuple(oc.get("betas", (0.9, 0.999)))
        fused = bool(oc.pop("fused", True))
        if fused and next(module.parameters()).is_cuda:
            # clip_grad_norm_ + Adam as two native launches over all tensors (pbt_b200/optim.py); same state layout
            from pbt_b200.optim import FusedClipAdam
            return FusedClipAdam(module.parameters(), **oc)
        if self.use_cuda_graph and next(module.parameters()).is_cuda:
            oc["capturable"] = True   # the whole step is replayed as one CUDA graph (pbt_b200/graphs.py)
        return torch.optim.Adam(module.parameters(), **oc)

    def configure_optimizers(self):
        """[opt_g] or [opt_g, opt_d] (reference lightning_model.py:323-341)"""
        opts = [self._adam(self.generator, "generator")]
        if self.discriminator is not None:
            opts.append(self._adam(self.discriminator, "discriminator"))
        return opts

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
        out = self.full_step(combined_input, batch["post"])
        self._log_losses(out)
        return out

    def full_step(self, combined_input: torch.Tensor, post: torch.Tensor) -> Dict[str, torch.Tensor]:
        """critic update (when enabled), then generator update — reference training_step, lightning_model.py:224-250.
        No host synchronisation inside: the same body is captured by the CUDA-graph steps."""
        opts = self.optimizers()
        opt_g = opts[0]
        out: Dict[str, torch.Tensor] = {}
        generated = None
        if self.discriminator is not None:
            opt_d = opts[1]
            opt_d.zero_grad(set_to_none=True)
            if self.share_generator_pass:
                generated = self._shared_generator_pass(combined_input)
            d_loss = self._discriminator_step(combined_input, post, generated)
            d_loss["loss"].backward()
            if self.d_grad_sync is not None:
                self.d_grad_sync.collect_from_params()
                self.d_grad_sync.finish()
            self._clip_and_step(opt_d, self.discriminator)
            out.update({k: v.detach() for k, v in d_loss.items() if k != "loss"})
        opt_g.zero_grad(set_to_none=True)
        g_loss = self._generator_step(combined_input, {"post": post}, generated)
        g_loss["loss"].backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()          # mean over ranks, before the clip (DDP semantics of the reference)
        self._clip_and_step(opt_g, self.generator)
        out.update({k: v.detach() for k, v in g_loss.items()})
        return out

    def _clip_and_step(self, opt, module: nn.Module) -> None:
        clip = self.training_config["gradient_clip_val"] if self.training_config.get("use_gradient_clipping", False) else None
        if hasattr(opt, "last_grad_norm"):       # FusedClipAdam: the clip is part of the optimiser launch
            opt.step(max_grad_norm=clip)
        else:
            # torch.optim.Adam (optimizer.*.fused=false) has no skip-on-overflow: a sweep whose fp16 gradients overflowed would
            # write NaN into the weights.  Replace non-finite gradients by zeros without a host round trip (graph-capturable);
            # FusedClipAdam skips such a step entirely.
            grads = [p.grad for p in module.parameters() if p.grad is not None]
            if grads:
                finite = torch.stack([torch.isfinite(g).all() for g in grads]).all()
                for g in grads:
                    g.copy_(torch.where(finite, torch.nan_to_num(g, nan=0.0, posinf=0.0, neginf=0.0), torch.zeros_like(g)))
            if clip is not None:
                torch.nn.utils.clip_grad_norm_(module.parameters(), clip)
            opt.step()

    def _log_losses(self, out: Dict[str, torch.Tensor]) -> None:
        tr = getattr(self, "trainer", None)
        if tr is not None and getattr(tr, "global_step", 0) % max(1, int(self.training_config.get("log_every_n_steps", 10))) == 0:
            names = {"margin_loss": "g_image_loss"}      # metric names of the reference's _log_metrics
            self.log_dict({names.get(k, k): float(v) for k, v in out.items() if k != "loss"})

    def _shared_generator_pass(self, combined_input: torch.Tensor) -> torch.Tensor:
        """ONE generator forward for both halves of the step.  The reference runs the generator twice per step on the
        same input with the same weights (no-grad for the critic, lightning_model.py:296-297, then with autograd, :262;
        opt_g steps only afterwards), so the two outputs are identical: the pass with autograd serves both, its
        detached result feeds the critic.  The only side effect of the second train()-mode pass — another momentum
        update of the BatchNorm running statistics with the same batch moments — is applied in closed form:
        r1 = (1-mu) r0 + mu m,  r2 = (1-mu) r1 + mu m  =>  r2 = (2-mu) r1 - (1-mu) r0."""
        bns = [m for m in self.generator.modules() if isinstance(m, nn.BatchNorm2d)] if self.generator.training else []
        before = [(bn.running_mean.clone(), bn.running_var.clone()) for bn in bns]
        generated = self.generator(combined_input)
        with torch.no_grad():
            for bn, (m0, v0) in zip(bns, before):
                mu = bn.momentum
                bn.running_mean.mul_(2.0 - mu).sub_(m0, alpha=1.0 - mu)
                bn.running_var.mul_(2.0 - mu).sub_(v0, alpha=1.0 - mu)
                bn.num_batches_tracked += 1
        return generated

    def _discriminator_step(self, combined_input: torch.Tensor, post: torch.Tensor,
                            generated: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """reference lightning_model.py:294-319.  Without a shared pass the generator runs the native no-grad path in
        train() mode (its BatchNorm running statistics advance, as they do in the reference).  Real and generated
        patches go through the critic as ONE batch: its convolutions and InstanceNorms are per sample, so the logits
        are those of two separate calls (BatchNorm critics, whose statistics would mix, keep the two calls)."""
        if generated is None:
            with torch.no_grad():
                generated = self.generator(combined_input)
        generated = generated.detach()
        per_sample = not any(isinstance(m, nn.BatchNorm2d) for m in self.discriminator.modules()) and \
            not (self.discriminator.use_noise and self.discriminator.training)
        if per_sample and generated.shape == post.shape:
            real_labels, fake_labels = self.discriminator(torch.cat([post, generated], dim=0))[0].chunk(2, dim=0)
        else:
            real_labels, _ = self.discriminator(post)
            fake_labels, _ = self.discriminator(generated)
        real_loss = self.adversarial_criterion(real_labels, torch.ones_like(real_labels))
        fake_loss = self.adversarial_criterion(fake_labels, torch.zeros_like(fake_labels))
        d_loss = (real_loss + fake_loss) * 0.5
        return {"loss": d_loss, "d_real_loss": real_loss, "d_fake_loss": fake_loss, "d_total_loss": d_loss}

    def graphed_training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int):
        """the same generator step replayed as ONE CUDA graph (static shapes; falls back to training_step for the
        ragged last batch of an epoch)"""
        from pbt_b200.graphs import GraphedGeneratorStep
        x = batch["combined_input"]
        if self.discriminator is not None or self.perception_loss_model is not None:
            from pbt_b200.graphs import GraphedGanStep
            if self._graphed is None:
                self._graphed = GraphedGanStep(self, tuple(x.shape))
            if tuple(x.shape) != tuple(self._graphed.x.shape):
                return self.training_step(batch, batch_idx)
            out = self._graphed(x, batch["post"])
            self._log_losses(out)
            return out
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

    def _generator_step(self, combined_input, batch, generated: Optional[torch.Tensor] = None):
        if generated is None:
            generated = self.generator(combined_input)
        losses = {}
        if self.training_config["use_image_loss"]:
            losses["margin_loss"] = self.reconstruction_criterion(generated, batch["post"]) * \
                self.training_config["reconstruction_weight"]
        if self.perception_loss_model is not None:      # reference lightning_model.py:270-275
            pm = self.perception_loss_model
            if hasattr(pm, "feature_mse"):              # PerceptualVGG19: native kernels on CUDA (pbt_b200/perceptual.py)
                per = pm.feature_mse(generated, batch["post"])
            else:
                per = ((pm(generated)[1] - pm(batch["post"].detach())[1]) ** 2).mean()
            losses["g_perception_loss"] = per * self.perception_loss_weight
        if self.discriminator is not None:              # reference lightning_model.py:277-283
            # only dL/d(generated) is needed from this pass: the critic's own gradients of the generator loss are
            # discarded by the reference (opt_d.zero_grad() opens the next step), so they are not computed at all
            frozen = [p for p in self.discriminator.parameters() if p.requires_grad]
            for p in frozen:
                p.requires_grad_(False)
            try:
                fake_labels, _ = self.discriminator(generated)
            finally:
                for p in frozen:
                    p.requires_grad_(True)
            losses["g_adversarial_loss"] = self.adversarial_criterion(fake_labels, torch.ones_like(fake_labels)) * \
                self.training_config["adversarial_weight"]
        total = sum(losses.values())
        losses["g_total_loss"] = total
        return {"loss": total, **losses}
