"""Drop-in for the reference's train.py: `python train.py [key=value ...]` with the reference's config schema.
Multi-GPU: `torchrun --nproc-per-node N train.py training.devices=N` (one process per GPU, NCCL all-reduce)."""
import os
import sys

import yaml

from lightning_model import StyleTransferModel
from pbt_b200.config import compose, to_container
from pbt_b200.trainer import Trainer


def train(cfg) -> None:
    print(yaml.safe_dump(to_container(cfg), sort_keys=False))
    os.makedirs(cfg.training.output_dir, exist_ok=True)
    with open(os.path.join(cfg.training.output_dir, "config.yaml"), "w") as f:
        yaml.safe_dump(to_container(cfg), f, sort_keys=False)
    model = StyleTransferModel(generator_config=cfg.model.generator, discriminator_config=cfg.model.get("discriminator"),
                               training_config=cfg.training, optimizer_config=cfg.optimizer, data_config=cfg.data,
                               perception_loss_config=cfg.model.get("perception_loss"))
    trainer = Trainer(max_epochs=cfg.training.max_epochs, max_steps=cfg.training.get("max_steps"),
                      output_dir=cfg.training.output_dir, log_every_n_steps=cfg.training.log_every_n_steps,
                      early_stopping_patience=cfg.training.early_stopping_patience if cfg.training.early_stopping else None,
                      steps_per_epoch=cfg.training.get("steps_per_epoch"), devices=cfg.training.devices)
    trainer.fit(model)
    print("Training completed!")


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    train(compose(os.path.join(here, "config"), "config", sys.argv[1:]))
