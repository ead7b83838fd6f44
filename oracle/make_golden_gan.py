"""Generate tests/golden/gan_step.npz by running the UNMODIFIED reference GeneratorJ and DiscriminatorN_IN through two
full adversarial training steps (needs /root/reference; build container only):   python oracle/make_golden_gan.py

The reference's lightning_model.py cannot be imported here (pytorch_lightning / hydra / omegaconf are not installed), so
the step below drives the reference MODULES with torch.optim.Adam and clip_grad_norm_ in the order of
lightning_model.py:201-250 (critic first, then the generator against the updated critic), with the shipped
hyper-parameters (config/training/default.yaml, config/optimizer/default.yaml, config/model/default.yaml).
Contents: critic init / final state, per-step losses, first-step raw gradients (critic: all; generator: tensors up to
20k elements in full + (sum, sum of squares) of every tensor), generator state after two steps in the same form.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)

from src.models.discriminator import DiscriminatorN_IN  # noqa: E402
from src.models.generator import GeneratorJ  # noqa: E402

SMALL = 20000


def digest(prefix, tensors, out):
    for k, v in tensors.items():
        v = v.detach().double()
        out[f"{prefix}/moments/{k}"] = np.array([float(v.sum()), float((v * v).sum())])
        if v.numel() <= SMALL:
            out[f"{prefix}/full/{k}"] = v.float().numpy()


def main():
    torch.set_num_threads(8)
    sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, "gen_c3_trained.npz")).items()}
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    x, post = torch.from_numpy(vec["x"][:8]).contiguous(), torch.from_numpy(vec["target"][:8]).contiguous()
    G = GeneratorJ(input_channels=3, use_bias=True)
    G.load_state_dict(sd, strict=True)
    torch.manual_seed(2024)
    D = DiscriminatorN_IN(input_channels=3, num_filters=12, n_layers=2, use_noise=False, noise_sigma=0.2,
                          norm_layer="instance_norm", use_bias=True)
    G.train(), D.train()
    adam = dict(lr=0.0004, betas=(0.9, 0.999), weight_decay=0.00001)
    opt_g, opt_d = torch.optim.Adam(G.parameters(), **adam), torch.optim.Adam(D.parameters(), **adam)
    rec, adv = torch.nn.L1Loss(), torch.nn.MSELoss()
    out = {"seed": np.array(2024), "n": np.array(8)}
    for k, v in D.state_dict().items():
        out[f"d_init/{k}"] = v.numpy().copy()
    for step in range(2):
        opt_d.zero_grad()
        with torch.no_grad():
            generated = G(x)
        real_labels, _ = D(post)
        real_loss = adv(real_labels, torch.ones_like(real_labels))
        fake_labels, _ = D(generated)
        fake_loss = adv(fake_labels, torch.zeros_like(fake_labels))
        d_loss = (real_loss + fake_loss) * 0.5
        d_loss.backward()
        if step == 0:
            digest("d_grad0", {k: p.grad for k, p in D.named_parameters()}, out)
        torch.nn.utils.clip_grad_norm_(D.parameters(), 0.5)
        opt_d.step()
        opt_g.zero_grad()
        generated = G(x)
        margin = rec(generated, post) * 4.0
        fake_labels, _ = D(generated)
        g_adv = adv(fake_labels, torch.ones_like(fake_labels)) * 0.5
        total = sum({"margin_loss": margin, "g_adversarial_loss": g_adv}.values())
        total.backward()
        if step == 0:
            digest("g_grad0", {k: p.grad for k, p in G.named_parameters() if p.grad is not None}, out)
        torch.nn.utils.clip_grad_norm_(G.parameters(), 0.5)
        opt_g.step()
        out[f"losses/{step}"] = np.array([float(real_loss), float(fake_loss), float(d_loss), float(margin), float(g_adv),
                                          float(total)])
    for k, v in D.state_dict().items():
        out[f"d_final/{k}"] = v.numpy().copy()
    digest("g_final", {k: v for k, v in G.state_dict().items() if v.is_floating_point()}, out)
    np.savez_compressed(os.path.join(GOLD, "gan_step.npz"), **out)
    print("wrote gan_step.npz:", {k: out[k] for k in ("losses/0", "losses/1")})


if __name__ == "__main__":
    main()
