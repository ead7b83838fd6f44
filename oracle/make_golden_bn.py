"""Generate tests/golden/gen_bn_vectors.npz (and gen_plain_vectors.npz, the norm-free variant) from the UNMODIFIED reference GeneratorJ(norm_layer='batch_norm')
(needs /root/reference; build container only):   python oracle/make_golden_bn.py

Weights are not stored: the module is built under torch.manual_seed(31) (the drop-in module tree reproduces the reference
initialisation bit for bit) and its BatchNorm affine parameters are then set to gamma_i = 1 + 0.25 sin(i), beta_i =
0.1 cos(i) so that the affine terms matter.  Stored: input / target patches, train-mode output, loss, running statistics
after that pass, one-step gradients (full for tensors up to 20k elements, (sum, sum of squares) for all), eval-mode output.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
from src.models.generator import GeneratorJ  # noqa: E402


def set_affine(g):
    with torch.no_grad():
        for m in g.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                i = torch.arange(m.num_features, dtype=torch.float32)
                m.weight.copy_(1 + 0.25 * torch.sin(i))
                m.bias.copy_(0.1 * torch.cos(i))


def plain():
    """norm_layer='none': no norm layers at all (reference :83-87); weights scaled up x3 after the seeded init so that the
    un-normalised activations do not decay to nothing through 20 convs of N(0, 0.02) weights"""
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    x, tgt = torch.from_numpy(vec["x"][:12]).contiguous(), torch.from_numpy(vec["target"][:12]).contiguous()
    torch.manual_seed(32)
    g = GeneratorJ(input_channels=3, use_bias=True, norm_layer="none")
    with torch.no_grad():
        for m in g.modules():
            if isinstance(m, torch.nn.Conv2d):
                m.weight.mul_(3.0)
                m.bias.copy_(0.05 * torch.sin(torch.arange(m.bias.numel(), dtype=torch.float32)))
    g.train()
    y = g(x)
    loss = torch.nn.functional.l1_loss(y, tgt) * 4.0
    loss.backward()
    out = {"y_train": y.detach().numpy(), "loss": np.array(float(loss))}
    for k, p in g.named_parameters():
        v = p.grad.double()
        out[f"grad/moments/{k}"] = np.array([float(v.sum()), float((v * v).sum())])
        if v.numel() <= 20000:
            out[f"grad/full/{k}"] = p.grad.numpy().copy()
    g.eval()
    with torch.no_grad():
        out["y_eval"] = g(x).numpy()
    np.savez_compressed(os.path.join(GOLD, "gen_plain_vectors.npz"), **out)
    print("wrote gen_plain_vectors.npz, loss", float(loss), "y std", float(y.std()))


def main():
    torch.set_num_threads(8)
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    x, tgt = torch.from_numpy(vec["x"][:12]).contiguous(), torch.from_numpy(vec["target"][:12]).contiguous()
    torch.manual_seed(31)
    g = GeneratorJ(input_channels=3, use_bias=True, norm_layer="batch_norm")
    set_affine(g)
    g.train()
    y = g(x)
    loss = torch.nn.functional.l1_loss(y, tgt) * 4.0
    loss.backward()
    out = {"x": x.numpy(), "target": tgt.numpy(), "y_train": y.detach().numpy(), "loss": np.array(float(loss))}
    for k, p in g.named_parameters():
        v = p.grad.double()
        out[f"grad/moments/{k}"] = np.array([float(v.sum()), float((v * v).sum())])
        if v.numel() <= 20000:
            out[f"grad/full/{k}"] = p.grad.numpy().copy()
    for k, v in g.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out[f"running/{k}"] = v.numpy().copy()
    g.eval()
    with torch.no_grad():
        out["y_eval"] = g(x).numpy()
    np.savez_compressed(os.path.join(GOLD, "gen_bn_vectors.npz"), **out)
    print("wrote gen_bn_vectors.npz, loss", float(loss))


if __name__ == "__main__":
    main()
    plain()
