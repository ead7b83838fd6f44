"""per-parameter gradient PSNR of the smoke configuration (2 x 3 x 32 x 32, random init); arg 'noside' runs the weight
gradients on the main stream"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import generator_oracle as go  # noqa: E402
from pbt_b200.generator import GeneratorJ  # noqa: E402

n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
g = GeneratorJ(input_channels=3, use_bias=True).cuda().train()
sd = {k: v.detach().cpu().clone() for k, v in g.state_dict().items()}
x = torch.rand(n, 3, 32, 32) * 2 - 1
t = torch.rand(n, 3, 32, 32) * 2 - 1
y = g(x.cuda())
if len(sys.argv) > 1 and sys.argv[1] == "noside":
    g._engine.side_stream = lambda: torch.cuda.current_stream()
loss = torch.nn.functional.l1_loss(y, t.cuda()) * 4.0
loss.backward()
torch.cuda.synchronize()
ry, rloss, rg = go.loss_and_grads(sd, x, t)
for k, p in g.named_parameters():
    ref = rg[k]
    peak = float(ref.abs().max())
    if peak < 1e-12 or float(p.grad.abs().max()) == 0.0:
        continue
    mse = float(((p.grad.cpu().double() - ref.double()) ** 2).mean())
    ps = 200.0 if mse == 0 else 10 * math.log10(peak * peak / mse)
    print(f"{k:40s} psnr {ps:6.1f} dB  peak {peak:.3e}")
