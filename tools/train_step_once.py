"""a few G-only training steps at the C3 shape (for ncu launch lists)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200.generator import GeneratorJ  # noqa: E402

n, cin, p = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (80, 9, 80)))
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(0)
g = GeneratorJ(input_channels=cin, use_bias=True).cuda().train()
opt = torch.optim.Adam(g.parameters(), lr=4e-4, weight_decay=1e-5)
x = torch.rand(n, cin, p, p, device="cuda") * 2 - 1
t = torch.rand(n, 3, p, p, device="cuda") * 2 - 1
for i in range(steps):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.l1_loss(g(x), t) * 4.0
    loss.backward()
    torch.nn.utils.clip_grad_norm_(g.parameters(), 0.5)
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
