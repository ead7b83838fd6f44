"""eager vs CUDA-graph G-only training step: python tools/train_graph_bench.py N CIN P"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200.generator import GeneratorJ  # noqa: E402
from pbt_b200.graphs import GraphedGeneratorStep  # noqa: E402

n, cin, p = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (80, 9, 80)))


def make():
    torch.manual_seed(0)
    g = GeneratorJ(input_channels=cin, use_bias=True).cuda().train()
    return g


x = torch.rand(n, cin, p, p, device="cuda") * 2 - 1
t = torch.rand(n, 3, p, p, device="cuda") * 2 - 1
# eager reference run
g1 = make()
o1 = torch.optim.Adam(g1.parameters(), lr=4e-4, weight_decay=1e-5)
losses1 = []
for i in range(6):
    o1.zero_grad(set_to_none=True)
    loss = torch.nn.functional.l1_loss(g1(x), t) * 4.0
    loss.backward()
    torch.nn.utils.clip_grad_norm_(g1.parameters(), 0.5)
    o1.step()
    losses1.append(float(loss))
# graphed run: 3 warm-up steps happen inside the constructor, then replays
g2 = make()
from pbt_b200.optim import FusedClipAdam  # noqa: E402
o2 = FusedClipAdam(g2.parameters(), lr=4e-4, weight_decay=1e-5, max_grad_norm=0.5)
step = GraphedGeneratorStep(g2, o2, (n, cin, p, p), clip=0.5)
step.x.copy_(x)
step.target.copy_(t)
losses2 = []
# the constructor ran warm-up + capture on zero inputs; restart from identical weights for the comparison
g2.load_state_dict(make().state_dict())
o2._m.zero_()
o2._v.zero_()
o2._state.zero_()
for i in range(6):
    losses2.append(float(step(x, t)))
print("eager  losses", [round(v, 5) for v in losses1])
print("graph  losses", [round(v, 5) for v in losses2])
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    step(x, t)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
print(f"[graph {n}x{cin}x{p}^2] {ms:.2f} ms/step -> {n / ms * 1e3:.0f} patches/s")
