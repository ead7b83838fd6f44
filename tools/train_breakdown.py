"""host/device time breakdown of the G-only training step (fresh process; one config)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200 import _native  # noqa: E402
from pbt_b200.generator import GeneratorJ  # noqa: E402

n, cin, p = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (80, 9, 80)))
torch.manual_seed(0)
g = GeneratorJ(input_channels=cin, use_bias=True).cuda().train()
opt = torch.optim.Adam(g.parameters(), lr=4e-4, weight_decay=1e-5)
x = torch.rand(n, cin, p, p, device="cuda") * 2 - 1
t = torch.rand(n, 3, p, p, device="cuda") * 2 - 1


def step(sync):
    tm = {}
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    y = g(x)
    loss = torch.nn.functional.l1_loss(y, t) * 4.0
    if sync:
        torch.cuda.synchronize()
    tm["fwd"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    loss.backward()
    if sync:
        torch.cuda.synchronize()
    tm["bwd"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    torch.nn.utils.clip_grad_norm_(g.parameters(), 0.5)
    opt.step()
    if sync:
        torch.cuda.synchronize()
    tm["opt"] = time.perf_counter() - t0
    return tm


for _ in range(3):
    step(True)
for mode in (True, False):
    torch.cuda.synchronize()
    l0 = _native.LAUNCHES[0]
    t0 = time.perf_counter()
    acc = {}
    reps = 5
    for _ in range(reps):
        for k, v in step(mode).items():
            acc[k] = acc.get(k, 0.0) + v
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / reps * 1e3
    print(f"[{n}x{cin}x{p}^2 sync={mode}] step {tot:.2f} ms; host-side phases: " +
          ", ".join(f"{k} {v / reps * 1e3:.2f} ms" for k, v in acc.items()) +
          f"; native launches/step {(_native.LAUNCHES[0] - l0) // reps}; mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
