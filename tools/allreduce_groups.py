"""run under torchrun (N >= 2): the graph-replayed C3 training step with the gradient exchange split into different numbers of
all-reduce groups, against the same step without any exchange.   torchrun --nproc-per-node 2 tools/allreduce_groups.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbt_b200.generator import GeneratorJ  # noqa: E402
from pbt_b200.graphs import GraphedGeneratorStep  # noqa: E402
from pbt_b200.optim import FusedClipAdam  # noqa: E402
from pbt_b200.parallel import GradAllReduce, init_distributed  # noqa: E402

rank, world, local = init_distributed("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
z = np.load(os.path.join(ROOT, "tests", "golden", "gen_cin9_trained.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
B, P = 80, 80
g0 = torch.Generator(device=dev).manual_seed(5 + rank)
x = torch.rand((B, 9, P, P), generator=g0, device=dev) * 2 - 1
t = torch.rand((B, 3, P, P), generator=g0, device=dev) * 2 - 1


def run(grouping):
    gen = GeneratorJ(input_channels=9, use_bias=True)
    gen.load_state_dict(sd, strict=True)
    gen = gen.to(dev).train()
    opt = FusedClipAdam(gen.parameters(), lr=4e-4, weight_decay=1e-5, max_grad_norm=0.5)
    ar = None if grouping == "none" else GradAllReduce(list(gen.named_parameters()), world=world, grouping=grouping).attach(gen)
    step = GraphedGeneratorStep(gen, opt, (B, 9, P, P), clip=0.5, grad_sync=ar)
    for _ in range(5):
        step(x, t)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 60
    for _ in range(reps):
        step(x, t)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    del step
    torch.cuda.synchronize()
    return float(ms), (len(ar.group_bounds) if ar else 0)


base = None
for grouping in ("none", "block", "pairs", "coarse", "single", "none"):
    ms, n = run(grouping)
    if base is None:
        base = ms
    if rank == 0:
        print(f"world {world} grouping {grouping:7s} ({n:2d} all-reduce calls per step): {ms:.3f} ms/step  (+{ms - base:.3f} ms vs no exchange)", flush=True)
sys.stdout.flush()
os._exit(0)
