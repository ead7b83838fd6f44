"""world_size-2 tests of the multi-GPU host logic on the CPU (gloo): gradient bucket all-reduce, frame sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from pbt_b200.generator import GeneratorJ
    from pbt_b200.parallel import GradAllReduce, init_distributed, shard_range
    r, w, _ = init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    g = GeneratorJ(input_channels=3, use_bias=True)        # parameters only; no forward on the CPU
    named = list(g.named_parameters())
    ar = GradAllReduce(named, world=world)
    # emulate the order in which the backward sweep publishes gradients: tail, decoder, trunk
    order = sorted(range(len(named)), key=lambda i: (0 if named[i][0].startswith(("output", "smoothers", "conv11")) else
                                                     1 if named[i][0].startswith("upsample") else 2))
    gen = torch.Generator().manual_seed(100 + rank)
    local = {}
    for i in order:
        name, p = named[i]
        local[name] = torch.randn(p.shape, generator=gen)
        ar.grad_ready(name, local[name])
    ar.finish()
    # expected mean over ranks, recomputed locally from both ranks' generators
    ok = True
    gens = [torch.Generator().manual_seed(100 + k) for k in range(world)]
    for i in order:
        name, p = named[i]
        exp = sum(torch.randn(p.shape, generator=gk) for gk in gens) / world
        ok &= torch.allclose(p.grad, exp, atol=1e-6)
    # a second step reuses the bucket
    for i in order:
        ar.grad_ready(named[i][0], torch.full(named[i][1].shape, float(rank + 1)))
    ar.finish()
    ok &= all(torch.allclose(p.grad, torch.full_like(p, (1 + world) / 2)) for _, p in named)
    lo, hi = shard_range(2000, rank, world)
    t = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(t)
    ok &= int(t) == 2000
    q.put((rank, bool(ok), ar.nbytes))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_grad_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    assert all(nb == 3265027 * 4 for _, _, nb in res)   # 13.06 MB fp32 bucket
