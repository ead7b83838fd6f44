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
    from pbt_b200.parallel import GradAllReduce, broadcast_module_state, init_distributed, replicas_identical, shard_range
    r, w, _ = init_distributed("gloo")
    assert (r, w) == (rank, world)
    # NO common seed: every process initialises its own weights; the trainer's broadcast (DDP's initial broadcast in the
    # reference, train.py:93-94) is what makes the replicas identical
    g = GeneratorJ(input_channels=3, use_bias=True)        # parameters only; no forward on the CPU
    differs = not replicas_identical(g)
    broadcast_module_state(g)
    same = replicas_identical(g)
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
    ok = differs and same
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


def _critic_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from lightning_model import StyleTransferModel
    from pbt_b200.parallel import GradAllReduce, init_distributed
    init_distributed("gloo")
    tcfg = {"batch_size": 4, "reconstruction_weight": 4.0, "adversarial_weight": 0.5, "use_image_loss": True,
            "reconstruction_criterion": "L1Loss", "adversarial_criterion": "MSELoss", "use_gradient_clipping": True,
            "gradient_clip_val": 0.5, "cuda_graph": False}
    adam = {"lr": 4e-4, "betas": [0.9, 0.999], "weight_decay": 1e-5}
    torch.manual_seed(0)       # identical initial critic on every rank
    m = StyleTransferModel({"args": {"input_channels": 3, "use_bias": True}},
                           {"args": {"input_channels": 3, "num_filters": 12, "n_layers": 2, "use_bias": True}}, tcfg,
                           {"generator": dict(adam), "discriminator": dict(adam)}, {"additional_channels": {}})
    d = m.discriminator
    m.d_grad_sync = GradAllReduce(list(d.named_parameters()), world=world)
    gen = torch.Generator().manual_seed(7)
    post_all, fake_all = torch.rand(4 * world, 3, 32, 32, generator=gen) * 2 - 1, torch.rand(4 * world, 3, 32, 32, generator=gen) * 2 - 1
    # the full-batch critic gradient, computed locally (per-sample InstanceNorm: mean of rank means = full-batch mean)
    full = m._discriminator_step(None, post_all, fake_all)["loss"]
    exp = torch.autograd.grad(full, list(d.parameters()))
    # the rank's share: critic half of full_step (the generator half needs the GPU), generated patches given
    sl = slice(4 * rank, 4 * rank + 4)
    m._discriminator_step(None, post_all[sl], fake_all[sl])["loss"].backward()
    m.d_grad_sync.collect_from_params()
    m.d_grad_sync.finish()
    ok = all(torch.allclose(p.grad, e, rtol=1e-4, atol=1e-7) for p, e in zip(d.parameters(), exp))
    (opt_d,) = m.configure_optimizers()[1:]
    m._clip_and_step(opt_d, d)
    flat = torch.cat([p.detach().flatten() for p in d.parameters()])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    ok &= all(torch.equal(both[0], b) for b in both[1:])       # replicas stay identical after the step
    q.put((rank, bool(ok), m.d_grad_sync.nbytes))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_critic_gradients_are_averaged_across_ranks_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_critic_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    assert all(nb == 4 * (12 * 3 * 16 + 12 + 24 * 12 * 16 + 24 + 48 * 24 * 16 + 48 + 48 * 16 + 1) for _, _, nb in res)
