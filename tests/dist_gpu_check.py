"""Run under torchrun with >= 2 GPUs (tests/test_gpu_round2.py launches it): NCCL data-parallel parity of the training path
and the sharded frame loop.  Prints DIST_GPU_CHECK_OK on rank 0 when every assertion held on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    from pbt_b200.generator import GeneratorJ
    from pbt_b200.graphs import GraphedGeneratorStep
    from pbt_b200.inference import FrameStylizer
    from pbt_b200.optim import FusedClipAdam
    from pbt_b200.parallel import GradAllReduce, broadcast_module_state, init_distributed, replicas_identical
    rank, world, local = init_distributed("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # (0) every rank initialises differently (no seed); the broadcast makes them identical
    g = GeneratorJ(input_channels=3, use_bias=True).to(dev).train()
    assert not replicas_identical(g), "unseeded ranks should start from different weights"
    broadcast_module_state(g)
    assert replicas_identical(g)
    z = np.load(os.path.join(GOLD, "gen_c3_trained.npz"))
    g.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    n = vec["x"].shape[0] // world
    x = torch.from_numpy(vec["x"][rank * n:(rank + 1) * n]).to(dev)
    t = torch.from_numpy(vec["target"][rank * n:(rank + 1) * n]).to(dev)
    # (1) NCCL-reduced gradient == mean of the per-rank gradients
    (torch.nn.functional.l1_loss(g(x), t) * 4.0).backward()
    mine = g._engine.grad_bucket().flat.clone()
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    expect = torch.stack(every).mean(0)
    g.zero_grad(set_to_none=True)
    ar = GradAllReduce(list(g.named_parameters()), world=world).attach(g)
    (torch.nn.functional.l1_loss(g(x), t) * 4.0).backward()
    ar.finish()
    torch.cuda.synchronize()
    err = float((ar.flat - expect).abs().max() / expect.abs().max())
    assert err < 1e-3, err
    assert all(p.grad.data_ptr() == ar.bucket.views[k].data_ptr() for k, p in g.named_parameters())
    # (2) graph-replayed data-parallel steps on different per-rank batches keep the weights bit-identical
    g.zero_grad(set_to_none=True)
    opt = FusedClipAdam(g.parameters(), lr=4e-4, betas=(0.9, 0.999), weight_decay=1e-5, max_grad_norm=0.5)
    step = GraphedGeneratorStep(g, opt, tuple(x.shape), clip=0.5, grad_sync=ar)
    w0 = g.conv11[0].weight.detach().clone()
    losses = [float(step(x, t)) for _ in range(12)]
    assert replicas_identical(g, buffers=False), "replicas diverged"
    assert not torch.equal(w0, g.conv11[0].weight.detach()) and losses[-1] < losses[0], losses
    # (3) sharded frame loop: the shares of all ranks tile the video and equal the single-process result
    gen = torch.Generator(device=dev).manual_seed(3)
    video = torch.randint(0, 256, (9, 64, 96, 3), generator=gen, device=dev, dtype=torch.uint8)
    sty = FrameStylizer(g)
    out = torch.zeros((9, 64, 96, 3), dtype=torch.uint8, device=dev)
    lo, hi = sty.stylize_video(video, out)          # rank / world from the torchrun environment
    whole = sty.stylize_device(video)
    assert torch.equal(out[lo:hi], whole[lo:hi]) and int(out[:lo].sum()) == 0 and int(out[hi:].sum()) == 0
    cnt = torch.tensor([hi - lo], device=dev)
    dist.all_reduce(cnt)
    assert int(cnt) == 9
    # (4) the reference-shaped trainer end to end under torchrun: unseeded ranks, adversarial branch on, CUDA-graph steps,
    # early-stop / top-k decisions on rank-averaged losses; afterwards generator and critic replicas are bit-identical
    import tempfile
    import lightning_model as lm
    from pbt_b200.config import compose
    from pbt_b200.trainer import Trainer
    mini = os.path.join(GOLD, "mini_dataset")
    tmp = tempfile.mkdtemp(prefix=f"pbt_dist_{rank}_")
    cfg = compose(os.path.join(ROOT, "config"), "config",
                  [f"data.dir_pre={mini}/input", f"data.dir_post={mini}/output", f"data.dir_mask={mini}/mask", "data.patch_size=32",
                   f"data.additional_channels.point_vector.path={mini}/guide", "training.batch_size=8", f"training.output_dir={tmp}",
                   "+training.max_steps=5", "training.max_epochs=1", "training.log_every_n_steps=2"])
    model = lm.StyleTransferModel(cfg.model.generator, cfg.model.discriminator, cfg.training, cfg.optimizer, cfg.data,
                                  cfg.model.perception_loss)
    tr = Trainer(max_epochs=1, max_steps=5, output_dir=tmp, log_every_n_steps=2)
    tr.fit(model)
    assert tr.global_step == 5
    assert replicas_identical(model.generator, buffers=False), "trainer: generator replicas diverged"
    assert replicas_identical(model.discriminator, buffers=False), "trainer: critic replicas diverged"
    if rank == 0:
        assert os.path.exists(os.path.join(tmp, "checkpoints", "last.ckpt"))
    model._graphed = None
    del step                      # a live CUDA graph that captured NCCL work keeps the communicator busy at shutdown
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"DIST_GPU_CHECK_OK world={world} allreduce_err={err:.2e} losses {losses[0]:.4f}->{losses[-1]:.4f}")
    sys.stdout.flush()
    os._exit(0)                   # skip the process-group teardown: it can block behind graph-captured collectives


if __name__ == "__main__":
    main()
