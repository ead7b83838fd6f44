"""probe (torchrun, N >= 2): NVLS multicast all-reduce of the 13 MB gradient bucket through torch's symmetric memory against
ncclAllReduce - availability, time, CUDA-graph capturability"""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbt_b200.parallel import init_distributed  # noqa: E402

rank, world, local = init_distributed("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
n = 3293251 + 61        # C3 bucket size (floats), 16-byte multiple
n = (n + 1023) // 1024 * 1024
group = dist.group.WORLD
try:
    buf = sm.empty(n, dtype=torch.float32, device=dev)
    hdl = sm.rendezvous(buf, group.group_name)
    mc = hdl.multicast_ptr
    print(f"[rank {rank}] symmetric bucket ok: multicast_ptr={'0x%x' % mc if mc else mc}, world {hdl.world_size}, signal pad {hdl.signal_pad_size}", flush=True)
except Exception as e:  # noqa: BLE001
    print(f"[rank {rank}] symmetric memory unavailable: {type(e).__name__}: {e}", flush=True)
    os._exit(0)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


plain = torch.zeros(n, device=dev)
res = {}
res["nccl all_reduce (plain buffer)"] = timeit(lambda: dist.all_reduce(plain))
res["nccl all_reduce (symmetric buffer)"] = timeit(lambda: dist.all_reduce(buf))
for name, fn in (("multimem_all_reduce_", lambda: torch.ops.symm_mem.multimem_all_reduce_(buf, "sum", group.group_name)),
                 ("two_shot_all_reduce_", lambda: torch.ops.symm_mem.two_shot_all_reduce_(buf, "sum", group.group_name)),
                 ("one_shot_all_reduce", lambda: torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", group.group_name))):
    try:
        res[name] = timeit(fn)
    except Exception as e:  # noqa: BLE001
        res[name] = f"{type(e).__name__}: {str(e)[:120]}"
# correctness + graph capture of the multimem variant
try:
    buf.fill_(float(rank + 1))
    torch.cuda.synchronize()
    dist.barrier()
    torch.ops.symm_mem.multimem_all_reduce_(buf, "sum", group.group_name)
    torch.cuda.synchronize()
    ok = bool((buf == world * (world + 1) / 2).all())
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        torch.ops.symm_mem.multimem_all_reduce_(buf, "sum", group.group_name)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        torch.ops.symm_mem.multimem_all_reduce_(buf, "sum", group.group_name)
    buf.fill_(1.0)
    dist.barrier()
    g.replay()
    torch.cuda.synchronize()
    res["multimem correct / graph replay correct"] = f"{ok} / {bool((buf == world).all())}"
except Exception as e:  # noqa: BLE001
    res["multimem correctness / capture"] = f"{type(e).__name__}: {str(e)[:160]}"
if rank == 0:
    for k, v in res.items():
        print(f"  {k}: {v if isinstance(v, str) else '%.1f us' % v}", flush=True)
sys.stdout.flush()
time.sleep(0.5)
os._exit(0)
