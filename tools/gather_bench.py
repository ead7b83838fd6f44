"""patch sampler gather: the stress case of SURVEY section 8(d) (16,384 patches x 12 channels x 80^2 fp32 = 5.03 GB read
+ 5.03 GB written in one launch) and the C3 batch (80 patches), timed with CUDA events"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200 import ops  # noqa: E402

dev = "cuda"
H, W, K, P = 1080, 1920, 14, 80            # 14 keyframes, pre + 2 guides + post = 4 sources of 3 channels
g = torch.Generator(device=dev).manual_seed(0)
imgs = [[torch.rand((3, H, W), generator=g, device=dev) for _ in range(K)] for _ in range(4)]
table = torch.tensor([[t.data_ptr() for t in row] for row in imgs], dtype=torch.int64, device=dev)
hw = torch.tensor([[H, W]] * K, dtype=torch.int32, device=dev)
for B in (80, 16384):
    pos = torch.stack([torch.randint(0, K, (B,), generator=g, device=dev), torch.randint(40, H - 40, (B,), generator=g, device=dev),
                       torch.randint(40, W - 40, (B,), generator=g, device=dev)], 1).int().contiguous()
    comb = torch.empty((B, 9, P, P), device=dev)
    post = torch.empty((B, 3, P, P), device=dev)
    run = lambda: ops.patch_gather(table, 4, K, 3, hw, pos, P, [comb, comb, comb, post], [0, 3, 6, 0], [9, 9, 9, 3])
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if B < 1000 else 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 2.0 * B * 12 * P * P * 4
    print(f"gather B={B}: {ms * 1e3:9.1f} us  {nbytes / 1e9:6.3f} GB moved  {nbytes / ms / 1e6:8.1f} GB/s "
          f"({nbytes / ms / 1e6 / 6544.3:.2f} of measured HBM peak)")
# spot-check against slicing
b = 5
i, y, x = (int(v) for v in pos[b])
ref = torch.cat([imgs[s][i][:, y - 40:y + 40, x - 40:x + 40] for s in range(3)], 0)
print("bit-exact:", bool(torch.equal(comb[b], ref)) and bool(torch.equal(post[b], imgs[3][i][:, y - 40:y + 40, x - 40:x + 40])))
