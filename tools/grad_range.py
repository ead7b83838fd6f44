"""dynamic range of the 16-bit gradient tensors of one backward sweep (fp16 overflow diagnosis): max |g| of every
dgrad / norm-backward output relative to the scaled output gradient"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200 import generator_bwd, ops  # noqa: E402
from pbt_b200.generator import GeneratorJ  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
target = float(sys.argv[1]) if len(sys.argv) > 1 else None
sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, "gen_c3_trained.npz")).items()}
vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
frame, yframe = torch.from_numpy(vec["frame"])[0], torch.from_numpy(vec["y_frame"])[0]
rng = np.random.RandomState(5)
xs, ts = [], []
for _ in range(80):
    y0, x0 = rng.randint(0, frame.shape[1] - 80), rng.randint(0, frame.shape[2] - 80)
    xs.append(frame[:, y0:y0 + 80, x0:x0 + 80])
    ts.append(yframe[:, y0:y0 + 80, x0:x0 + 80].flip(2))
x, t = torch.stack(xs).cuda(), torch.stack(ts).cuda()
g = GeneratorJ(input_channels=3, use_bias=True)
g.load_state_dict(sd, strict=True)
g = g.cuda().train()
y = g(x)
if target is not None:
    g._engine.grad_scale_target = target
log = []
orig_conv, orig_nb = ops.conv_fwd, ops.norm_bwd


def conv_fwd(*a, **k):
    orig_conv(*a, **k)
    out = k.get("out")
    if out is not None:
        v = out.t.float()
        log.append(("dgrad  c=%d %dx%d" % (out.c, out.h, out.w), float(v.abs().nan_to_num(posinf=1e30).max()), bool(torch.isfinite(v).all())))


def norm_bwd(xx, *a, **k):
    orig_nb(xx, *a, **k)
    v = k["dx"].t.float()
    log.append(("normbw c=%d %dx%d" % (xx.c, xx.h, xx.w), float(v.abs().nan_to_num(posinf=1e30).max()), bool(torch.isfinite(v).all())))


ops.conv_fwd, ops.norm_bwd = conv_fwd, norm_bwd
loss = torch.nn.functional.l1_loss(y, t) * 4.0
loss.backward()
torch.cuda.synchronize()
for name, mx, fin in log:
    print(f"{name:28s} max|g| {mx:12.4g}  finite={fin}")
bad = [k for k, p in g.named_parameters() if not torch.isfinite(p.grad).all()]
print("non-finite parameter gradients:", bad)
