"""Full adversarial training step (critic + generator, one CUDA-graph replay) against the G-only step at the same
shape:   python tools/gan_step_bench.py [N CIN P]      (default: config C3, 80 x 9 x 80^2)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightning_model import StyleTransferModel  # noqa: E402

n, cin, p = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (80, 9, 80)))
gen_cfg = {"type": "GeneratorJ", "args": {"input_channels": cin, "use_bias": True}}
dis_cfg = {"type": "DiscriminatorN_IN", "args": dict(input_channels=3, num_filters=12, n_layers=2, norm_layer="instance_norm",
                                                     use_bias=True)}
train_cfg = {"batch_size": n, "reconstruction_weight": 4.0, "adversarial_weight": 0.5, "use_image_loss": True,
             "reconstruction_criterion": "L1Loss", "adversarial_criterion": "MSELoss", "use_gradient_clipping": True,
             "gradient_clip_val": 0.5, "cuda_graph": True}
adam = {"lr": 0.0004, "betas": [0.9, 0.999], "weight_decay": 0.00001}
x = torch.rand(n, cin, p, p, device="cuda") * 2 - 1
t = torch.rand(n, 3, p, p, device="cuda") * 2 - 1
batch = {"combined_input": x, "post": t}
for name, dcfg in (("G-only", None), ("GAN (critic + generator)", dis_cfg)):
    torch.manual_seed(0)
    m = StyleTransferModel(gen_cfg, dcfg, dict(train_cfg), {"generator": dict(adam), "discriminator": dict(adam)},
                           {"additional_channels": {}}).cuda().train()
    m._optimizers = m.configure_optimizers()
    for i in range(5):
        out = m.graphed_training_step(batch, i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for i in range(reps):
        out = m.graphed_training_step(batch, i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"[{name} {n}x{cin}x{p}^2] {ms:.3f} ms/step -> {n / ms * 1e3:.0f} patches/s   loss {float(out['loss']):.4f}")
