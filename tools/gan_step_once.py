"""a few eager adversarial training steps at the C3 shape (for ncu launch lists):
python tools/gan_step_once.py [N CIN P [STEPS [perceptual]]]     (a 5th argument adds the VGG19 tap term, weight 6.0: the
reference's default three-term generator loss; synthetic VGG weights)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightning_model import StyleTransferModel  # noqa: E402

n, cin, p = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (80, 9, 80)))
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
tcfg = {"batch_size": n, "reconstruction_weight": 4.0, "adversarial_weight": 0.5, "use_image_loss": True,
        "reconstruction_criterion": "L1Loss", "adversarial_criterion": "MSELoss", "use_gradient_clipping": True,
        "gradient_clip_val": 0.5, "cuda_graph": False}
adam = {"lr": 4e-4, "betas": [0.9, 0.999], "weight_decay": 1e-5}
torch.manual_seed(0)
m = StyleTransferModel({"args": {"input_channels": cin, "use_bias": True}},
                       {"args": {"input_channels": 3, "num_filters": 12, "n_layers": 2, "use_bias": True}}, tcfg,
                       {"generator": dict(adam), "discriminator": dict(adam)}, {"additional_channels": {}}).cuda().train()
if len(sys.argv) > 5:
    from pbt_b200.perceptual import vgg19_prefix
    from src.models.perception import PerceptualVGG19
    m.perception_loss_model = PerceptualVGG19.from_features(vgg19_prefix(6), [0, 3, 5], use_normalization=False).cuda()
    m.perception_loss_weight = 6.0
m._optimizers = m.configure_optimizers()
x = torch.rand(n, cin, p, p, device="cuda") * 2 - 1
t = torch.rand(n, 3, p, p, device="cuda") * 2 - 1
for i in range(steps):
    out = m.training_step({"combined_input": x, "post": t}, i)
torch.cuda.synchronize()
print({k: round(float(v), 4) for k, v in out.items()})
