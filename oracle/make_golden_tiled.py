"""Generate tests/golden/tiled_golden.npz by calling the UNMODIFIED reference methods of /root/reference/generator.py
(`_process_mask`, `_get_valid_patch_positions`, `process_large_image`) — build container only:
    python oracle/make_golden_tiled.py
hydra / pytorch_lightning / omegaconf are not installed; the reference file only needs them to exist at import time, so
empty stub modules are registered (the reference's lightning_model.py is stubbed too: the methods under test only call
`self.model.generator(patch)` and `self.model.to(device)`).
"""
import importlib.util
import logging
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)

for name in ("pytorch_lightning", "omegaconf", "hydra"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["omegaconf"].DictConfig = type("DictConfig", (dict,), {})
sys.modules["hydra"].main = lambda **kw: (lambda f: f)
_lm = types.ModuleType("lightning_model")
_lm.StyleTransferModel = type("StyleTransferModel", (), {})
sys.modules["lightning_model"] = _lm

spec = importlib.util.spec_from_file_location("ref_generator_driver", os.path.join(REF, "generator.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
from src.models.generator import GeneratorJ  # noqa: E402  (the reference's)


class _Model(torch.nn.Module):
    def __init__(self, g):
        super().__init__()
        self.generator = g


def make_instance(gen, patch):
    inst = object.__new__(ref.StyleTransferInference)
    inst.patch_size = patch
    inst.patch_positions = []
    inst.logger = logging.getLogger("ref")
    inst.cfg = types.SimpleNamespace(inference=types.SimpleNamespace(use_gpu=False))
    inst.model = _Model(gen)
    return inst


def main():
    torch.manual_seed(0)
    sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, "gen_c3_trained.npz")).items()}
    gen = GeneratorJ(input_channels=3, use_bias=True)
    gen.load_state_dict(sd, strict=True)
    gen.eval()
    vec = np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))
    frame = torch.from_numpy(vec["frame"])[:, :, :152, :212].contiguous()      # real-frame crop, [-1, 1]
    h, w = frame.shape[-2:]
    yy, xx = np.mgrid[0:h, 0:w]
    raw = ((((yy - 70) / 52.0) ** 2 + ((xx - 120) / 85.0) ** 2) <= 1).astype(np.float32)
    raw[:30, :44] = 1.0            # touches the top-left corner: clipped windows
    raw[h - 26:, w - 60:] = 1.0    # touches the bottom-right corner
    raw[60:64, 100:104] = 0.3      # below the 0.4 threshold
    raw_t = torch.from_numpy(raw)[None]
    out = {}
    inst = make_instance(gen, 32)
    mask = inst._process_mask(raw_t.clone())                                      # [1, H, W]
    out["frame"], out["mask_raw"], out["mask"] = frame.numpy(), raw, mask.numpy()
    for patch, overlap in ((32, 30.0), (80, 30.0), (48, 50.0)):
        inst = make_instance(gen, patch)
        pos = inst._get_valid_patch_positions(mask.unsqueeze(0), overlap_percent=overlap)
        y = inst.process_large_image(frame.clone(), mask.unsqueeze(0).clone(), overlap_percent=overlap)
        out[f"pos_p{patch}"] = np.asarray(pos, dtype=np.int32)
        out[f"y_p{patch}"] = y.numpy()
        print(f"patch {patch} overlap {overlap}: {len(pos)} windows, output range [{float(y.min()):.3f}, {float(y.max()):.3f}]")
    inst = make_instance(gen, 32)
    y = inst.process_large_image(frame.clone(), None, overlap_percent=30.0)      # no mask: every pixel is a candidate
    out["y_p32_nomask"] = y.numpy()
    np.savez_compressed(os.path.join(GOLD, "tiled_golden.npz"), **out)
    print("wrote", os.path.join(GOLD, "tiled_golden.npz"))


if __name__ == "__main__":
    logging.disable(logging.CRITICAL)
    main()
