"""Drop-in for the reference's generator.py (inference driver): ``StyleTransferInference(cfg).process_directory()``.

Default (north-star semantics): every frame goes through ONE full-frame ``GeneratorJ.forward`` on the GPU.
``inference.tiled=true`` selects the reference's own behaviour instead: the frame is cut into patch_size windows around
sampled mask pixels, every window is stylised separately (per-window InstanceNorm statistics) and the outputs are
blended with Gaussian weights (reference generator.py:327-565, native path in pbt_b200/tiled.py; fixtures from the
unmodified reference methods in tests/golden/tiled_golden.npz).  Pre/post-processing follows the reference: RGB + RGB-converted guide images
normalised to [-1,1] (:584-616), mask thresholded at 128 and eroded 7x7 (:327-351,627-631), composite
``rgb*(1-m) + out*m`` (:562-563), clamp / (x+1)*127.5 / round to uint8 (:643-647).
"""
import glob
import logging
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

from pbt_b200 import ops, tiled
from pbt_b200.config import compose, to_container
from pbt_b200.inference import FrameStylizer
from src.models.generator import GeneratorJ


class StyleTransferInference:
    def __init__(self, cfg):
        self.cfg = cfg
        self.debug_mode = bool(cfg.inference.get("debug_mode", False))
        logging.basicConfig(level=logging.DEBUG if self.debug_mode else logging.INFO,
                            format="%(asctime)s - %(levelname)s - %(message)s")
        self.logger = logging.getLogger(__name__)
        if not (torch.cuda.is_available() and cfg.inference.get("use_gpu", True)):
            raise RuntimeError("the B200-native inference path needs a CUDA device (inference.use_gpu=true); no CPU path")
        self.device = torch.device("cuda")
        self.additional_channels = to_container(dict(cfg.paths.get("additional_channels", {}) or {}))
        self.patch_size = cfg.data.patch_size
        self.tiled = bool(cfg.inference.get("tiled", False))
        self.overlap_percent = float(cfg.inference.get("overlap_percent", 30.0))   # reference default (:431)
        self._setup_model()

    def _setup_model(self):
        ckpt = torch.load(self.cfg.paths.checkpoint, map_location="cpu")
        sd = {k[len("generator."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("generator.")}
        cin = sd["initial_conv.0.weight"].shape[1]
        expect = 3 + 3 * len(self.additional_channels)  # every guide directory is RGB-converted (reference :92,606)
        if cin != expect:
            raise ValueError(f"checkpoint expects {cin} input channels, configuration provides {expect}")
        args = to_container(dict(self.cfg.model.generator.get("args", {})))
        args["input_channels"] = cin
        self.generator = GeneratorJ(**args)
        self.generator.load_state_dict(sd, strict=True)
        self.generator.to(self.device).eval()
        self.stylizer = FrameStylizer(self.generator)
        self.logger.info(f"generator loaded: {cin} input channels")

    @staticmethod
    def _find_corresponding_image(base_dir, image_path):
        if isinstance(base_dir, dict):
            base_dir = base_dir.get("path")
        stem = os.path.splitext(os.path.basename(image_path))[0]
        for ext in (".png", ".jpg", ".jpeg", ".PNG", ".JPG", ".JPEG"):
            p = os.path.join(str(base_dir), stem + ext)
            if os.path.exists(p):
                return p
        return os.path.join(str(base_dir), os.path.basename(image_path))

    def _process_mask(self, mask_path, h, w):
        m = Image.open(mask_path).point(lambda p: p > 128 and 255).convert("L")
        t = torch.from_numpy(np.asarray(m, dtype=np.uint8).copy()).to(self.device).float().div_(255.0)[None, None]
        box = F.conv2d(t, torch.ones((1, 1, 7, 7), device=self.device), padding=3)
        return torch.where(box < 49, torch.zeros_like(t), t)[:, :, :h, :w]   # true erosion: all 49 pixels set

    @torch.no_grad()
    def process_image(self, input_path, mask_path, save_path):
        imgs = [Image.open(input_path).convert("RGB")]
        for name, cdir in self.additional_channels.items():
            p = self._find_corresponding_image(cdir, input_path)
            if not os.path.exists(p):
                raise FileNotFoundError(f"Required channel {name} not found: {p}")
            imgs.append(Image.open(p).convert("RGB"))
        u8 = torch.from_numpy(np.concatenate([np.asarray(i, dtype=np.uint8) for i in imgs], axis=2)).to(self.device)
        h, w = u8.shape[0], u8.shape[1]
        if self.tiled:
            return self._process_image_tiled(u8, mask_path, save_path)
        ph, pw = (-h) % 4, (-w) % 4
        if ph or pw:  # GeneratorJ needs multiples of 4: replicate the border, crop afterwards
            u8 = F.pad(u8.permute(2, 0, 1)[None].float(), (0, pw, 0, ph), mode="replicate")[0].permute(1, 2, 0).to(torch.uint8)
        y = self.stylizer.eng.forward(u8[None].contiguous(), save=False, u8_hwc=True)[:, :, :h, :w]
        mask_file = self._find_corresponding_image(os.path.dirname(mask_path), mask_path)
        if not os.path.exists(mask_file):
            raise FileNotFoundError(f"Mask file not found: {mask_file}")
        m = self._process_mask(mask_file, h, w)
        rgb = ((u8[:h, :w, :3].permute(2, 0, 1)[None].float() / 255.0) - 0.5) / 0.5
        out = (rgb * (1 - m) + y * m).contiguous()
        res = torch.empty((1, h, w, 3), dtype=torch.uint8, device=self.device)
        ops.nchw_to_u8hwc(out, res)
        os.makedirs(os.path.dirname(save_path) or ".", exist_ok=True)
        Image.fromarray(res[0].cpu().numpy()).save(save_path)

    def _process_image_tiled(self, u8, mask_path, save_path):
        """reference behaviour (generator.py:567-652): windows -> generator -> Gaussian blend -> mask composite"""
        h, w = u8.shape[0], u8.shape[1]
        x = ((u8.permute(2, 0, 1)[None].float() / 255.0) - 0.5) / 0.5          # ToTensor + Normalize(0.5, 0.5) (:91-95)
        mask_file = self._find_corresponding_image(os.path.dirname(mask_path), mask_path)
        if not os.path.exists(mask_file):
            raise FileNotFoundError(f"Mask file not found: {mask_file}")
        m = Image.open(mask_file).point(lambda p: p > 128 and 255).convert("L")  # (:629-631, GrayscaleConvert + ToTensor)
        mt = torch.from_numpy(np.asarray(m, dtype=np.uint8).copy()).to(self.device).float().div_(255.0)[None]
        mt = tiled.process_mask(mt)[:, :h, :w].unsqueeze(0)
        out = tiled.process_large_image(self.generator, x, mt, int(self.patch_size), self.overlap_percent)
        res = torch.empty((1, h, w, 3), dtype=torch.uint8, device=self.device)
        ops.nchw_to_u8hwc(out.contiguous(), res)                                 # clamp, (x+1)*127.5, round (:643-647)
        os.makedirs(os.path.dirname(save_path) or ".", exist_ok=True)
        Image.fromarray(res[0].cpu().numpy()).save(save_path)

    def process_directory(self):
        paths = self.cfg.paths
        os.makedirs(paths.output_dir, exist_ok=True)
        files = sorted(glob.glob(os.path.join(paths.input_dir, "*.[pj][np][g]")))
        self.logger.info(f"Found {len(files)} images to process")
        for f in files:
            try:
                self.process_image(f, os.path.join(paths.mask_dir, os.path.basename(f)),
                                   os.path.join(paths.output_dir, os.path.basename(f)))
            except Exception as e:  # the reference logs and continues with the next frame (:700-705)
                self.logger.error(f"Failed to process {os.path.basename(f)}: {e}")


def main(cfg):
    StyleTransferInference(cfg).process_directory()


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(compose(os.path.join(here, "config"), "inference", sys.argv[1:]))
