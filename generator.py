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
from pbt_b200.parallel import dist_env, shard_range
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
        self.device = torch.device("cuda", dist_env()[2] % max(1, torch.cuda.device_count()))   # torchrun: LOCAL_RANK
        torch.cuda.set_device(self.device)
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

    # ------------------------------------------------------------------ frame pipeline: decode | GPU | encode
    def _load_frame(self, input_path, mask_path):
        """CPU stage (thread-safe, PIL releases the GIL while decoding): RGB frame + RGB-converted guides (:584-616),
        mask thresholded at 128 (:629-631) as uint8 arrays"""
        imgs = [Image.open(input_path).convert("RGB")]
        for name, cdir in self.additional_channels.items():
            p = self._find_corresponding_image(cdir, input_path)
            if not os.path.exists(p):
                raise FileNotFoundError(f"Required channel {name} not found: {p}")
            imgs.append(Image.open(p).convert("RGB"))
        u8 = np.concatenate([np.asarray(i, dtype=np.uint8) for i in imgs], axis=2)
        mask_file = self._find_corresponding_image(os.path.dirname(mask_path), mask_path)
        if not os.path.exists(mask_file):
            raise FileNotFoundError(f"Mask file not found: {mask_file}")
        m = Image.open(mask_file).point(lambda p: p > 128 and 255).convert("L")
        return u8, np.asarray(m, dtype=np.uint8).copy()

    def _erode(self, mask_u8):
        """thresholded uint8 mask [H,W] -> device fp32 [1,H,W]: 1 where all 49 pixels of the 7x7 window are set (:327-351)"""
        m = torch.from_numpy(mask_u8).to(self.device, non_blocking=True)[None].contiguous()
        out = torch.empty(m.shape, dtype=torch.float32, device=self.device)
        ops.mask_erode7(m, out)
        return out

    @torch.no_grad()
    def _stylize(self, u8_np, mask_np):
        """GPU stage: uint8 HWC frame (+guides) and mask -> uint8 HWC stylised frame (host array)"""
        u8 = torch.from_numpy(u8_np).to(self.device, non_blocking=True)
        h, w = u8.shape[0], u8.shape[1]
        if self.tiled:
            # reference behaviour (generator.py:567-652): windows -> generator -> Gaussian blend -> mask composite
            x = ((u8.permute(2, 0, 1)[None].float() / 255.0) - 0.5) / 0.5      # ToTensor + Normalize(0.5, 0.5) (:91-95)
            mt = self._erode(mask_np)[:, :h, :w].unsqueeze(0)      # the thresholded mask is binary: process_mask == erosion
            out = tiled.process_large_image(self.generator, x, mt, int(self.patch_size), self.overlap_percent)
            res = torch.empty((1, h, w, 3), dtype=torch.uint8, device=self.device)
            ops.composite_to_u8(out.contiguous(), res)                         # clamp, (x+1)*127.5, round (:643-647)
            return res[0].cpu().numpy()
        ph, pw = (-h) % 4, (-w) % 4
        m = self._erode(mask_np)
        if ph or pw:  # GeneratorJ needs multiples of 4: replicate the border, crop afterwards
            u8 = F.pad(u8.permute(2, 0, 1)[None].float(), (0, pw, 0, ph), mode="replicate")[0].permute(1, 2, 0).to(torch.uint8)
            m = F.pad(m, (0, pw, 0, ph))
        # generator pass, mask composite rgb*(1-m) + y*m (:562-563) and uint8 conversion (:643-647), all native
        res = self.stylizer.stylize_device(u8[None].contiguous(), masks=m.contiguous())
        return res[0, :h, :w].cpu().numpy()

    @staticmethod
    def _save(arr, save_path):
        os.makedirs(os.path.dirname(save_path) or ".", exist_ok=True)
        Image.fromarray(arr).save(save_path)

    def process_image(self, input_path, mask_path, save_path):
        u8, m = self._load_frame(input_path, mask_path)
        self._save(self._stylize(u8, m), save_path)

    def process_directory(self):
        """Frames are independent, so decode, GPU work and encode overlap: `inference.io_workers` threads decode ahead of
        the GPU and encode behind it (PNG codecs run at ~10-40 frames/s per core while the generator runs at ~200
        full-HD frames/s).  io_workers=0 is the reference's strictly sequential loop.
        Like the reference (:700-705), a frame that fails is logged and skipped."""
        from concurrent.futures import ThreadPoolExecutor
        paths = self.cfg.paths
        os.makedirs(paths.output_dir, exist_ok=True)
        files = sorted(glob.glob(os.path.join(paths.input_dir, "*.[pj][np][g]")))
        self.logger.info(f"Found {len(files)} images to process")
        jobs = [(f, os.path.join(paths.mask_dir, os.path.basename(f)), os.path.join(paths.output_dir, os.path.basename(f)))
                for f in files]
        # one process per GPU under torchrun: this rank takes its contiguous range of the frame list (no collective;
        # the reference's loop :674-705 is the world-size-1 case)
        rank, world, local = dist_env()
        if world > 1:
            lo, hi = shard_range(len(jobs), rank, world)
            self.logger.info(f"rank {rank}/{world}: frames [{lo}, {hi})")
            jobs = jobs[lo:hi]
        workers = int(self.cfg.inference.get("io_workers", min(16, os.cpu_count() or 1)))
        if workers <= 0:
            for f, mp, sp in jobs:
                try:
                    self.process_image(f, mp, sp)
                except Exception as e:
                    self.logger.error(f"Failed to process {os.path.basename(f)}: {e}")
            return
        ahead = 2 * workers
        with ThreadPoolExecutor(workers) as dec, ThreadPoolExecutor(workers) as enc:
            loads = {}
            saves = []
            nxt = 0
            for i, (f, mp, sp) in enumerate(jobs):
                while nxt < len(jobs) and nxt < i + ahead:          # keep the decode window full
                    loads[nxt] = dec.submit(self._load_frame, jobs[nxt][0], jobs[nxt][1])
                    nxt += 1
                try:
                    u8, m = loads.pop(i).result()
                    saves.append((f, enc.submit(self._save, self._stylize(u8, m), sp)))
                except Exception as e:
                    self.logger.error(f"Failed to process {os.path.basename(f)}: {e}")
                while len(saves) > ahead:                           # bound the frames held for encoding
                    f0, fut = saves.pop(0)
                    try:
                        fut.result()
                    except Exception as e:
                        self.logger.error(f"Failed to save {os.path.basename(f0)}: {e}")
            for f0, fut in saves:
                try:
                    fut.result()
                except Exception as e:
                    self.logger.error(f"Failed to save {os.path.basename(f0)}: {e}")


def main(cfg):
    StyleTransferInference(cfg).process_directory()


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(compose(os.path.join(here, "config"), "inference", sys.argv[1:]))
