"""ORACLE (test infrastructure — never imported by the product path).

numpy restatement of the reference patch sampler:

  load_rgb_norm   <- RGBConvert + ToTensor + Normalize(0.5, 0.5)          reference src/data/dataset.py:34-38
  valid_centres   <- mask.point(>128) -> 'L' -> 7x7 box sum != 0 -> nonzero  reference src/data/dataset.py:150-170
  cut_patch       <- StyleTransferDataset._cut_patch                      reference src/data/dataset.py:209-232
  OracleSampler   <- __getitem__ / __len__ (global numpy RNG, draw without replacement by list.pop)
                                                                          reference src/data/dataset.py:234-298

Pinning: tests/test_oracle.py replays tests/golden/sampler_golden.npz, recorded from the UNMODIFIED reference
StyleTransferDataset driven by torch DataLoader(shuffle=True, num_workers=0) (oracle/make_golden.py).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
from PIL import Image

_EXTS = (".png", ".jpg", ".jpeg", ".PNG", ".JPG", ".JPEG")


def find_image(base_dir: str, name: str) -> str:
    stem = os.path.splitext(name)[0]
    for e in _EXTS:
        p = os.path.join(base_dir, stem + e)
        if os.path.exists(p):
            return p
    return os.path.join(base_dir, name)


def load_rgb_norm(path: str) -> np.ndarray:
    """fp32 CHW in [-1,1]:  (u8/255 - 0.5)/0.5 with every step rounded to fp32 (as torch does)"""
    img = Image.open(path)
    if img.mode != "RGB":
        img = img.convert("RGB")
    a = np.asarray(img, dtype=np.uint8).transpose(2, 0, 1).astype(np.float32)
    a = a / np.float32(255.0)
    return ((a - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)


def load_mask_binary(path: str) -> np.ndarray:
    """uint8 HxW, 1 where the thresholded mask is non-zero after conversion to 'L'"""
    m = Image.open(path)
    m = m.point(lambda p: p > 128 and 255)
    if m.mode != "L":
        m = m.convert("L")
    return (np.asarray(m) > 0).astype(np.uint8)


def dilate7(binary: np.ndarray) -> np.ndarray:
    """1 where any pixel of the 7x7 window (zero padded) is set == (7x7 ones conv, pad 3) != 0"""
    h, w = binary.shape
    p = np.zeros((h + 6, w + 6), dtype=np.int32)
    p[3:3 + h, 3:3 + w] = binary
    ii = np.zeros((h + 7, w + 7), dtype=np.int64)
    ii[1:, 1:] = p.cumsum(0).cumsum(1)
    s = ii[7:, 7:] - ii[:-7, 7:] - ii[7:, :-7] + ii[:-7, :-7]
    return (s > 0).astype(np.uint8)


def valid_centres(mask_path: str) -> np.ndarray:
    """int64 [K,2] (y,x) in row-major order"""
    return np.argwhere(dilate7(load_mask_binary(mask_path)) > 0).astype(np.int64)


def cut_patch(t: np.ndarray, y: int, x: int, size: int) -> np.ndarray:
    half = size // 2
    hn, hx = max(0, y - half), min(y + half, t.shape[1] - 1)
    xn, xx = max(0, x - half), min(x + half, t.shape[2] - 1)
    patch = t[:, hn:hx, xn:xx]
    if patch.shape[1] != size or patch.shape[2] != size:
        res = np.zeros((t.shape[0], size, size), dtype=np.float32)
        res[:, :patch.shape[1], :patch.shape[2]] = patch
        patch = res
    return patch


class OracleSampler:
    def __init__(self, dir_pre: str, dir_post: str, dir_mask: str, patch_size: int, augmentation_factor: int = 1,
                 additional_channels: Optional[Dict[str, object]] = None):
        self.patch_size = patch_size
        self.aug = max(1, augmentation_factor)
        self.channels = dict(additional_channels or {})
        names = sorted(f for f in os.listdir(dir_pre) if f.lower().endswith((".png", ".jpg", ".jpeg")))
        self.pre: List[np.ndarray] = []
        self.post: List[np.ndarray] = []
        self.valid: List[np.ndarray] = []
        self.left: List[List[int]] = []
        self.extra: Dict[str, List[np.ndarray]] = {k: [] for k in self.channels}
        for nm in names:
            self.pre.append(load_rgb_norm(find_image(dir_pre, nm)))
            self.post.append(load_rgb_norm(find_image(dir_post, nm)))
            v = valid_centres(find_image(dir_mask, nm))
            self.valid.append(v)
            self.left.append(list(range(len(v))))
            for k, cfg in self.channels.items():
                d = cfg.get("path") if isinstance(cfg, dict) else cfg
                self.extra[k].append(load_rgb_norm(find_image(d, nm)))
        self.last_patch_positions: List[List[int]] = []

    def __len__(self) -> int:
        return sum(len(v) for v in self.valid) * self.aug

    def draw(self, idx: int):
        """RNG + bookkeeping only: returns (img, y, x)"""
        i = idx % len(self.pre)
        if not self.left[i]:
            self.left[i] = list(range(len(self.valid[i])))
        c = np.random.randint(0, len(self.left[i]))
        y, x = self.valid[i][self.left[i][c]]
        self.left[i].pop(c)
        return i, int(y), int(x)

    def __getitem__(self, idx: int) -> Dict[str, np.ndarray]:
        i, y, x = self.draw(idx)
        self.last_patch_positions = [[y, x]]
        P = self.patch_size
        out = {"pre": cut_patch(self.pre[i], y, x, P), "post": cut_patch(self.post[i], y, x, P)}
        for k in self.channels:
            out[f"channel_{k}"] = cut_patch(self.extra[k][i], y, x, P)
        if self.aug > 1:
            r = np.random.randint(0, len(self.valid[i]))
            yr, xr = (int(v) for v in self.valid[i][r])
            self.last_patch_positions.append([yr, xr])
            out["already"] = cut_patch(self.post[i], yr, xr, P)
            for k in self.channels:
                out[f"channel_{k}_aug"] = cut_patch(self.extra[k][i], yr, xr, P)
        return out
