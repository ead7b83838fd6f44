"""Drop-in ``StyleTransferDataset`` (reference src/data/dataset.py:13-298) with device-resident keyframes.

What stays identical to the reference (bit-exact for a given seed):
  * file discovery and pairing, RGB conversion, ToTensor+Normalize arithmetic (done by a CUDA kernel with
    IEEE division, so the resident fp32 images equal the reference's CPU tensors bit for bit)
  * the valid-centre list: mask.point(>128) -> 'L' -> 7x7 box-sum != 0 -> row-major (y, x)
  * the draw: ``np.random.randint(0, n_left)`` on numpy's global legacy RNG, WITHOUT replacement, the k-th
    smallest unused index is taken (list.pop on a sorted list in the reference), refill when empty,
    one extra draw when augmentation_factor > 1
  * `_cut_patch` clamping / top-left zero padding
What changes: the O(n) list.pop becomes an O(log n) order-statistics tree (native, host side), and the
patches of a whole batch are cut by ONE gather launch on the GPU, already concatenated as the generator
input (torch.cat of lightning_model.py:221) — see ``sample_batch``.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

from . import ops
from ._native import lib

_EXTS = (".png", ".jpg", ".jpeg", ".PNG", ".JPG", ".JPEG")


class _OsTree:
    """set of unused indices {0..n-1} with 'take k-th smallest' in O(log n) (libpbt host routine)"""

    def __init__(self, n: int):
        self.n = n
        self.tree = np.zeros(n + 1, dtype=np.int32)
        self._ptr = self.tree.ctypes.data
        self.reset()

    def reset(self):
        lib().pbt_ostree_reset(self._ptr, self.n)

    def __len__(self):
        return int(self.tree[0])

    def take(self, k: int) -> int:
        r = lib().pbt_ostree_take(self._ptr, self.n, k)
        if r < 0:
            raise IndexError(k)
        return r


class StyleTransferDataset(Dataset):
    def __init__(self, dir_pre: str, dir_post: str, dir_mask: str, patch_size: int, augmentation_factor: int = 1,
                 additional_channels: Optional[Dict[str, object]] = None, device: Optional[str] = None,
                 verbose: bool = False):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("StyleTransferDataset (B200-native) keeps its keyframes on a CUDA device; none is available")
        self.dir_pre, self.dir_post, self.dir_mask = dir_pre, dir_post, dir_mask
        self.patch_size = int(patch_size)
        self.additional_channels = dict(additional_channels or {})
        self.augmentation_factor = max(1, int(augmentation_factor))
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.verbose = verbose
        self.image_paths = sorted(f for f in os.listdir(dir_pre) if f.lower().endswith((".png", ".jpg", ".jpeg")))
        self.images_pre: List[torch.Tensor] = []
        self.images_post: List[torch.Tensor] = []
        self.valid_indices: List[torch.Tensor] = []          # int64 [K,2] (y,x), CPU
        self._valid_np: List[np.ndarray] = []
        self._left: List[_OsTree] = []
        self.additional_channel_data: Dict[str, List[torch.Tensor]] = {k: [] for k in self.additional_channels}
        self.last_patch_positions: List[List[int]] = []
        self._load_images()
        self._build_tables()

    # ------------------------------------------------------------------ loading
    @staticmethod
    def _find_corresponding_image(base_dir, image_name: str) -> str:
        if isinstance(base_dir, dict) or hasattr(base_dir, "get") and not isinstance(base_dir, str):
            base_dir = base_dir.get("path")
        stem = os.path.splitext(image_name)[0]
        for ext in _EXTS:
            p = os.path.join(base_dir, stem + ext)
            if os.path.exists(p):
                return p
        return os.path.join(base_dir, image_name)

    def _norm_u8(self, arr: np.ndarray) -> torch.Tensor:
        """uint8 HWC RGB -> resident fp32 CHW in [-1,1]: ToTensor + Normalize(0.5, 0.5) (reference dataset.py:34-38)"""
        u8 = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint8)).to(self.device)
        out = torch.empty((3, u8.shape[0], u8.shape[1]), device=self.device)
        ops.u8hwc_to_norm_chw(u8, out)
        return out

    def _to_device_norm(self, path: str) -> torch.Tensor:
        img = Image.open(path)
        if img.mode != "RGB":
            img = img.convert("RGB")
        return self._norm_u8(np.asarray(img, dtype=np.uint8).copy())

    def _valid_from_mask_u8(self, m: np.ndarray) -> torch.Tensor:
        """thresholded 'L' mask (0 / 255) -> int64 [K,2] (y,x) centres whose 7x7 window touches the mask (dataset.py:150-174)"""
        u8 = torch.from_numpy(np.ascontiguousarray(m, dtype=np.uint8)).to(self.device)
        dil = torch.empty_like(u8)
        ops.mask_dilate7(u8, dil)
        return dil.nonzero(as_tuple=False).cpu()

    def _valid_from_mask(self, path: str) -> torch.Tensor:
        m = Image.open(path)
        m = m.point(lambda p: p > 128 and 255)
        if m.mode != "L":
            m = m.convert("L")
        return self._valid_from_mask_u8(np.asarray(m, dtype=np.uint8).copy())

    @classmethod
    def from_arrays(cls, pre: Sequence[np.ndarray], post: Sequence[np.ndarray], mask: Sequence[np.ndarray], patch_size: int,
                    augmentation_factor: int = 1, additional: Optional[Dict[str, Sequence[np.ndarray]]] = None,
                    device: Optional[str] = None) -> "StyleTransferDataset":
        """the same dataset from decoded keyframes already in memory: uint8 [H,W,3] RGB arrays for pre / post / every
        guide and uint8 [H,W] single-band masks (thresholded at 128 like the file path).  Used where the frames come from
        a video decoder or a generator instead of image directories."""
        if not torch.cuda.is_available():
            raise RuntimeError("StyleTransferDataset (B200-native) keeps its keyframes on a CUDA device; none is available")
        self = cls.__new__(cls)
        Dataset.__init__(self)
        self.dir_pre = self.dir_post = self.dir_mask = None
        self.patch_size = int(patch_size)
        self.additional_channels = {k: {"depth": 3} for k in (additional or {})}
        self.augmentation_factor = max(1, int(augmentation_factor))
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.verbose = False
        self.image_paths = [f"{i:03d}" for i in range(len(pre))]
        self.images_pre, self.images_post, self.valid_indices, self._valid_np, self._left = [], [], [], [], []
        self.additional_channel_data = {k: [] for k in self.additional_channels}
        self.last_patch_positions = []
        for i in range(len(pre)):
            self.images_pre.append(self._norm_u8(pre[i]))
            self.images_post.append(self._norm_u8(post[i]))
            valid = self._valid_from_mask_u8(np.where(np.asarray(mask[i]) > 128, 255, 0).astype(np.uint8))
            self.valid_indices.append(valid)
            self._valid_np.append(valid.numpy())
            self._left.append(_OsTree(len(valid)))
            for k in self.additional_channels:
                self.additional_channel_data[k].append(self._norm_u8(additional[k][i]))
        self._build_tables()
        return self

    def _load_images(self):
        for name in self.image_paths:
            try:
                pre = self._to_device_norm(self._find_corresponding_image(self.dir_pre, name))
                post = self._to_device_norm(self._find_corresponding_image(self.dir_post, name))
                valid = self._valid_from_mask(self._find_corresponding_image(self.dir_mask, name))
                extra = {k: self._to_device_norm(self._find_corresponding_image(cfg, name))
                         for k, cfg in self.additional_channels.items()}
            except Exception as e:  # the reference skips an image set that fails to load (dataset.py:134-207)
                print(f"Error loading image set {name}: {e}")
                continue
            self.images_pre.append(pre)
            self.images_post.append(post)
            self.valid_indices.append(valid)
            self._valid_np.append(valid.numpy())
            self._left.append(_OsTree(len(valid)))
            for k, t in extra.items():
                self.additional_channel_data[k].append(t)
            if self.verbose:
                print(f"loaded {name}: {tuple(pre.shape)}, {len(valid)} valid patch positions")

    def _build_tables(self):
        n = len(self.images_pre)
        srcs = [self.images_pre] + [self.additional_channel_data[k] for k in self.additional_channels] + [self.images_post]
        self._n_src = len(srcs)
        self._table = torch.tensor([[t.data_ptr() for t in s] for s in srcs], dtype=torch.int64, device=self.device) \
            if n else None
        self._hw = torch.tensor([[t.shape[1], t.shape[2]] for t in self.images_pre], dtype=torch.int32, device=self.device) \
            if n else None
        # second table for the augmentation patches: [post | guides]
        srcs2 = [self.images_post] + [self.additional_channel_data[k] for k in self.additional_channels]
        self._table_aug = torch.tensor([[t.data_ptr() for t in s] for s in srcs2], dtype=torch.int64, device=self.device) \
            if n else None

    @property
    def valid_indices_left(self) -> List[int]:
        """number of not-yet-used centres per image (the reference keeps the explicit lists)"""
        return [len(t) for t in self._left]

    def __len__(self) -> int:
        return sum(len(v) for v in self.valid_indices) * self.augmentation_factor

    # ------------------------------------------------------------------ sampling
    def draw(self, idx: int):
        """RNG draw(s) for dataset index `idx`: ((img, y, x), (img, yr, xr) or None)"""
        img = idx % len(self.images_pre)
        left = self._left[img]
        if len(left) == 0:
            left.reset()
        c = np.random.randint(0, len(left))
        y, x = self._valid_np[img][left.take(c)]
        second = None
        if self.augmentation_factor > 1:
            r = np.random.randint(0, len(self._valid_np[img]))
            yr, xr = self._valid_np[img][r]
            second = (img, int(yr), int(xr))
        return (img, int(y), int(x)), second

    def _gather(self, table, n_src, pos: Sequence, outs, ch_off, ch_total):
        p = torch.tensor(pos, dtype=torch.int32).pin_memory().to(self.device, non_blocking=True)
        ops.patch_gather(table, n_src, len(self.images_pre), 3, self._hw, p, self.patch_size, outs, ch_off, ch_total)

    def sample_batch(self, indices: Sequence[int]) -> Dict[str, torch.Tensor]:
        """One gather launch for a whole batch.  Returns the reference's collated batch dict (views) plus
        'combined_input' = cat([pre] + guides, dim=1), contiguous, ready for the generator."""
        B, P = len(indices), self.patch_size
        k = len(self.additional_channels)
        pos, pos2 = [], []
        for idx in indices:
            a, b = self.draw(int(idx))
            pos.append(a)
            if b is not None:
                pos2.append(b)
        self.last_patch_positions = [[pos[-1][1], pos[-1][2]]] if pos else []
        comb = torch.empty((B, 3 + 3 * k, P, P), device=self.device)
        post = torch.empty((B, 3, P, P), device=self.device)
        if B:
            self._gather(self._table, self._n_src, pos, [comb] * (1 + k) + [post], [3 * i for i in range(1 + k)] + [0],
                         [3 + 3 * k] * (1 + k) + [3])
        batch = {"pre": comb[:, 0:3], "post": post, "combined_input": comb,
                 "positions": torch.tensor(pos, dtype=torch.int64).reshape(-1, 3)}
        for i, name in enumerate(self.additional_channels):
            batch[f"channel_{name}"] = comb[:, 3 + 3 * i: 6 + 3 * i]
        if pos2:
            self.last_patch_positions.append([pos2[-1][1], pos2[-1][2]])
            aug = torch.empty((B, 3 + 3 * k, P, P), device=self.device)
            self._gather(self._table_aug, 1 + k, pos2, [aug] * (1 + k), [3 * i for i in range(1 + k)], [3 + 3 * k] * (1 + k))
            batch["already"] = aug[:, 0:3]
            for i, name in enumerate(self.additional_channels):
                batch[f"channel_{name}_aug"] = aug[:, 3 + 3 * i: 6 + 3 * i]
        return batch

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        """reference-compatible single item (CUDA tensors [3,P,P]); training uses sample_batch instead"""
        b = self.sample_batch([idx])
        return {k: v[0] for k, v in b.items() if k not in ("combined_input", "positions")}
