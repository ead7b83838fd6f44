"""CPU restatement (test infrastructure only) of the reference's tiled inference mode, generator.py:327-565.

Pinned by tests/golden/tiled_golden.npz, which oracle/make_golden_tiled.py produces by calling the UNMODIFIED reference
methods `StyleTransferInference._process_mask / _get_valid_patch_positions / process_large_image` (imported in the build
container with stub modules for hydra / pytorch_lightning / omegaconf, which are not installed).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.nn.functional as F


def process_mask(mask_tensor: torch.Tensor) -> torch.Tensor:
    """generator.py:327-351 — threshold 0.4, 7x7 box sum, keep only fully covered pixels (erosion), value sum/49.
    mask_tensor: [1, H, W] in [0, 1].  Returns [1, H, W]."""
    m = mask_tensor.clone()
    m[m < 0.4] = 0
    w = torch.ones((1, 1, 7, 7), dtype=m.dtype)
    conv = F.conv2d(m.unsqueeze(0), w, stride=1, padding=3)
    conv[conv < w.numel()] = 0
    conv /= w.numel()
    return conv.squeeze(0)


def valid_patch_positions(mask_tensor: torch.Tensor, patch_size: int, overlap_percent: float = 50.0) -> List[Tuple[int, int, int, int]]:
    """generator.py:353-398 — every `stride`-th non-zero mask pixel (row-major) is a candidate centre; one window per
    (y // stride, x // stride) cell; windows are clipped to the frame: (y_start, y_end, x_start, x_end)."""
    overlap = min(max(overlap_percent, 0.0), 100.0) / 100.0
    stride = max(1, int(patch_size * (1 - overlap)))
    idx = mask_tensor.squeeze().nonzero()
    half = patch_size // 2
    h, w = mask_tensor.shape[-2:]
    out, used = [], set()
    for i in range(0, len(idx), stride):
        y, x = int(idx[i][0]), int(idx[i][1])
        key = (y // stride, x // stride)
        if key not in used:
            out.append((max(0, y - half), min(h, y + half), max(0, x - half), min(w, x + half)))
            used.add(key)
    return out


def ensure_valid_patch_size(patch: torch.Tensor, patch_size: int) -> torch.Tensor:
    """generator.py:470-497 — smaller windows are centred in a zero patch"""
    _, _, h, w = patch.shape
    if h == patch_size and w == patch_size:
        return patch
    new = torch.zeros((patch.size(0), patch.size(1), patch_size, patch_size), dtype=patch.dtype)
    hc, wc = min(h, patch_size), min(w, patch_size)
    ho, wo = (patch_size - hc) // 2, (patch_size - wc) // 2
    new[:, :, ho:ho + hc, wo:wo + wc] = patch[:, :, :hc, :wc]
    return new


def gaussian_weight(ph: int, pw: int, out_hw: Tuple[int, int], dtype=torch.float32) -> torch.Tensor:
    """generator.py:519-532 — separable Gaussian of the WINDOW size, bilinearly resized (align_corners=False) to the
    generator output size when the window was smaller than the patch.  Returns [1, 1, H, W]."""
    wy = torch.exp(-((torch.arange(ph) - ph / 2) ** 2 / (ph / 4) ** 2))[:, None]
    wx = torch.exp(-((torch.arange(pw) - pw / 2) ** 2 / (pw / 4) ** 2))[None, :]
    wt = (wy * wx).to(dtype)[None, None]
    if tuple(wt.shape[-2:]) != tuple(out_hw):
        wt = F.interpolate(wt, size=out_hw, mode="bilinear", align_corners=False)
    return wt


def process_large_image(generator: Callable[[torch.Tensor], torch.Tensor], input_tensor: torch.Tensor,
                        mask_tensor: Optional[torch.Tensor], patch_size: int, overlap_percent: float = 30.0):
    """generator.py:427-565.  input_tensor [1, C, H, W], mask_tensor [1, 1, H, W] (already through process_mask).
    Returns (output [1, 3, H, W], windows)."""
    b, c, h, w = input_tensor.shape
    dtype = input_tensor.dtype
    output = torch.zeros((b, 3, h, w), dtype=dtype)
    weights = torch.zeros((b, 1, h, w), dtype=dtype)
    if mask_tensor is None:
        mask_tensor = torch.ones((b, 1, h, w), dtype=dtype)
    boxes = valid_patch_positions(mask_tensor, patch_size, overlap_percent)
    for y0, y1, x0, x1 in boxes:
        patch = ensure_valid_patch_size(input_tensor[..., y0:y1, x0:x1], patch_size)
        with torch.no_grad():
            proc = ensure_valid_patch_size(generator(patch), patch_size)
        wt = gaussian_weight(y1 - y0, x1 - x0, tuple(proc.shape[-2:]), dtype)
        hs = slice(y0, min(y0 + proc.shape[2], h))
        ws = slice(x0, min(x0 + proc.shape[3], w))
        nh, nw = hs.stop - hs.start, ws.stop - ws.start
        output[..., hs, ws] += proc[..., :nh, :nw] * wt[..., :nh, :nw]
        weights[..., hs, ws] += wt[..., :nh, :nw]
    valid = weights > 1e-8
    output = output / weights.repeat(1, 3, 1, 1).where(valid, torch.ones_like(weights))
    output = input_tensor[:, :3] * (1 - mask_tensor) + output * mask_tensor
    return output, boxes
