"""Tiled inference mode of the reference driver (reference generator.py:327-565), native path.

The reference stylises a frame window by window: `_get_valid_patch_positions` samples window centres from the eroded
mask, every patch_size window goes through the generator on its own (so InstanceNorm statistics are per window), and
the outputs are blended with Gaussian weights, normalised and composited with the input through the mask.
Here the windows of a frame are gathered into batches by one native launch (`pbt_tile_gather`), go through the
B200-native GeneratorJ as a batch (InstanceNorm is per sample, so batching does not change the result), and are blended
/ normalised / composited by `pbt_tile_blend` and `pbt_tile_finish`.  Window selection is host logic and follows the
reference literally, including its quirks (a window smaller than the patch is centred in a zero patch, but its output
is blended from the window's top-left corner; its weight map is the bilinearly resized Gaussian of the window size).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from ._native import check, lib, stream_ptr

Box = Tuple[int, int, int, int]


def process_mask(mask_tensor: torch.Tensor) -> torch.Tensor:
    """reference generator.py:327-351: threshold 0.4, 7x7 box sum (zero padded), zero where the sum is below 49, divide by
    49.  For mask values in [0, 1] (ToTensor output) the sum reaches 49 only where all 49 pixels equal 1.0, so the result
    is the binary erosion of (mask == 1) - computed by the native `pbt_mask_erode7`.
    mask_tensor [1, H, W] on the GPU, values in [0, 1]."""
    from . import ops
    full = (mask_tensor >= 1.0).to(torch.uint8).contiguous()
    out = torch.empty(full.shape, dtype=torch.float32, device=full.device)
    ops.mask_erode7(full, out)
    return out.to(mask_tensor.dtype)


def valid_patch_positions(mask_tensor: torch.Tensor, patch_size: int, overlap_percent: float = 50.0) -> List[Box]:
    """reference generator.py:353-398 (host logic): every `stride`-th non-zero mask pixel is a candidate centre, one
    window per (y // stride, x // stride) cell, windows clipped to the frame"""
    overlap = min(max(overlap_percent, 0.0), 100.0) / 100.0
    stride = max(1, int(patch_size * (1 - overlap)))
    idx = mask_tensor.squeeze().nonzero()[::stride].cpu().tolist()   # only the sampled candidates cross to the host
    half = patch_size // 2
    h, w = mask_tensor.shape[-2:]
    out: List[Box] = []
    used = set()
    for y, x in idx:
        key = (y // stride, x // stride)
        if key not in used:
            out.append((max(0, y - half), min(h, y + half), max(0, x - half), min(w, x + half)))
            used.add(key)
    return out


def _weight_table(boxes: Sequence[Box], patch: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Gaussian blend weights per distinct window shape (reference generator.py:519-532), and the shape index per window"""
    shapes: Dict[Tuple[int, int], int] = {}
    maps, index = [], []
    for y0, y1, x0, x1 in boxes:
        key = (y1 - y0, x1 - x0)
        if key not in shapes:
            ph, pw = key
            wy = torch.exp(-((torch.arange(ph, device=device) - ph / 2) ** 2 / (ph / 4) ** 2))[:, None]
            wx = torch.exp(-((torch.arange(pw, device=device) - pw / 2) ** 2 / (pw / 4) ** 2))[None, :]
            wt = (wy * wx).float()[None, None]
            if (ph, pw) != (patch, patch):
                wt = F.interpolate(wt, size=(patch, patch), mode="bilinear", align_corners=False)
            shapes[key] = len(maps)
            maps.append(wt[0, 0])
        index.append(shapes[key])
    return torch.stack(maps).contiguous(), torch.tensor(index, dtype=torch.int32, device=device)


@torch.no_grad()
def process_large_image(generator, input_tensor: torch.Tensor, mask_tensor: Optional[torch.Tensor], patch_size: int,
                        overlap_percent: float = 30.0, tile_batch: int = 128,
                        return_windows: bool = False):
    """reference generator.py:427-565.  input_tensor [1, C, H, W] fp32 on the GPU in [-1, 1]; mask_tensor [1, 1, H, W]
    (already through `process_mask`) or None.  Returns [1, 3, H, W] fp32."""
    if not input_tensor.is_cuda:
        raise RuntimeError("tiled inference runs on the GPU only (no CPU path)")
    b, c, h, w = input_tensor.shape
    if b != 1:
        raise ValueError("process_large_image handles one frame at a time (as the reference does)")
    if patch_size % 4:
        raise ValueError("patch_size must be a multiple of 4")
    dev = input_tensor.device
    x = input_tensor.float().contiguous()
    if mask_tensor is None:
        mask_tensor = torch.ones((1, 1, h, w), device=dev)
    mask = mask_tensor.float().contiguous()
    boxes = valid_patch_positions(mask, patch_size, overlap_percent)
    acc = torch.zeros((3, h, w), device=dev)
    wsum = torch.zeros((h, w), device=dev)
    if boxes:
        wtab, widx = _weight_table(boxes, patch_size, dev)
        boxes_dev = torch.tensor(boxes, dtype=torch.int32, device=dev)
        was_training = generator.training
        generator.eval()
        try:
            for lo in range(0, len(boxes), tile_batch):
                n = min(tile_batch, len(boxes) - lo)
                tiles = torch.empty((n, c, patch_size, patch_size), device=dev)
                bsl = boxes_dev[lo:lo + n].contiguous()
                check(lib().pbt_tile_gather(x.data_ptr(), c, h, w, bsl.data_ptr(), n, patch_size, tiles.data_ptr(), stream_ptr()),
                      "pbt_tile_gather")
                proc = generator(tiles).float().contiguous()          # [n, 3, P, P]; InstanceNorm is per window
                check(lib().pbt_tile_blend(proc.data_ptr(), bsl.data_ptr(), widx[lo:lo + n].contiguous().data_ptr(), wtab.data_ptr(),
                                           n, patch_size, h, w, acc.data_ptr(), wsum.data_ptr(), stream_ptr()), "pbt_tile_blend")
        finally:
            generator.train(was_training)
    out = torch.empty((1, 3, h, w), device=dev)
    rgb = x[0, :3].contiguous()
    check(lib().pbt_tile_finish(acc.data_ptr(), wsum.data_ptr(), rgb.data_ptr(), mask.data_ptr(), h, w, out.data_ptr(),
                                stream_ptr()), "pbt_tile_finish")
    return (out, boxes) if return_windows else out
